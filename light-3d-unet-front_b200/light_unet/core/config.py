"""YAML config helper with the reference's semantics (light_unet/core/config.py:13-27):
plain nested dict in, plain nested dict out."""
import os

import yaml


class ConfigManager:
    @staticmethod
    def load(path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"Config file not found: {path}")
        with open(path, "r", encoding="utf-8") as fh:
            return yaml.safe_load(fh)

    @staticmethod
    def save(config, path):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w", encoding="utf-8") as fh:
            yaml.dump(config, fh, default_flow_style=False)
