from .inferencer import Inferencer
from .config import ConfigManager

__all__ = ["Inferencer", "ConfigManager"]
