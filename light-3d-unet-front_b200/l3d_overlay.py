"""Lay the B200-native modules over a Light-3D-Unet checkout, module by module.

    import l3d_overlay; l3d_overlay.install("/path/to/Light-3D-Unet-Front")
    from light_unet.core.trainer import Trainer        # the reference's own Trainer ...
    # ... whose imports of light_unet.models.{unet3d,losses,metrics}, light_unet.utils and light_unet.core.inferencer now
    # resolve to this package, while light_unet.datasets.{loader,case_dataset,patch_dataset}, core.trainer, scripts/* stay
    # the reference's.

Why not sys.path order: both trees are REGULAR packages named light_unet; whichever comes first on sys.path shadows the
other completely, so the reference's trainer.py (which imports light_unet.datasets.loader, trainer.py:16) fails to import
with this package in front.  The finder below resolves every `light_unet.*` name against BOTH trees: a module that exists
here wins, anything else comes from the reference; a package uses the reference's __init__.py when it has one (so its
re-exports keep working, models/__init__.py:18-24) and searches both directories for its submodules.
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG = "light_unet"


class _OverlayFinder(importlib.abc.MetaPathFinder):
    def __init__(self, reference_root: str):
        self.roots = [os.path.join(_HERE, PKG), os.path.join(os.path.abspath(reference_root), PKG)]   # ours first
        if not os.path.isdir(self.roots[1]):
            raise FileNotFoundError(f"{self.roots[1]}: not a Light-3D-Unet checkout (no light_unet/ package)")

    def find_spec(self, fullname, path=None, target=None):
        if fullname != PKG and not fullname.startswith(PKG + "."):
            return None
        rel = fullname.split(".")[1:]
        dirs = [os.path.join(r, *rel) for r in self.roots]
        pkg_dirs = [d for d in dirs if os.path.isdir(d)]
        if pkg_dirs:
            # a package: the reference's __init__ when it exists (index 1), else ours; submodules from both trees
            inits = [os.path.join(d, "__init__.py") for d in reversed(dirs) if os.path.isfile(os.path.join(d, "__init__.py"))]
            if inits:
                return importlib.util.spec_from_file_location(fullname, inits[0], submodule_search_locations=pkg_dirs)
        for d in dirs:                                   # a module: ours wins
            if os.path.isfile(d + ".py"):
                return importlib.util.spec_from_file_location(fullname, d + ".py")
        return None


def install(reference_root: str) -> None:
    """Idempotent; must run before the first `import light_unet`."""
    if any(isinstance(f, _OverlayFinder) for f in sys.meta_path):
        return
    loaded = [m for m in sys.modules if m == PKG or m.startswith(PKG + ".")]
    if loaded:
        raise RuntimeError(f"l3d_overlay.install() must run before light_unet is imported (already loaded: {loaded[:3]} ...)")
    sys.meta_path.insert(0, _OverlayFinder(reference_root))


def uninstall() -> None:
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _OverlayFinder)]
    for m in [m for m in sys.modules if m == PKG or m.startswith(PKG + ".")]:
        del sys.modules[m]
