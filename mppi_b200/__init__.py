"""Importable alias of the `husky-rover-mppi-isaacsim_b200/` package directory (whose name is not a
valid Python identifier).  All code lives there; this module only points `__path__` at it."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "husky-rover-mppi-isaacsim_b200")
__path__.insert(0, _PKG_DIR)

from . import capi                                            # noqa: E402
from .controller import MPPI_Controller, Robot, Surface       # noqa: E402
from .devarray import DeviceArray                             # noqa: E402
from .costmap import build_obstacle_costmap                   # noqa: E402

DEFAULT_CONFIG = _os.path.join(_PKG_DIR, "config", "default.yaml")
__all__ = ["capi", "MPPI_Controller", "Robot", "Surface", "DeviceArray", "DEFAULT_CONFIG", "build_obstacle_costmap"]
