"""shockwave_b200 — B200-native (sm_100a) per-round schedule solver behind Shockwave's own Python API.

Public surface (drop-in for the reference's scheduler/shockwave.py and scheduler/policies/*):
    ShockwaveScheduler      shockwave_b200.scheduler
    Engine, make_params     shockwave_b200.engine   (ctypes binding of libswb200.so, include/swb200.h)
"""
from .engine import Engine, LIB_PATH, load_library, make_params  # noqa: F401
from .scheduler import ShockwaveScheduler  # noqa: F401

__all__ = ["ShockwaveScheduler", "Engine", "make_params", "load_library", "LIB_PATH"]
