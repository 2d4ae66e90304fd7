"""B200-native path-tracing core — Python face of the C ABI in include/rt_b200.h."""
from . import ctypes_types, scene_desc, scenes, sharding  # noqa: F401
from .api import Context, RtError, Scene, SYMBOLS, build, device_count, lib, save_ppm  # noqa: F401
from .ctypes_types import MODE_AUTO, MODE_EXACT, MODE_FAST, SCHED_AUTO, SCHED_MEGAKERNEL, SCHED_WAVEFRONT  # noqa: F401
