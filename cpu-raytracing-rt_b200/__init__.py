"""B200-native path-tracing core — Python face of the C ABI in include/rt_b200.h."""
from . import ctypes_types, scene_desc, scenes  # noqa: F401
