"""B200-native path-tracing core — Python face of the C ABI in include/rt_b200.h."""
from . import ctypes_types, scene_desc, scenes, sharding  # noqa: F401
from .api import (COMM_ID_BYTES, Context, Group, RtError, Scene, SYMBOLS, build, comm_unique_id, device_count, lib,  # noqa: F401
                  save_ppm, shard_sample_range)  # noqa: F401
from .ctypes_types import MODE_AUTO, MODE_EXACT, MODE_FAST, SCHED_AUTO, SCHED_MEGAKERNEL, SCHED_WAVEFRONT  # noqa: F401
