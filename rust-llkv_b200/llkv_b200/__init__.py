"""llkv_b200 — host-side mirror of LLKV's scan/filter/MVCC/aggregate interfaces over the B200 C ABI.

`expr` / `table` only build and flatten trees and buffers (importable without a GPU).
`gpu` binds libllkv_gpu.so; it raises loudly when the CUDA library or a device is missing (no CPU fallback).
"""
from . import ffi  # noqa: F401
from .expr import *  # noqa: F401,F403
from .table import *  # noqa: F401,F403
