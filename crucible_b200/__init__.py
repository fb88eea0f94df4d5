"""crucible_b200 — B200-native (sm_100a) path-tracing backend for the Crucible renderer.

Layout: csrc/ (CUDA kernels + the C ABI of include/crucible_gpu.h), abi.py (ctypes mirror),
scene.py / demo_builder.py (host-side mirror of the reference's scene API), gpu.py (the backend behind
`Camera::render`), multigpu.py (row / frame sharding over torch.distributed).
"""
from . import abi  # noqa: F401

__version__ = "0.1.0"
