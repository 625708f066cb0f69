"""realtrace_b200 — B200-native ray-tracing core behind RealTrace's Serial scene/render API.

The product is the C-ABI CUDA library (include/realtrace_b200.h, realtrace_b200/csrc) and the
C++ mirror of the reference's classes (realtrace_b200/host).  This Python package is the harness
side only: ctypes binding, scene records and the named workloads of BASELINE.json.
"""
from . import api, objio, scene, scenes  # noqa: F401

__all__ = ["api", "objio", "scene", "scenes"]
