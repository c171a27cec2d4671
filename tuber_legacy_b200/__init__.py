"""tuber-b200: the per-frame data-parallel hot path of Lisible/tuber-legacy (ECS query over
Transform + Sprite, Transform -> Matrix4 compose with parent propagation, sprite batching into
instance/vertex buffers) as hand-written sm_100a CUDA behind a C ABI (include/tuber_b200.h).

The product is csrc/ + the shared library; `capi` is its ctypes binding, `scenes` the seeded
synthetic scenes of the BASELINE configs. There is no CPU fallback.
"""
__version__ = "0.1.0"
