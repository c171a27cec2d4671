"""ctypes binding of libtuber_b200.so (include/tuber_b200.h).

This is the same boundary the reference's Rust FFI crate would bind: plain
pointers and sizes, no torch types. There is no fallback: if the library is missing or no
sm_100 GPU is present, loading / context creation raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

TB_OK = 0
TB_ERR_INVALID, TB_ERR_CUDA, TB_ERR_OOM, TB_ERR_RANGE, TB_ERR_NO_DEVICE = 1, 2, 3, 4, 5
TB_ERR_CAPACITY, TB_ERR_CYCLE, TB_ERR_COMM, TB_ERR_UNSUPPORTED = 6, 7, 8, 9
TB_TRANSFORM, TB_SPRITE, TB_VELOCITY, TB_PARENT, TB_SPIN = 0, 1, 2, 3, 4
TB_SYSTEM_VELOCITY, TB_SYSTEM_SPIN = 0, 1
TB_PATH_NONE, TB_PATH_BUCKET, TB_PATH_SORT = 0, 1, 2
TB_NO_PARENT = 0xFFFFFFFF

# numpy views of the C structs
TRANSFORM_DTYPE = np.dtype([("translation", "<f4", 3), ("angle", "<f4", 3),
                            ("rotation_center", "<f4", 3), ("scale", "<f4", 3)])
SPRITE_DTYPE = np.dtype([("size", "<f4", 2), ("texture", "<u4"), ("layer", "<u4")])
VELOCITY_DTYPE = np.dtype([("v", "<f4", 3)])
INSTANCE_DTYPE = np.dtype([("model", "<f4", 16), ("size", "<f4", 2), ("texture", "<u4"), ("layer", "<u4")])
VERTEX_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("uv", "<u4")])
BATCH_DTYPE = np.dtype([("layer", "<u4"), ("texture", "<u4"), ("first_instance", "<u4"), ("count", "<u4")])
assert TRANSFORM_DTYPE.itemsize == 48 and SPRITE_DTYPE.itemsize == 16 and INSTANCE_DTYPE.itemsize == 80
assert VERTEX_DTYPE.itemsize == 16 and BATCH_DTYPE.itemsize == 16


class FrameOut(C.Structure):
    _fields_ = [
        ("num_instances", C.c_uint64),
        ("num_batches", C.c_uint64),
        ("d_instances", C.c_void_p),
        ("d_vertices", C.c_void_p),
        ("d_batches", C.c_void_p),
        ("d_order", C.c_void_p),
        ("path", C.c_uint32),
        ("key_bits", C.c_uint32),
        ("sort_passes", C.c_uint32),
        ("kernel_launches", C.c_uint32),
        ("hierarchy_levels", C.c_uint32),
        ("replanned", C.c_uint32),
    ]


class TuberError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"tb_status {status}: {message}")
        self.status = status


_SIGNATURES = {
    "tb_version": (C.c_char_p, []),
    "tb_status_string": (C.c_char_p, [C.c_int32]),
    "tb_ctx_create": (C.c_int32, [C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]),
    "tb_ctx_destroy": (None, [C.c_void_p]),
    "tb_last_error": (C.c_char_p, [C.c_void_p]),
    "tb_ctx_set_stream": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "tb_ctx_synchronize": (C.c_int32, [C.c_void_p]),
    "tb_ctx_set_max_batches": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "tb_ctx_set_option": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int64]),
    "tb_set_entity_count": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "tb_entity_count": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "tb_upload_transforms": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_upload_sprites": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_upload_velocities": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_upload_parents": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_upload_spins": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_insert": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "tb_delete_by_ids": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "tb_delete_by_query": (C.c_int32, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
    "tb_add_component": (C.c_int32, [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]),
    "tb_remove_component": (C.c_int32, [C.c_void_p, C.c_int, C.c_uint64]),
    "tb_entity_generations": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_entities_alive": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_system_spin": (C.c_int32, [C.c_void_p, C.c_float]),
    "tb_system_access": (C.c_int32, [C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "tb_bundle_step": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_void_p]),
    "tb_upload_presence": (C.c_int32, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_download_presence": (C.c_int32, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_download_transforms": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_query": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "tb_as_matrix4": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_quat_from_euler": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_quat_from_axis_angle": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_quat_rotation_matrix": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_quat_mul": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_quat_normalize": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_mat4_mul": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_mat4_transform_vec3": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_mat4_try_inverse": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "tb_system_integrate": (C.c_int32, [C.c_void_p, C.c_float]),
    "tb_frame_transform_batch": (C.c_int32, [C.c_void_p, C.POINTER(FrameOut)]),
    "tb_frame_force_path": (C.c_int32, [C.c_void_p, C.c_uint32]),
    "tb_frame_set_view_projection": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "tb_orthographic": (C.c_int32, [C.c_float] * 6 + [C.c_void_p]),
    "tb_download_instances": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_download_vertices": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_download_batches": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_download_order": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_download_world_matrices": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_present_headless": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64)]),
    "tb_profile_enable": (C.c_int32, [C.c_void_p, C.c_int]),
    "tb_profile_count": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "tb_profile_get": (C.c_int32, [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "tb_launch_count": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "tb_shard_set": (C.c_int32, [C.c_void_p, C.c_int, C.c_int, C.c_uint64]),
    "tb_frame_set_output": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "tb_shard_key_stats": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "tb_shard_plan": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tb_shard_key_histogram": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "tb_shard_place": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "tb_shard_unique_id": (C.c_int32, [C.c_void_p]),
    "tb_shard_local_id": (C.c_int32, [C.c_void_p]),
    "tb_shard_init": (C.c_int32, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_int]),
    "tb_frame_sharded": (C.c_int32, [C.c_void_p, C.POINTER(FrameOut)]),
    "tb_shard_finalize": (C.c_int32, [C.c_void_p]),
    "tb_device_alloc": (C.c_int32, [C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]),
    "tb_device_free": (C.c_int32, [C.c_void_p]),
    "tb_ipc_export": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "tb_ipc_open": (C.c_int32, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "tb_ipc_close": (C.c_int32, [C.c_void_p]),
    "tb_device_read": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tb_vertices_from_instances": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "tb_host_alloc": (C.c_int32, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "tb_host_free": (None, [C.c_void_p]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Loads (building first if the sources are newer) the CUDA library. Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TB_LIB_PATH")  # a deployed build elsewhere; default: the in-tree library
    if not path:
        path = _build.LIB_PATH
        if not os.path.exists(path) or os.environ.get("TB_REBUILD"):
            path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    return list(_SIGNATURES.keys())


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _as(a, dtype):
    """`a` as a contiguous 1-D array of `dtype` records: either already that dtype, or a
    2-D array of 4-byte scalars whose rows are exactly one record."""
    a = np.ascontiguousarray(a)
    if a.dtype == dtype:
        return a.reshape(-1)
    if a.ndim == 2 and a.dtype.itemsize == 4 and a.shape[1] * 4 == dtype.itemsize:
        return a.view(dtype).reshape(-1)
    raise TypeError(f"cannot view array of dtype {a.dtype} shape {a.shape} as {dtype}")


class Context:
    """One tb_ctx: the device-resident ECS columns and frame pipeline of one GPU."""

    def __init__(self, max_entities, device=0, **options):
        self._lib = load()
        h = C.c_void_p()
        st = self._lib.tb_ctx_create(int(device), int(max_entities), C.byref(h))
        if st != TB_OK:
            raise TuberError(st, self._lib.tb_status_string(st).decode())
        self._h = h
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def _check(self, st):
        if st != TB_OK:
            raise TuberError(st, self._lib.tb_last_error(self._h).decode() or self._lib.tb_status_string(st).decode())

    # -- context
    def set_option(self, name, value):
        self._check(self._lib.tb_ctx_set_option(self._h, name.encode(), int(value)))

    def set_stream(self, cuda_stream):
        self._check(self._lib.tb_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(self._lib.tb_ctx_synchronize(self._h))

    def set_max_batches(self, n):
        self._check(self._lib.tb_ctx_set_max_batches(self._h, int(n)))

    # -- storage
    def set_entity_count(self, n):
        self._check(self._lib.tb_set_entity_count(self._h, int(n)))

    def entity_count(self):
        n = C.c_uint64()
        self._check(self._lib.tb_entity_count(self._h, C.byref(n)))
        return n.value

    def upload_transforms(self, first, transforms):
        a = _as(transforms, TRANSFORM_DTYPE)
        self._check(self._lib.tb_upload_transforms(self._h, int(first), a.shape[0], _ptr(a)))

    def upload_sprites(self, first, sprites):
        a = _as(sprites, SPRITE_DTYPE)
        self._check(self._lib.tb_upload_sprites(self._h, int(first), a.shape[0], _ptr(a)))

    def upload_velocities(self, first, velocities):
        a = _as(velocities, VELOCITY_DTYPE)
        self._check(self._lib.tb_upload_velocities(self._h, int(first), a.shape[0], _ptr(a)))

    def upload_parents(self, first, parents):
        a = np.ascontiguousarray(parents, dtype=np.uint32)
        self._check(self._lib.tb_upload_parents(self._h, int(first), a.shape[0], _ptr(a)))

    def upload_presence(self, component, first_word, words):
        a = np.ascontiguousarray(words, dtype=np.uint64)
        self._check(self._lib.tb_upload_presence(self._h, int(component), int(first_word), a.shape[0], _ptr(a)))

    def download_presence(self, component, first_word, nwords):
        a = np.zeros(int(nwords), dtype=np.uint64)
        self._check(self._lib.tb_download_presence(self._h, int(component), int(first_word), int(nwords), _ptr(a)))
        return a

    def download_transforms(self, first, n):
        a = np.zeros(int(n), dtype=TRANSFORM_DTYPE)
        self._check(self._lib.tb_download_transforms(self._h, int(first), int(n), _ptr(a)))
        return a

    # -- query / compose / systems
    def query(self, required_components, want_indices=True):
        """Ascending indices of entities holding every component in `required_components`."""
        bits = 0
        for c in required_components:
            bits |= 1 << c
        count = C.c_uint64()
        cap = self.entity_count() if want_indices else 0
        out = np.zeros(max(cap, 1), dtype=np.uint32)
        self._check(self._lib.tb_query(self._h, bits, _ptr(out) if want_indices else C.c_void_p(0), cap, C.byref(count)))
        return out[:count.value] if want_indices else count.value

    def as_matrix4(self, transforms):
        a = _as(transforms, TRANSFORM_DTYPE)
        out = np.zeros((a.shape[0], 16), dtype=np.float32)
        self._check(self._lib.tb_as_matrix4(self._h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    # -- the rest of tuber-math, batched (f32 arrays; quaternions (w, x, y, z), matrices row-major)
    def _math(self, fn, ins, out_shapes):
        ins = [np.ascontiguousarray(a, dtype=np.float32) for a in ins]
        n = ins[0].shape[0]
        outs = [np.zeros((n,) + tuple(sh), dtype=dt) for sh, dt in out_shapes]
        self._check(fn(self._h, *[_ptr(a) for a in ins], n, *[_ptr(o) for o in outs]))
        return outs[0] if len(outs) == 1 else tuple(outs)

    def quat_from_euler(self, euler_xyz):
        return self._math(self._lib.tb_quat_from_euler, [np.reshape(euler_xyz, (-1, 3))], [((4,), np.float32)])

    def quat_from_axis_angle(self, axis_angle):
        return self._math(self._lib.tb_quat_from_axis_angle, [np.reshape(axis_angle, (-1, 4))], [((4,), np.float32)])

    def quat_rotation_matrix(self, q):
        return self._math(self._lib.tb_quat_rotation_matrix, [np.reshape(q, (-1, 4))], [((16,), np.float32)])

    def quat_mul(self, a, b):
        return self._math(self._lib.tb_quat_mul, [np.reshape(a, (-1, 4)), np.reshape(b, (-1, 4))], [((4,), np.float32)])

    def quat_normalize(self, q):
        return self._math(self._lib.tb_quat_normalize, [np.reshape(q, (-1, 4))], [((4,), np.float32)])

    def mat4_mul(self, a, b):
        return self._math(self._lib.tb_mat4_mul, [np.reshape(a, (-1, 16)), np.reshape(b, (-1, 16))], [((16,), np.float32)])

    def mat4_transform_vec3(self, m, v):
        return self._math(self._lib.tb_mat4_transform_vec3, [np.reshape(m, (-1, 16)), np.reshape(v, (-1, 3))], [((3,), np.float32)])

    def mat4_try_inverse(self, m):
        """-> (inverse (n, 16), some (n,) uint8); some == 0 is the reference's None."""
        return self._math(self._lib.tb_mat4_try_inverse, [np.reshape(m, (-1, 16))], [((16,), np.float32), ((), np.uint8)])

    # -- structural changes on the device (ecs.rs:93-124)
    def upload_spins(self, first, spins):
        a = np.ascontiguousarray(spins, dtype=np.float32).reshape(-1)
        self._check(self._lib.tb_upload_spins(self._h, int(first), a.shape[0], _ptr(a)))

    def insert(self, transform=None, sprite=None, velocity=None, parent=None):
        def one(v, dt):
            return None if v is None else np.ascontiguousarray(np.asarray(v, dtype=dt).reshape(1))
        t, s, v = one(transform, TRANSFORM_DTYPE), one(sprite, SPRITE_DTYPE), one(velocity, VELOCITY_DTYPE)
        p = None if parent is None else np.array([parent], np.uint32)
        idx = C.c_uint64()
        self._check(self._lib.tb_insert(self._h, _ptr(t), _ptr(s), _ptr(v), _ptr(p), C.byref(idx)))
        return idx.value

    def delete_by_ids(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        self._check(self._lib.tb_delete_by_ids(self._h, _ptr(a), a.shape[0]))

    def delete_by_query(self, required_mask):
        n = C.c_uint64()
        self._check(self._lib.tb_delete_by_query(self._h, int(required_mask), C.byref(n)))
        return n.value

    def add_component(self, component, index, value):
        dt = {TB_TRANSFORM: TRANSFORM_DTYPE, TB_SPRITE: SPRITE_DTYPE, TB_VELOCITY: VELOCITY_DTYPE, TB_PARENT: np.uint32,
              TB_SPIN: np.float32}[component]
        a = np.ascontiguousarray(np.asarray(value, dtype=dt).reshape(1))
        self._check(self._lib.tb_add_component(self._h, int(component), int(index), _ptr(a)))

    def remove_component(self, component, index):
        self._check(self._lib.tb_remove_component(self._h, int(component), int(index)))

    def entity_generations(self, first, n):
        out = np.zeros(int(n), np.uint32)
        self._check(self._lib.tb_entity_generations(self._h, int(first), int(n), _ptr(out)))
        return out

    def entities_alive(self, ids, generations):
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        g = np.ascontiguousarray(generations, dtype=np.uint32)
        out = np.zeros(a.shape[0], np.uint8)
        self._check(self._lib.tb_entities_alive(self._h, _ptr(a), _ptr(g), a.shape[0], _ptr(out)))
        return out

    # -- systems beyond the velocity system
    def system_spin(self, dt):
        self._check(self._lib.tb_system_spin(self._h, float(np.float32(dt))))

    def bundle_step(self, systems, dt):
        """SystemBundle::step for device systems; returns the stage every system was scheduled in."""
        a = np.ascontiguousarray(systems, dtype=np.int32)
        stages = np.zeros(a.shape[0], np.uint32)
        self._check(self._lib.tb_bundle_step(self._h, _ptr(a), a.shape[0], float(np.float32(dt)), _ptr(stages)))
        return stages

    def present_headless(self, width, height):
        """Rasterises the last frame's draw list: -> (ids (h, w) u32, rgba (h, w) u32, drawn, culled)."""
        ids = np.zeros((height, width), np.uint32)
        rgba = np.zeros((height, width), np.uint32)
        drawn, culled = C.c_uint64(), C.c_uint64()
        self._check(self._lib.tb_present_headless(self._h, int(width), int(height), _ptr(ids), _ptr(rgba), C.byref(drawn),
                                                  C.byref(culled)))
        return ids, rgba, drawn.value, culled.value

    def system_integrate(self, dt):
        self._check(self._lib.tb_system_integrate(self._h, float(np.float32(dt))))

    # -- frame
    def set_view_projection(self, view_proj):
        """Row-major 4x4 f32 folded into instances and vertices; None removes the camera."""
        if view_proj is None:
            self._check(self._lib.tb_frame_set_view_projection(self._h, None))
        else:
            m = np.ascontiguousarray(np.asarray(view_proj, np.float32).reshape(16))
            self._check(self._lib.tb_frame_set_view_projection(self._h, _ptr(m)))

    def force_path(self, path):
        self._check(self._lib.tb_frame_force_path(self._h, int(path)))

    def frame(self):
        out = FrameOut()
        self._check(self._lib.tb_frame_transform_batch(self._h, C.byref(out)))
        return out

    def download_instances(self, n, first=0, out=None):
        a = out if out is not None else np.zeros(int(n), dtype=INSTANCE_DTYPE)
        self._check(self._lib.tb_download_instances(self._h, int(first), int(n), _ptr(a)))
        return a

    def download_vertices(self, n, first=0, out=None):
        a = out if out is not None else np.zeros((int(n), 4), dtype=VERTEX_DTYPE)
        self._check(self._lib.tb_download_vertices(self._h, int(first), int(n), _ptr(a)))
        return a

    def download_batches(self, n, first=0):
        a = np.zeros(int(n), dtype=BATCH_DTYPE)
        self._check(self._lib.tb_download_batches(self._h, int(first), int(n), _ptr(a)))
        return a

    def download_order(self, n, first=0, out=None):
        a = out if out is not None else np.zeros(int(n), dtype=np.uint32)
        self._check(self._lib.tb_download_order(self._h, int(first), int(n), _ptr(a)))
        return a

    def download_world_matrices(self, first, n):
        a = np.zeros((int(n), 16), dtype=np.float32)
        self._check(self._lib.tb_download_world_matrices(self._h, int(first), int(n), _ptr(a)))
        return a

    # -- instrumentation
    def profile_enable(self, on=True):
        self._check(self._lib.tb_profile_enable(self._h, 1 if on else 0))

    def profile(self):
        """{kernel name: (total ms, expected bytes, launches)} since profile_enable()."""
        n = C.c_uint32()
        self._check(self._lib.tb_profile_count(self._h, C.byref(n)))
        rows = {}
        for i in range(n.value):
            name, ms, b, k = C.c_char_p(), C.c_double(), C.c_uint64(), C.c_uint32()
            self._check(self._lib.tb_profile_get(self._h, i, C.byref(name), C.byref(ms), C.byref(b), C.byref(k)))
            rows[name.value.decode()] = (ms.value, b.value, k.value)
        return rows

    def launch_count(self):
        n = C.c_uint64()
        self._check(self._lib.tb_launch_count(self._h, C.byref(n)))
        return n.value

    # -- sharding
    def shard_set(self, rank, nranks, global_first):
        self._check(self._lib.tb_shard_set(self._h, int(rank), int(nranks), int(global_first)))

    def shard_key_stats(self):
        """(OR of keys, OR of ~keys, renderable count) of this shard; reduce over ranks with OR, OR, +."""
        st = (C.c_uint64 * 3)()
        self._check(self._lib.tb_shard_key_stats(self._h, st))
        return [int(st[0]), int(st[1]), int(st[2])]

    def shard_plan(self, global_stats):
        st = (C.c_uint64 * 3)(*[int(v) for v in global_stats])
        nbins = C.c_uint64()
        self._check(self._lib.tb_shard_plan(self._h, st, C.byref(nbins)))
        return nbins.value

    def shard_key_histogram(self, d_hist):
        self._check(self._lib.tb_shard_key_histogram(self._h, C.c_void_p(d_hist)))

    def shard_place(self, d_all_hists):
        total = C.c_uint64()
        self._check(self._lib.tb_shard_place(self._h, C.c_void_p(d_all_hists), C.byref(total)))
        return total.value

    def shard_init(self, rank, nranks, shard_id, global_first, root=0):
        """Collective: joins the exchange group `shard_id` (bytes from shard_unique_id / shard_local_id)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(shard_id))
        self._check(self._lib.tb_shard_init(self._h, int(rank), int(nranks), buf, int(global_first), int(root)))

    def frame_sharded(self):
        """Collective: one frame of the sharded scene as one draw list on the presenting rank."""
        out = FrameOut()
        self._check(self._lib.tb_frame_sharded(self._h, C.byref(out)))
        return out

    def shard_finalize(self):
        self._check(self._lib.tb_shard_finalize(self._h))

    def vertices_from_instances(self, d_instances, first, count, d_vertices):
        self._check(self._lib.tb_vertices_from_instances(self._h, C.c_void_p(d_instances), int(first), int(count),
                                                         C.c_void_p(d_vertices)))

    def device_read(self, d_ptr, nbytes, dtype=np.uint8):
        out = np.zeros(int(nbytes) // np.dtype(dtype).itemsize, dtype=dtype)
        self._check(self._lib.tb_device_read(self._h, C.c_void_p(d_ptr), int(nbytes), _ptr(out)))
        return out

    def set_output(self, d_instances, d_vertices, d_order, capacity):
        self._check(self._lib.tb_frame_set_output(self._h, C.c_void_p(d_instances or 0), C.c_void_p(d_vertices or 0),
                                                  C.c_void_p(d_order or 0), int(capacity)))


def host_alloc(nbytes):
    """Page-locked host buffer (cudaMallocHost) as a uint8 numpy array; free with host_free."""
    lib = load()
    p = C.c_void_p()
    st = lib.tb_host_alloc(int(nbytes), C.byref(p))
    if st != TB_OK:
        raise TuberError(st, "tb_host_alloc failed")
    buf = (C.c_uint8 * int(nbytes)).from_address(p.value)
    return np.frombuffer(buf, dtype=np.uint8), p.value


def host_free(ptr):
    load().tb_host_free(C.c_void_p(ptr))


def device_alloc(nbytes, device=0):
    p = C.c_void_p()
    st = load().tb_device_alloc(int(device), int(nbytes), C.byref(p))
    if st != TB_OK:
        raise TuberError(st, "tb_device_alloc failed")
    return p.value


def device_free(ptr):
    load().tb_device_free(C.c_void_p(ptr))


def ipc_export(ptr):
    h = (C.c_uint8 * 64)()
    st = load().tb_ipc_export(C.c_void_p(ptr), h)
    if st != TB_OK:
        raise TuberError(st, "tb_ipc_export failed")
    return bytes(h)


def ipc_open(handle, device=0):
    p = C.c_void_p()
    h = (C.c_uint8 * 64)(*handle)
    st = load().tb_ipc_open(int(device), h, C.byref(p))
    if st != TB_OK:
        raise TuberError(st, "tb_ipc_open failed")
    return p.value


def ipc_close(ptr):
    load().tb_ipc_close(C.c_void_p(ptr))


def orthographic(left, right, bottom, top, near, far):
    """Matrix4::new_orthographic (matrix.rs:41-58): row-major 4x4 f32."""
    out = np.zeros(16, np.float32)
    if load().tb_orthographic(left, right, bottom, top, near, far, _ptr(out)) != 0:
        raise TuberError(1, "tb_orthographic")
    return out.reshape(4, 4)


def shard_unique_id():
    """A new NCCL exchange group id (128 bytes): rank 0 creates it and passes it to the other ranks."""
    buf = (C.c_uint8 * 128)()
    st = load().tb_shard_unique_id(buf)
    if st != TB_OK:
        raise TuberError(st, "tb_shard_unique_id (is libnccl.so.2 loadable?)")
    return bytes(buf)


def shard_local_id():
    """A new process-local exchange group id (several contexts of this process, one thread each)."""
    buf = (C.c_uint8 * 128)()
    st = load().tb_shard_local_id(buf)
    if st != TB_OK:
        raise TuberError(st, "tb_shard_local_id")
    return bytes(buf)


def system_access(system):
    """(reads, writes) column bit sets of a built-in device system (enum tb_column)."""
    r, w = C.c_uint32(), C.c_uint32()
    if load().tb_system_access(int(system), C.byref(r), C.byref(w)) != TB_OK:
        raise TuberError(TB_ERR_INVALID, "tb_system_access")
    return r.value, w.value
