"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

ctypes binding of oracle/_build/liboracle.so: the C++ CPU restatement of the reference's
tuber-math / tuber-core / tuber-ecs code on the hot path (ref_math.hpp, ref_ecs.hpp) plus the
builder-SPEC frame built from those primitives (ref_frame.hpp). The reference itself cannot be
compiled here (Rust, no cargo/rustc in the image), so there is no oracle/_ref.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
SOURCES = ["oracle_capi.cpp", "ref_math.hpp", "ref_ecs.hpp", "ref_frame.hpp", "Makefile"]

INSTANCE_DTYPE = np.dtype([("model", "<f4", 16), ("size", "<f4", 2), ("texture", "<u4"), ("layer", "<u4")])
VERTEX_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("uv", "<u4")])
BATCH_DTYPE = np.dtype([("layer", "<u4"), ("texture", "<u4"), ("first_instance", "<u4"), ("count", "<u4")])

_lib = None


def build(force=False):
    stale = force or not os.path.exists(LIB)
    if not stale:
        t = os.path.getmtime(LIB)
        stale = any(os.path.getmtime(os.path.join(HERE, s)) > t for s in SOURCES)
    if stale:
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, u64, u32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_double
        L.orc_scene_create.restype = vp
        L.orc_scene_create.argtypes = [u64, vp, vp, vp, vp, vp, vp, vp, u64, vp, vp]
        L.orc_scene_step_systems.restype = C.c_int
        L.orc_scene_step_systems.argtypes = [vp, dbl, vp, u32]
        L.orc_ecs_free.argtypes = [vp]
        L.orc_ecs_new.restype = vp
        L.orc_ecs_new.argtypes = [u64]
        L.orc_ecs_entity_count.restype = u64
        L.orc_ecs_entity_count.argtypes = [vp]
        L.orc_ecs_last_error.restype = C.c_char_p
        L.orc_ecs_last_error.argtypes = [vp]
        L.orc_scene_step.restype = dbl
        L.orc_scene_step.argtypes = [vp, dbl]
        L.orc_scene_read_transforms.argtypes = [vp, vp]
        L.orc_frame_run.restype = vp
        L.orc_frame_run.argtypes = [vp]
        L.orc_frame_run_camera.restype = vp
        L.orc_frame_run_camera.argtypes = [vp, vp]
        L.orc_orthographic.argtypes = [C.c_float] * 6 + [vp]
        L.orc_frame_free.argtypes = [vp]
        L.orc_frame_bounds.argtypes = [vp, vp, vp, vp]
        L.orc_frame_seconds.restype = dbl
        L.orc_frame_seconds.argtypes = [vp]
        for f in ("orc_frame_num_instances", "orc_frame_num_batches", "orc_frame_num_entities"):
            getattr(L, f).restype = u64
            getattr(L, f).argtypes = [vp]
        L.orc_frame_read_world.argtypes = [vp, vp, vp]
        for f in ("orc_frame_read_order", "orc_frame_read_keys", "orc_frame_read_instances",
                  "orc_frame_read_vertices", "orc_frame_read_batches"):
            getattr(L, f).argtypes = [vp, vp]
        L.orc_sort_key.restype = u64
        L.orc_sort_key.argtypes = [u32, u32, C.c_float]
        L.orc_ecs_insert.restype = C.c_int64
        L.orc_ecs_insert.argtypes = [vp, u32, vp, vp, vp]
        L.orc_ecs_query.restype = u64
        L.orc_ecs_query.argtypes = [vp, u32, vp, vp, vp, vp, vp, u64]
        L.orc_ecs_query_by_ids.restype = u64
        L.orc_ecs_query_by_ids.argtypes = [vp, u32, vp, vp, vp, vp, u64, vp, u64]
        L.orc_ecs_query_one.restype = C.c_int64
        L.orc_ecs_query_one.argtypes = [vp, u32, vp, vp, vp, vp]
        L.orc_ecs_get.restype = C.c_int64
        L.orc_ecs_get.argtypes = [vp, u64, u64, vp, u64]
        L.orc_ecs_set.restype = C.c_int
        L.orc_ecs_set.argtypes = [vp, u64, u64, vp, u64]
        L.orc_ecs_delete_by_ids.argtypes = [vp, vp, u64]
        L.orc_ecs_delete_by_query.argtypes = [vp, u32, vp, vp, vp]
        L.orc_ecs_add_component.argtypes = [vp, u64, vp, u64, u64]
        L.orc_ecs_remove_component.argtypes = [vp, u64, u64]
        L.orc_ecs_insert_shared_resource.argtypes = [vp, u64, vp, u64]
        L.orc_ecs_shared_resource.restype = C.c_int64
        L.orc_ecs_shared_resource.argtypes = [vp, u64, vp, u64]
        L.orc_ecs_double_borrow_conflicts.restype = C.c_int
        L.orc_ecs_double_borrow_conflicts.argtypes = [vp, u64, u64]
        L.orc_bundle_new.restype = vp
        L.orc_bundle_free.argtypes = [vp]
        L.orc_bundle_len.restype = u64
        L.orc_bundle_len.argtypes = [vp]
        L.orc_bundle_step.restype = C.c_int
        L.orc_bundle_step.argtypes = [vp, vp, vp]
        L.orc_as_matrix4.argtypes = [vp, u64, vp]
        L.orc_as_matrix4_closed.argtypes = [vp, u64, vp]
        L.orc_u64_set_bit.restype = u64
        L.orc_u64_set_bit.argtypes = [u64, u64]
        L.orc_u64_unset_bit.restype = u64
        L.orc_u64_unset_bit.argtypes = [u64, u64]
        L.orc_u64_bit.restype = C.c_int
        L.orc_u64_bit.argtypes = [u64, u64]
        for f in ("orc_words_set_bit", "orc_words_unset_bit", "orc_words_bit"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [vp, u64, u64]
        for f in ("orc_batch_quat_from_euler", "orc_batch_quat_from_axis_angle", "orc_batch_quat_rotation_matrix",
                  "orc_batch_quat_normalize"):
            getattr(L, f).argtypes = [vp, u64, vp]
        for f in ("orc_batch_quat_mul", "orc_batch_mat4_mul", "orc_batch_mat4_transform_vec3"):
            getattr(L, f).argtypes = [vp, vp, u64, vp]
        L.orc_batch_mat4_try_inverse.argtypes = [vp, u64, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _f32(a, cols):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    if a.dtype.names is not None or a.dtype != np.float32:
        a = a.view(np.float32) if a.dtype.itemsize == cols * 4 or a.dtype.names is not None else a.astype(np.float32)
    return np.ascontiguousarray(a.reshape(-1, cols))


def _u32(a, cols):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    if a.dtype.names is not None:
        a = a.view(np.uint32)
    return np.ascontiguousarray(a.reshape(-1, cols).view(np.uint32))


class Frame:
    """Result of ref::render_scene over a restated-reference Ecs."""

    def __init__(self, h, scene_handle=None, view_proj=None):
        L = lib()
        n = L.orc_frame_num_entities(h)
        m = L.orc_frame_num_instances(h)
        b = L.orc_frame_num_batches(h)
        self.seconds = L.orc_frame_seconds(h)
        self.world = np.zeros((n, 16), np.float32)
        self.has_world = np.zeros(n, np.uint8)
        L.orc_frame_read_world(h, _p(self.world), _p(self.has_world))
        self.order = np.zeros(m, np.uint32)
        self.keys = np.zeros(m, np.uint64)
        self.instances = np.zeros(m, INSTANCE_DTYPE)
        self.vertices = np.zeros((m, 4), VERTEX_DTYPE)
        self.batches = np.zeros(b, BATCH_DTYPE)
        if m:
            L.orc_frame_read_order(h, _p(self.order))
            L.orc_frame_read_keys(h, _p(self.keys))
            L.orc_frame_read_instances(h, _p(self.instances))
            L.orc_frame_read_vertices(h, _p(self.vertices))
        if b:
            L.orc_frame_read_batches(h, _p(self.batches))
        # forward bounds of the model matrices (row-major, per instance): the denominator of the parity metric
        self.bounds = np.zeros((m, 16), np.float64)
        if m and scene_handle:
            L.orc_frame_bounds(scene_handle, h, _p(view_proj), _p(self.bounds))
        L.orc_frame_free(h)


class Scene:
    """A restated-reference Ecs filled from component arrays (NULL array == no such store)."""

    def __init__(self, n, transforms=None, sprites=None, velocities=None, parents=None,
                 mask_transform=None, mask_sprite=None, mask_velocity=None, bitset_words=0, spins=None, mask_spin=None):
        L = lib()
        self.n = int(n)
        t = _f32(transforms, 12)
        s = _u32(sprites, 4)
        v = _f32(velocities, 3)
        p = None if parents is None else np.ascontiguousarray(parents, dtype=np.uint32)
        masks = [None if m is None else np.ascontiguousarray(m, dtype=np.uint64)
                 for m in (mask_transform, mask_sprite, mask_velocity)]
        w = None if spins is None else np.ascontiguousarray(spins, dtype=np.float32)
        mw = None if mask_spin is None else np.ascontiguousarray(mask_spin, dtype=np.uint64)
        self._h = L.orc_scene_create(self.n, _p(t), _p(masks[0]), _p(s), _p(masks[1]), _p(v), _p(masks[2]), _p(p),
                                     int(bitset_words), _p(w), _p(mw))
        if not self._h:
            raise OverflowError("entity index beyond the presence bitset (the reference panics)")

    def close(self):
        if self._h:
            lib().orc_ecs_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, dt):
        """One engine tick running the velocity system; returns elapsed seconds."""
        r = lib().orc_scene_step(self._h, float(dt))
        if r < 0:
            raise RuntimeError(lib().orc_ecs_last_error(self._h).decode())
        return r

    def step_systems(self, dt, systems):
        """One tick running the built-in systems (0: velocity, 1: spin) in this order, like SystemBundle::step."""
        a = np.ascontiguousarray(systems, dtype=np.int32)
        if lib().orc_scene_step_systems(self._h, float(dt), _p(a), len(a)) != 0:
            raise RuntimeError(lib().orc_ecs_last_error(self._h).decode())

    def transforms(self):
        out = np.zeros((self.n, 12), np.float32)
        lib().orc_scene_read_transforms(self._h, _p(out))
        return out

    def frame(self, view_proj=None, bounds=True):
        """render_scene; `view_proj` (4x4 row-major f32) folds a camera into instances and vertices."""
        if view_proj is None:
            return Frame(lib().orc_frame_run(self._h), self._h if bounds else None)
        m = np.ascontiguousarray(np.asarray(view_proj, np.float32).reshape(16))
        return Frame(lib().orc_frame_run_camera(self._h, _p(m)), self._h if bounds else None, m)


def orthographic(left, right, bottom, top, near, far):
    """Matrix4::new_orthographic (matrix.rs:41-58), row-major 4x4 f32."""
    out = np.zeros(16, np.float32)
    lib().orc_orthographic(left, right, bottom, top, near, far, _p(out))
    return out.reshape(4, 4)


def as_matrix4(transforms):
    t = _f32(transforms, 12)
    out = np.zeros((t.shape[0], 16), np.float32)
    lib().orc_as_matrix4(_p(t), t.shape[0], _p(out))
    return out


def as_matrix4_closed(transforms):
    t = _f32(transforms, 12)
    out = np.zeros((t.shape[0], 16), np.float32)
    lib().orc_as_matrix4_closed(_p(t), t.shape[0], _p(out))
    return out


def sort_key(layer, texture, z):
    return lib().orc_sort_key(int(layer), int(texture), float(np.float32(z)))


# ---- batched tuber-math (matrix.rs / quaternion.rs restatements), for the device-vs-oracle tests ----

def _batch(fn, ins, cols_out, dtype=np.float32):
    ins = [np.ascontiguousarray(a, dtype=np.float32) for a in ins]
    n = ins[0].shape[0]
    out = np.zeros((n, cols_out), dtype) if cols_out else np.zeros(n, dtype)
    getattr(lib(), fn)(*[_p(a) for a in ins], n, _p(out))
    return out


def quat_from_euler(e):
    return _batch("orc_batch_quat_from_euler", [np.reshape(e, (-1, 3))], 4)


def quat_from_axis_angle(aa):
    return _batch("orc_batch_quat_from_axis_angle", [np.reshape(aa, (-1, 4))], 4)


def quat_rotation_matrix(q):
    return _batch("orc_batch_quat_rotation_matrix", [np.reshape(q, (-1, 4))], 16)


def quat_mul(a, b):
    return _batch("orc_batch_quat_mul", [np.reshape(a, (-1, 4)), np.reshape(b, (-1, 4))], 4)


def quat_normalize(q):
    return _batch("orc_batch_quat_normalize", [np.reshape(q, (-1, 4))], 4)


def mat4_mul(a, b):
    return _batch("orc_batch_mat4_mul", [np.reshape(a, (-1, 16)), np.reshape(b, (-1, 16))], 16)


def mat4_transform_vec3(m, v):
    return _batch("orc_batch_mat4_transform_vec3", [np.reshape(m, (-1, 16)), np.reshape(v, (-1, 3))], 3)


def mat4_try_inverse(m):
    m = np.ascontiguousarray(np.reshape(m, (-1, 16)), dtype=np.float32)
    out = np.zeros_like(m)
    some = np.zeros(m.shape[0], np.uint8)
    lib().orc_batch_mat4_try_inverse(_p(m), m.shape[0], _p(out), _p(some))
    return out, some


# ---- CPU rasterisation of a draw list (checker of the headless presenter, tb_present_headless) ----

def rasterize(vertices, width, height):
    """Painter's-rule rasterisation of quads (n, 4) VERTEX_DTYPE in draw order, clip space, pixel centres, four edge
    functions in float32 with every operation rounded on its own (numpy float32 arithmetic): ids[j, i] = 1 + position of
    the last quad covering pixel (i, j), 0 for none. Row 0 is clip y = -1."""
    f = np.float32
    ids = np.zeros((height, width), np.uint32)
    sx, sy = f(2.0) / f(width), f(2.0) / f(height)
    cx = (np.arange(width, dtype=np.float32) + f(0.5)) * sx - f(1.0)
    cy = (np.arange(height, dtype=np.float32) + f(0.5)) * sy - f(1.0)
    vx = np.asarray(vertices["x"], np.float32)
    vy = np.asarray(vertices["y"], np.float32)
    for q in range(vx.shape[0]):
        x, y = vx[q], vy[q]
        if not (x.max() >= -1 and x.min() <= 1 and y.max() >= -1 and y.min() <= 1):
            continue
        i0 = max(0, int(np.floor((float(x.min()) + 1) * 0.5 * width)) - 2)
        i1 = min(width - 1, int(np.floor((float(x.max()) + 1) * 0.5 * width)) + 2)
        j0 = max(0, int(np.floor((float(y.min()) + 1) * 0.5 * height)) - 2)
        j1 = min(height - 1, int(np.floor((float(y.max()) + 1) * 0.5 * height)) + 2)
        if i1 < i0 or j1 < j0:
            continue
        px = cx[i0:i1 + 1][None, :]
        py = cy[j0:j1 + 1][:, None]
        pos = np.ones((j1 - j0 + 1, i1 - i0 + 1), bool)
        neg = pos.copy()
        for k in range(4):
            ax, ay, bx, by = x[k], y[k], x[(k + 1) % 4], y[(k + 1) % 4]
            e = (bx - ax) * (py - ay) - (by - ay) * (px - ax)   # float32 throughout, no fused operations
            pos &= e >= 0
            neg &= e <= 0
        inside = pos | neg
        view = ids[j0:j1 + 1, i0:i1 + 1]
        view[inside] = q + 1
    return ids
