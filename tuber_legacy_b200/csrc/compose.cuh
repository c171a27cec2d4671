// Device-side arithmetic of the hot path: Transform -> Matrix4 compose, parent * local,
// quad corners, sort key. Every parity-critical product and sum is an explicit
// __fmul_rn / __fadd_rn so nvcc can never contract them into FMAs: the reference is
// compiled by rustc, which rounds every operation individually
// (crates/tuber-math/src/matrix.rs:174-194, quaternion.rs:39-90).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "libm_sincosf.h"

namespace tb {

// Which build of glibc's sinf / cosf the HOST's libm uses (1: the FMA + AVX2 build, 0: the generic one); set per
// device by tb_ctx_create from the same CPU-feature test as glibc's own selector. See libm_sincosf.h.
__constant__ int c_libm_fma = 1;

// f32::sin / f32::cos of the reference (number_traits.rs:137-143 -> the platform libm), bit for bit.
// Not inlined: per-frame kernels only reach it for entities with pitch or roll (yaw's sin / cos are cached in the
// row), and one shared copy of the double-precision code keeps those kernels' hot paths small.
__device__ __noinline__ void ref_sincosf(float a, float* s, float* c) {
    if (c_libm_fma) tbm::sincosf_ref<true>(a, s, c);
    else tbm::sincosf_ref<false>(a, s, c);
}

// Device row of the (Transform, Sprite) table: four 16-byte chunks per entity.
//   A0 = {t.x, t.y, t.z, c.z}                A1 = {s.z, angle.x, angle.y, texture | layer << 16}
//   B0 = {sin(angle.z/2), c.x, c.y, s.x}     B1 = {s.y, size.w, size.h, cos(angle.z/2)}
// A0 is everything the velocity system writes; A0 + A1 everything the sort key of a planar entity
// needs. The yaw enters Transform::as_matrix4 only through sin / cos of its half (quaternion.rs:39-59), so the row
// keeps those two values — evaluated ONCE, when the Transform is uploaded (or an angular system changes it), with
// the host libm's sinf / cosf reproduced bit for bit (libm_sincosf.h) — and the per-frame kernels do no
// trigonometry at all for entities that rotate about z only (every 2-D sprite). angle.z itself lives in the cold
// `yaw` array (4 B/entity, read only by downloads and angular systems). Chunk A0 of entity i is t[i * st], A1 is a[i * st], B0 / B1 are b[i * sb] and b[i * sb + 1]
// (strides in float4 units). Two layouts:
//   split (start-up): three dense arrays T (16 B/entity), A (16 B), B (32 B), st = 1, sb = 2 — the
//     velocity system's write-back is a dense 16 B/entity stream and the key pass reads 32 B/entity
//   64-byte rows: t = R, a = R + 1, b = R + 2, st = sb = 4 — one DRAM atom per gathered entity
struct RowView {
    float4* t;
    float4* a;
    float4* b;
    uint32_t st, sb;
    float* yaw;  // angle.z per entity (cold)
};

// The four 16-byte quarters of one row, in registers.
struct Row {
    float4 a0, a1, b0, b1;
};

// Affine 3x4 world matrix, row-major rows 0..2 (row 3 is always 0,0,0,1).
struct Affine {
    float4 r0, r1, r2;
};

__device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addr(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subr(float a, float b) { return __fsub_rn(a, b); }

// The velocity system over `ticks` engine ticks of the same dt (the fixed-timestep loop of
// tuber-winit/src/lib.rs:117-141 runs 1-2 updates per render): translation += velocity * dt, once per
// tick, every product and sum rounded on its own — the value the ticks give one after the other.
__device__ __forceinline__ void integrate_ticks(float4& a0, float vx, float vy, float vz, float dt, uint32_t ticks) {
    const float dx = mulr(vx, dt), dy = mulr(vy, dt), dz = mulr(vz, dt);
    for (uint32_t k = 0; k < ticks; ++k) {
        a0.x = addr(a0.x, dx);
        a0.y = addr(a0.y, dy);
        a0.z = addr(a0.z, dz);
    }
}

// Quaternion::from_euler (quaternion.rs:39-59) followed by rotation_matrix (:63-90),
// un-normalised, returning the 3x3 block R[j][i].
// Fast path: angle.x == 0 && angle.y == 0 (every 2-D sprite) needs one sincos instead of
// three and is value-identical, because x1, x0 and +0 are exact:
//   w = 1*1*cy + 0*0*sy = cy,  x = 0,  y = 0,  z = 1*1*sy - 0*0*cy = sy.
// `sy_`, `cy_`: sin / cos of yaw / 2 as the row caches them.
__device__ __forceinline__ void rotation3x3_yaw(float ax, float ay, float sy_, float cy_, float R[9]) {
    float w, x, y, z;
    if (ax == 0.0f && ay == 0.0f) {
        // 2-D: R = [[1-zz2, -wz2, 0], [wz2, 1-zz2, 0], [0, 0, 1]]
        float z2 = addr(sy_, sy_);
        float w2 = addr(cy_, cy_);
        float zz2 = mulr(z2, sy_);
        float wz2 = mulr(w2, sy_);
        float d = subr(1.0f, zz2);
        R[0] = d;     R[1] = -wz2;  R[2] = 0.0f;
        R[3] = wz2;   R[4] = d;     R[5] = 0.0f;
        R[6] = 0.0f;  R[7] = 0.0f;  R[8] = 1.0f;
        return;
    }
    float sp, cp, sr, cr;
    ref_sincosf(mulr(ay, 0.5f), &sp, &cp);
    ref_sincosf(mulr(ax, 0.5f), &sr, &cr);
    w = addr(mulr(mulr(cr, cp), cy_), mulr(mulr(sr, sp), sy_));
    x = subr(mulr(mulr(sr, cp), cy_), mulr(mulr(cr, sp), sy_));
    y = addr(mulr(mulr(cr, sp), cy_), mulr(mulr(sr, cp), sy_));
    z = subr(mulr(mulr(cr, cp), sy_), mulr(mulr(sr, sp), cy_));
    float x2 = addr(x, x), y2 = addr(y, y), z2 = addr(z, z), w2 = addr(w, w);
    float xx2 = mulr(x2, x), xy2 = mulr(x2, y), xz2 = mulr(x2, z);
    float yy2 = mulr(y2, y), yz2 = mulr(y2, z), zz2 = mulr(z2, z);
    float wx2 = mulr(w2, x), wy2 = mulr(w2, y), wz2 = mulr(w2, z);
    R[0] = subr(subr(1.0f, yy2), zz2);  R[1] = subr(xy2, wz2);              R[2] = addr(xz2, wy2);
    R[3] = addr(xy2, wz2);              R[4] = subr(subr(1.0f, xx2), zz2);  R[5] = subr(yz2, wx2);
    R[6] = subr(xz2, wy2);              R[7] = addr(yz2, wx2);              R[8] = subr(subr(1.0f, xx2), yy2);
}
__device__ __forceinline__ void rotation3x3(float ax, float ay, float az, float R[9]) {
    float sy_, cy_;
    ref_sincosf(mulr(az, 0.5f), &sy_, &cy_);
    rotation3x3_yaw(ax, ay, sy_, cy_, R);
}
// sin / cos of yaw / 2 for the row's B0.x / B1.w slots (upload conversion, angular systems)
__device__ __forceinline__ float2 yaw_sincos(const float yaw) {
    float2 sc;
    ref_sincosf(mulr(yaw, 0.5f), &sc.x, &sc.y);
    return sc;
}

// One row j of Transform::as_matrix4 (transform.rs:28-38) in closed form
// (equivalence with the 4-multiply chain is proven on the CPU in
// tests/test_oracle_math.py):
//   r_ji   = s_j * R[j][i]
//   tr_j   = s_j*c_j + s_j*t_j
//   M[j][3] = ((r_j0*(-c_x) + r_j1*(-c_y)) + r_j2*(-c_z)) + tr_j
__device__ __forceinline__ float4 compose_row(const float* Rj, float sj, float cj, float tj, float cx, float cy,
                                              float cz) {
    float r0 = mulr(sj, Rj[0]);
    float r1 = mulr(sj, Rj[1]);
    float r2 = mulr(sj, Rj[2]);
    float trj = addr(mulr(sj, cj), mulr(sj, tj));
    float acc = addr(mulr(r0, -cx), mulr(r1, -cy));
    acc = addr(acc, mulr(r2, -cz));
    acc = addr(acc, trj);
    return make_float4(r0, r1, r2, acc);
}

// Full local matrix of a row.
__device__ __forceinline__ Affine compose_local(const Row& r) {
    const float tx = r.a0.x, ty = r.a0.y, tz = r.a0.z, cz = r.a0.w;
    const float sz = r.a1.x, ax = r.a1.y, ay = r.a1.z;
    const float cx = r.b0.y, cy = r.b0.z, sx = r.b0.w;
    const float sy = r.b1.x;
    float R[9];
    rotation3x3_yaw(ax, ay, r.b0.x, r.b1.w, R);
    Affine m;
    m.r0 = compose_row(R + 0, sx, cx, tx, cx, cy, cz);
    m.r1 = compose_row(R + 3, sy, cy, ty, cx, cy, cz);
    m.r2 = compose_row(R + 6, sz, cz, tz, cx, cy, cz);
    return m;
}

// True when the row's rotation is about z only (every 2-D sprite): then M[2][3] needs no
// trigonometry and nothing from chunks B0 B1.
__device__ __forceinline__ bool is_planar(const float4 a1) { return a1.y == 0.0f && a1.z == 0.0f; }

// M[2][3] (the world z of a flat entity) from chunks A0 A1 alone, valid when is_planar(a1):
// R[2] = (0, 0, 1), so with finite c.x, c.y the row-2 sum is
//   ((+-0 + +-0) + s.z*(-c.z)) + (s.z*c.z + s.z*t.z).
__device__ __forceinline__ float planar_z(const float4 a0, const float4 a1) {
    const float tz = a0.z, cz = a0.w, sz = a1.x;
    float trj = addr(mulr(sz, cz), mulr(sz, tz));
    float acc = addr(0.0f, mulr(sz, -cz));
    return addr(acc, trj);
}

// M[2][3] of any row (general 3-D rotation needs chunks B0 B1 too).
__device__ __forceinline__ float local_z(const Row& r) {
    const float tz = r.a0.z, cz = r.a0.w, sz = r.a1.x, ax = r.a1.y, ay = r.a1.z;
    const float cx = r.b0.y, cy = r.b0.z;
    float R[9];
    rotation3x3_yaw(ax, ay, r.b0.x, r.b1.w, R);
    return compose_row(R + 6, sz, cz, tz, cx, cy, cz).w;
}

// Matrix4::mul (matrix.rs:174-194) for world = parent * local with both bottom rows
// (0,0,0,1): out[j][i] = ((p[j][0]*l[0][i] + p[j][1]*l[1][i]) + p[j][2]*l[2][i]) + p[j][3]*l[3][i].
// l[3][i] is 0 for i<3 and 1 for i==3, so the last term is +0 (value-neutral) or + p[j][3].
__device__ __forceinline__ float4 mul_row(const float4 p, const Affine& l) {
    float4 o;
    o.x = addr(addr(mulr(p.x, l.r0.x), mulr(p.y, l.r1.x)), mulr(p.z, l.r2.x));
    o.y = addr(addr(mulr(p.x, l.r0.y), mulr(p.y, l.r1.y)), mulr(p.z, l.r2.y));
    o.z = addr(addr(mulr(p.x, l.r0.z), mulr(p.y, l.r1.z)), mulr(p.z, l.r2.z));
    o.w = addr(addr(addr(mulr(p.x, l.r0.w), mulr(p.y, l.r1.w)), mulr(p.z, l.r2.w)), p.w);
    return o;
}
__device__ __forceinline__ Affine mul_affine(const Affine& p, const Affine& l) {
    Affine o;
    o.r0 = mul_row(p.r0, l);
    o.r1 = mul_row(p.r1, l);
    o.r2 = mul_row(p.r2, l);
    return o;
}

// [SPEC] sort key: layer:16 | texture:16 | orderable(world z + 0.0f):32.
__device__ __forceinline__ uint64_t sort_key(uint32_t layer, uint32_t texture, float world_z) {
    float z = addr(world_z, 0.0f);  // folds -0 into +0
    uint32_t u = __float_as_uint(z);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
    return (uint64_t(layer & 0xFFFFu) << 48) | (uint64_t(texture & 0xFFFFu) << 32) | uint64_t(u);
}

// Matrix4::transform_vec3 (matrix.rs:160-171) of corner (x, y, 0, 1):
//   ((m0*x + m1*y) + m2*0) + m3*1  — the m2*0 term is a value-neutral +-0.
__device__ __forceinline__ float corner_coord(float m0, float m1, float m3, float x, float y) {
    return addr(addr(mulr(m0, x), mulr(m1, y)), m3);
}

}  // namespace tb
