// Device-side arithmetic of the hot path: Transform -> Matrix4 compose, parent * local,
// quad corners, sort key. Every parity-critical product and sum is an explicit
// __fmul_rn / __fadd_rn so nvcc can never contract them into FMAs: the reference is
// compiled by rustc, which rounds every operation individually
// (crates/tuber-math/src/matrix.rs:174-194, quaternion.rs:39-90).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

// One packed 64-byte row of the (Transform, Sprite) table — exactly two 32-byte DRAM
// sectors, so a random gather of one entity moves no wasted bytes.
//   a = {t.x, t.y, t.z, angle.x}   b = {angle.y, angle.z, c.x, c.y}
//   c = {c.z, s.x, s.y, s.z}       d = {size.w, size.h, texture, layer}
struct __align__(16) Row {
    float4 a, b, c, d;
};
static_assert(sizeof(Row) == 64, "Row is 64 bytes");

// Affine 3x4 world matrix, row-major rows 0..2 (row 3 is always 0,0,0,1).
struct Affine {
    float4 r0, r1, r2;
};

__device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addr(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subr(float a, float b) { return __fsub_rn(a, b); }

// Quaternion::from_euler (quaternion.rs:39-59) followed by rotation_matrix (:63-90),
// un-normalised, returning the 3x3 block R[j][i].
// Fast path: angle.x == 0 && angle.y == 0 (every 2-D sprite) needs one sincos instead of
// three and is value-identical, because x1, x0 and +0 are exact:
//   w = 1*1*cy + 0*0*sy = cy,  x = 0,  y = 0,  z = 1*1*sy - 0*0*cy = sy.
__device__ __forceinline__ void rotation3x3(float ax, float ay, float az, float R[9]) {
    float w, x, y, z;
    float sy_, cy_;
    sincosf(mulr(az, 0.5f), &sy_, &cy_);
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
    sincosf(mulr(ay, 0.5f), &sp, &cp);
    sincosf(mulr(ax, 0.5f), &sr, &cr);
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
__device__ __forceinline__ Affine compose_local(const float4 a, const float4 b, const float4 c) {
    const float tx = a.x, ty = a.y, tz = a.z, ax = a.w;
    const float ay = b.x, az = b.y, cx = b.z, cy = b.w;
    const float cz = c.x, sx = c.y, sy = c.z, sz = c.w;
    float R[9];
    rotation3x3(ax, ay, az, R);
    Affine m;
    m.r0 = compose_row(R + 0, sx, cx, tx, cx, cy, cz);
    m.r1 = compose_row(R + 3, sy, cy, ty, cx, cy, cz);
    m.r2 = compose_row(R + 6, sz, cz, tz, cx, cy, cz);
    return m;
}

// Only M[2][3] (the world z of a flat entity) — all the key pass needs. For a 2-D
// transform this is trig-free: R[2] = (0, 0, 1).
__device__ __forceinline__ float compose_local_z(const float4 a, const float4 b, const float4 c) {
    const float tz = a.z, ax = a.w;
    const float ay = b.x, az = b.y, cx = b.z, cy = b.w;
    const float cz = c.x, sz = c.w;
    if (ax == 0.0f && ay == 0.0f) {
        const float R2[3] = {0.0f, 0.0f, 1.0f};
        return compose_row(R2, sz, cz, tz, cx, cy, cz).w;
    }
    float R[9];
    rotation3x3(ax, ay, az, R);
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

// Bit-gather (pext) of the varying key bits, as up to 8 contiguous runs. Built on the host
// (or by make_plan_kernel) from the frame's OR/AND masks.
struct PextPlan {
    uint32_t nruns;
    uint32_t nbits;       // total extracted bits
    uint8_t src_shift[16];
    uint8_t width[16];
    uint8_t dst_shift[16];
};

__device__ __forceinline__ uint64_t pext_runs(uint64_t key, const PextPlan& p) {
    uint64_t out = 0;
#pragma unroll 4
    for (uint32_t r = 0; r < p.nruns; ++r) {
        uint64_t m = (p.width[r] >= 64) ? ~0ull : ((1ull << p.width[r]) - 1ull);
        out |= ((key >> p.src_shift[r]) & m) << p.dst_shift[r];
    }
    return out;
}

__host__ __device__ inline bool build_pext_plan(uint64_t varying, PextPlan* p) {
    p->nruns = 0;
    p->nbits = 0;
    uint32_t bit = 0;
    while (bit < 64) {
        if (!((varying >> bit) & 1ull)) { ++bit; continue; }
        uint32_t start = bit;
        while (bit < 64 && ((varying >> bit) & 1ull)) ++bit;
        if (p->nruns >= 16) return false;
        p->src_shift[p->nruns] = (uint8_t)start;
        p->width[p->nruns] = (uint8_t)(bit - start);
        p->dst_shift[p->nruns] = (uint8_t)p->nbits;
        p->nbits += bit - start;
        p->nruns += 1;
    }
    return true;
}

}  // namespace tb
