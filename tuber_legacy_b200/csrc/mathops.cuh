// The rest of tuber-math on the device, batched: quaternion construction / product /
// normalisation (crates/tuber-math/src/quaternion.rs:28-158), dense Matrix4 product, inverse and point transform
// (crates/tuber-math/src/matrix.rs:98-149,160-194). Not on the per-frame path (picking, screen -> world, 3-D
// rotation composition), so one thread per element and plain coalescable loads. Same discipline as compose.cuh:
// every product, sum, quotient and root is individually rounded in the reference's own order, so the results are
// the reference's bit for bit (sin / cos through libm_sincosf.h).
#pragma once
#include "common.cuh"

namespace tb {

__device__ __forceinline__ float divr(float a, float b) { return __fdiv_rn(a, b); }

// Quaternion::from_euler (quaternion.rs:39-59): angles = (roll, pitch, yaw); out = (w, x, y, z).
__global__ void __launch_bounds__(kThreads) quat_from_euler_kernel(const float* __restrict__ euler, uint32_t n,
                                                                    float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float roll = euler[3 * i], pitch = euler[3 * i + 1], yaw = euler[3 * i + 2];
    float sy, cy, sp, cp, sr, cr;
    ref_sincosf(mulr(yaw, 0.5f), &sy, &cy);
    ref_sincosf(mulr(pitch, 0.5f), &sp, &cp);
    ref_sincosf(mulr(roll, 0.5f), &sr, &cr);
    float4 q;
    q.x = addr(mulr(mulr(cr, cp), cy), mulr(mulr(sr, sp), sy));  // w
    q.y = subr(mulr(mulr(sr, cp), cy), mulr(mulr(cr, sp), sy));  // x
    q.z = addr(mulr(mulr(cr, sp), cy), mulr(mulr(sr, cp), sy));  // y
    q.w = subr(mulr(mulr(cr, cp), sy), mulr(mulr(sr, sp), cy));  // z
    out[i] = q;
}

// Quaternion::from_axis_angle (quaternion.rs:28-37): in = (axis.x, axis.y, axis.z, angle); out = (w, x, y, z).
__global__ void __launch_bounds__(kThreads) quat_from_axis_angle_kernel(const float4* __restrict__ in, uint32_t n,
                                                                         float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = in[i];
    float s, c;
    ref_sincosf(mulr(a.w, 0.5f), &s, &c);
    out[i] = make_float4(c, mulr(a.x, s), mulr(a.y, s), mulr(a.z, s));
}

// Quaternion::rotation_matrix (quaternion.rs:63-90): (w, x, y, z), NOT normalised first -> row-major 4x4.
__global__ void __launch_bounds__(kThreads) quat_rotation_matrix_kernel(const float4* __restrict__ q, uint32_t n,
                                                                         float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = q[i];
    const float w = v.x, x = v.y, y = v.z, z = v.w;
    const float x2 = addr(x, x), y2 = addr(y, y), z2 = addr(z, z), w2 = addr(w, w);
    const float xx2 = mulr(x2, x), xy2 = mulr(x2, y), xz2 = mulr(x2, z);
    const float yy2 = mulr(y2, y), yz2 = mulr(y2, z), zz2 = mulr(z2, z);
    const float wx2 = mulr(w2, x), wy2 = mulr(w2, y), wz2 = mulr(w2, z);
    float4* o = out + (size_t)i * 4;
    o[0] = make_float4(subr(subr(1.0f, yy2), zz2), subr(xy2, wz2), addr(xz2, wy2), 0.0f);
    o[1] = make_float4(addr(xy2, wz2), subr(subr(1.0f, xx2), zz2), subr(yz2, wx2), 0.0f);
    o[2] = make_float4(subr(xz2, wy2), addr(yz2, wx2), subr(subr(1.0f, xx2), yy2), 0.0f);
    o[3] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
}

// impl Mul for Quaternion (quaternion.rs:132-158), terms summed in the reference's order.
__global__ void __launch_bounds__(kThreads) quat_mul_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                             uint32_t n, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = a[i], r = b[i];
    const float w1 = p.x, x1 = p.y, y1 = p.z, z1 = p.w;
    const float w2 = r.x, x2 = r.y, y2 = r.z, z2 = r.w;
    float4 o;
    o.x = subr(subr(subr(mulr(w1, w2), mulr(x1, x2)), mulr(y1, y2)), mulr(z1, z2));
    o.y = addr(subr(addr(mulr(x1, w2), mulr(y1, z2)), mulr(z1, y2)), mulr(w1, x2));
    o.z = subr(addr(addr(mulr(y1, w2), mulr(z1, x2)), mulr(w1, y2)), mulr(x1, z2));
    o.w = subr(addr(addr(mulr(z1, w2), mulr(w1, z2)), mulr(x1, y2)), mulr(y1, x2));
    out[i] = o;
}

// Quaternion::normalized (quaternion.rs:92-120): norm = sqrt(ww + xx + yy + zz), every part divided by it.
__global__ void __launch_bounds__(kThreads) quat_normalize_kernel(const float4* __restrict__ q, uint32_t n,
                                                                   float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = q[i];
    const float w = v.x, x = v.y, y = v.z, z = v.w;
    const float xx = mulr(x, x), yy = mulr(y, y), zz = mulr(z, z), ww = mulr(w, w);
    const float norm = __fsqrt_rn(addr(addr(addr(ww, xx), yy), zz));
    out[i] = make_float4(divr(w, norm), divr(x, norm), divr(y, norm), divr(z, norm));
}

// impl Mul for Matrix4 (matrix.rs:174-194): dense row-major 4x4, each entry summed left to right.
__global__ void __launch_bounds__(kThreads) mat4_mul_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                             uint32_t n, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* A = a + (size_t)i * 4;
    const float4 b0 = b[(size_t)i * 4], b1 = b[(size_t)i * 4 + 1], b2 = b[(size_t)i * 4 + 2], b3 = b[(size_t)i * 4 + 3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 r = A[j];
        float4 o;
        o.x = addr(addr(addr(mulr(r.x, b0.x), mulr(r.y, b1.x)), mulr(r.z, b2.x)), mulr(r.w, b3.x));
        o.y = addr(addr(addr(mulr(r.x, b0.y), mulr(r.y, b1.y)), mulr(r.z, b2.y)), mulr(r.w, b3.y));
        o.z = addr(addr(addr(mulr(r.x, b0.z), mulr(r.y, b1.z)), mulr(r.z, b2.z)), mulr(r.w, b3.z));
        o.w = addr(addr(addr(mulr(r.x, b0.w), mulr(r.y, b1.w)), mulr(r.z, b2.w)), mulr(r.w, b3.w));
        out[(size_t)i * 4 + j] = o;
    }
}

// Matrix4::transform_vec3 (matrix.rs:160-171) of one point per matrix: M * (v, 1), rows summed left to right.
__global__ void __launch_bounds__(kThreads) mat4_transform_vec3_kernel(const float4* __restrict__ m, const float* __restrict__ v,
                                                                        uint32_t n, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float4 r = m[(size_t)i * 4 + j];
        out[3 * i + j] = addr(addr(addr(mulr(r.x, x), mulr(r.y, y)), mulr(r.z, z)), mulr(r.w, 1.0f));
    }
}

// Matrix4::try_inverse (matrix.rs:98-149): cofactor expansion; None when |det| < 1e-8 (IsZero for f32,
// number_traits.rs:80-84) -> some[i] = 0 and the output matrix is left untouched.
__global__ void __launch_bounds__(kThreads) mat4_try_inverse_kernel(const float4* __restrict__ m, uint32_t n,
                                                                     float4* __restrict__ out, uint8_t* __restrict__ some) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 r0 = m[(size_t)i * 4], r1 = m[(size_t)i * 4 + 1], r2 = m[(size_t)i * 4 + 2], r3 = m[(size_t)i * 4 + 3];
    const float e00 = r0.x, e01 = r0.y, e02 = r0.z, e03 = r0.w;
    const float e10 = r1.x, e11 = r1.y, e12 = r1.z, e13 = r1.w;
    const float e20 = r2.x, e21 = r2.y, e22 = r2.z, e23 = r2.w;
    const float e30 = r3.x, e31 = r3.y, e32 = r3.z, e33 = r3.w;
    auto det2 = [](float a, float b, float c, float d) { return subr(mulr(a, b), mulr(c, d)); };  // a*b - c*d
    const float a2323 = det2(e22, e33, e23, e32), a1323 = det2(e21, e33, e23, e31), a1223 = det2(e21, e32, e22, e31);
    const float a0323 = det2(e20, e33, e23, e30), a0223 = det2(e20, e32, e22, e30), a0123 = det2(e20, e31, e21, e30);
    const float a2313 = det2(e12, e33, e13, e32), a1313 = det2(e11, e33, e13, e31), a1213 = det2(e11, e32, e12, e31);
    const float a2312 = det2(e12, e23, e13, e22), a1312 = det2(e11, e23, e13, e21), a1212 = det2(e11, e22, e12, e21);
    const float a0313 = det2(e10, e33, e13, e30), a0213 = det2(e10, e32, e12, e30), a0312 = det2(e10, e23, e13, e20);
    const float a0212 = det2(e10, e22, e12, e20), a0113 = det2(e10, e31, e11, e30), a0112 = det2(e10, e21, e11, e20);
    // p*a - q*b + r*c, left to right
    auto tri = [](float p, float a, float q, float b, float r, float c) { return addr(subr(mulr(p, a), mulr(q, b)), mulr(r, c)); };
    const float det = subr(addr(subr(mulr(e00, tri(e11, a2323, e12, a1323, e13, a1223)),
                                     mulr(e01, tri(e10, a2323, e12, a0323, e13, a0223))),
                                mulr(e02, tri(e10, a1323, e11, a0323, e13, a0123))),
                           mulr(e03, tri(e10, a1223, e11, a0223, e12, a0123)));
    if (fabsf(det) < 0.00000001f) {
        some[i] = 0;
        return;
    }
    some[i] = 1;
    const float d = divr(1.0f, det);
    float4* o = out + (size_t)i * 4;
    o[0] = make_float4(mulr(d, tri(e11, a2323, e12, a1323, e13, a1223)), mulr(d, -tri(e01, a2323, e02, a1323, e03, a1223)),
                       mulr(d, tri(e01, a2313, e02, a1313, e03, a1213)), mulr(d, -tri(e01, a2312, e02, a1312, e03, a1212)));
    o[1] = make_float4(mulr(d, -tri(e10, a2323, e12, a0323, e13, a0223)), mulr(d, tri(e00, a2323, e02, a0323, e03, a0223)),
                       mulr(d, -tri(e00, a2313, e02, a0313, e03, a0213)), mulr(d, tri(e00, a2312, e02, a0312, e03, a0212)));
    o[2] = make_float4(mulr(d, tri(e10, a1323, e11, a0323, e13, a0123)), mulr(d, -tri(e00, a1323, e01, a0323, e03, a0123)),
                       mulr(d, tri(e00, a1313, e01, a0313, e03, a0113)), mulr(d, -tri(e00, a1312, e01, a0312, e03, a0112)));
    o[3] = make_float4(mulr(d, -tri(e10, a1223, e11, a0223, e12, a0123)), mulr(d, tri(e00, a1223, e01, a0223, e02, a0123)),
                       mulr(d, -tri(e00, a1213, e01, a0213, e02, a0113)), mulr(d, tri(e00, a1212, e01, a0212, e02, a0112)));
}

}  // namespace tb
