// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load anything under oracle/.
//
// CPU restatement of the reference's tuber-math / tuber-core arithmetic on the hot path.
// Every operation is an individually rounded f32 multiply or add, summed left to right,
// exactly as rustc compiles the reference (rustc never contracts a*b+c into an FMA).
// Build with -O2 -ffp-contract=off and without -ffast-math (see oracle/Makefile).
//
// Parity status: PINNED for Matrix4::mul (exact i32 product, matrix.rs:309-341),
// Quaternion::from_euler (quaternion.rs:250-259) and rotation_matrix (quaternion.rs:180-201),
// try_inverse (matrix.rs:381-407) — the reference's own known-answer tests are replayed in
// tests/test_oracle_math.py. as_matrix4, transform_vec3 and the column-major conversion
// have no reference test; they follow the source text only.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace ref {

// tuber-math/src/vector.rs:9-187 — only the plain-old-data view and Neg (vector.rs:142-153)
// are on the hot path.
struct Vec3 {
    float x, y, z;
};
inline Vec3 neg(const Vec3& v) { return Vec3{-v.x, -v.y, -v.z}; }

// tuber-math/src/matrix.rs:9-12 — row-major values[16], column-vector convention.
template <typename T>
struct Mat4 {
    T v[16];
};
using Mat4f = Mat4<float>;

// matrix.rs:253-270
template <typename T>
inline Mat4<T> identity() {
    Mat4<T> m;
    for (int i = 0; i < 16; ++i) m.v[i] = T(0);
    m.v[0] = m.v[5] = m.v[10] = m.v[15] = T(1);
    return m;
}

// matrix.rs:61-71 — translation sits in elements 3, 7, 11.
inline Mat4f new_translation(const Vec3& t) {
    Mat4f m = identity<float>();
    m.v[3] = t.x;
    m.v[7] = t.y;
    m.v[11] = t.z;
    return m;
}

// matrix.rs:80-90
inline Mat4f new_scale(const Vec3& s) {
    Mat4f m = identity<float>();
    m.v[0] = s.x;
    m.v[5] = s.y;
    m.v[10] = s.z;
    return m;
}

// matrix.rs:41-58 — Matrix4::new_orthographic (unused by the reference itself; the camera a batching
// render_scene folds into its emit).
inline Mat4f new_orthographic(float left, float right, float bottom, float top, float near_, float far_) {
    Mat4f m;
    for (int k = 0; k < 16; ++k) m.v[k] = 0.0f;
    m.v[0] = 2.0f / (right - left);
    m.v[3] = -((right + left) / (right - left));
    m.v[5] = 2.0f / (top - bottom);
    m.v[7] = -((top + bottom) / (top - bottom));
    m.v[10] = -2.0f / (far_ - near_);
    m.v[11] = -((far_ + near_) / (far_ - near_));
    m.v[15] = 1.0f;
    return m;
}

// matrix.rs:174-194 — out[j][i] = a[j][0]*b[0][i] + a[j][1]*b[1][i] + a[j][2]*b[2][i]
// + a[j][3]*b[3][i], evaluated left to right.
template <typename T>
inline Mat4<T> mul(const Mat4<T>& a, const Mat4<T>& b) {
    Mat4<T> r;
    for (int j = 0; j < 4; ++j) {
        for (int i = 0; i < 4; ++i) {
            T p0 = a.v[j * 4 + 0] * b.v[i];
            T p1 = a.v[j * 4 + 1] * b.v[i + 4];
            T p2 = a.v[j * 4 + 2] * b.v[i + 8];
            T p3 = a.v[j * 4 + 3] * b.v[i + 12];
            T s = p0 + p1;
            s = s + p2;
            s = s + p3;
            r.v[j * 4 + i] = s;
        }
    }
    return r;
}

// matrix.rs:160-171 — M·(v,1), each row summed left to right; first three components kept.
inline Vec3 transform_vec3(const Mat4f& m, const Vec3& p) {
    const float w = 1.0f;
    float out[4];
    for (int j = 0; j < 4; ++j) {
        float s = m.v[j * 4 + 0] * p.x;
        s = s + m.v[j * 4 + 1] * p.y;
        s = s + m.v[j * 4 + 2] * p.z;
        s = s + m.v[j * 4 + 3] * w;
        out[j] = s;
    }
    return Vec3{out[0], out[1], out[2]};
}

// matrix.rs:219-251 — the only GPU-facing matrix layout the reference defines:
// [[v0,v4,v8,v12],[v1,v5,v9,v13],[v2,v6,v10,v14],[v3,v7,v11,v15]] (column-major).
inline void to_column_major(const Mat4f& m, float out[16]) {
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out[c * 4 + r] = m.v[r * 4 + c];
}

// matrix.rs:98-149 — cofactor inverse; is_zero(det) is |det| < 1e-8 (number_traits.rs:80-84).
// Off the per-frame path; restated so the reference's own test (matrix.rs:381-407) can pin
// the row-major indexing convention of this file.
inline bool try_inverse(const Mat4f& s, Mat4f* out) {
    auto e = [&](int r, int c) { return s.v[r * 4 + c]; };
    float a2323 = e(2, 2) * e(3, 3) - e(2, 3) * e(3, 2);
    float a1323 = e(2, 1) * e(3, 3) - e(2, 3) * e(3, 1);
    float a1223 = e(2, 1) * e(3, 2) - e(2, 2) * e(3, 1);
    float a0323 = e(2, 0) * e(3, 3) - e(2, 3) * e(3, 0);
    float a0223 = e(2, 0) * e(3, 2) - e(2, 2) * e(3, 0);
    float a0123 = e(2, 0) * e(3, 1) - e(2, 1) * e(3, 0);
    float a2313 = e(1, 2) * e(3, 3) - e(1, 3) * e(3, 2);
    float a1313 = e(1, 1) * e(3, 3) - e(1, 3) * e(3, 1);
    float a1213 = e(1, 1) * e(3, 2) - e(1, 2) * e(3, 1);
    float a2312 = e(1, 2) * e(2, 3) - e(1, 3) * e(2, 2);
    float a1312 = e(1, 1) * e(2, 3) - e(1, 3) * e(2, 1);
    float a1212 = e(1, 1) * e(2, 2) - e(1, 2) * e(2, 1);
    float a0313 = e(1, 0) * e(3, 3) - e(1, 3) * e(3, 0);
    float a0213 = e(1, 0) * e(3, 2) - e(1, 2) * e(3, 0);
    float a0312 = e(1, 0) * e(2, 3) - e(1, 3) * e(2, 0);
    float a0212 = e(1, 0) * e(2, 2) - e(1, 2) * e(2, 0);
    float a0113 = e(1, 0) * e(3, 1) - e(1, 1) * e(3, 0);
    float a0112 = e(1, 0) * e(2, 1) - e(1, 1) * e(2, 0);

    float det = e(0, 0) * (e(1, 1) * a2323 - e(1, 2) * a1323 + e(1, 3) * a1223) -
                e(0, 1) * (e(1, 0) * a2323 - e(1, 2) * a0323 + e(1, 3) * a0223) +
                e(0, 2) * (e(1, 0) * a1323 - e(1, 1) * a0323 + e(1, 3) * a0123) -
                e(0, 3) * (e(1, 0) * a1223 - e(1, 1) * a0223 + e(1, 2) * a0123);
    if (std::fabs(det) < 0.00000001f) return false;
    float d = 1.0f / det;
    float* o = out->v;
    o[0] = d * (e(1, 1) * a2323 - e(1, 2) * a1323 + e(1, 3) * a1223);
    o[1] = d * -(e(0, 1) * a2323 - e(0, 2) * a1323 + e(0, 3) * a1223);
    o[2] = d * (e(0, 1) * a2313 - e(0, 2) * a1313 + e(0, 3) * a1213);
    o[3] = d * -(e(0, 1) * a2312 - e(0, 2) * a1312 + e(0, 3) * a1212);
    o[4] = d * -(e(1, 0) * a2323 - e(1, 2) * a0323 + e(1, 3) * a0223);
    o[5] = d * (e(0, 0) * a2323 - e(0, 2) * a0323 + e(0, 3) * a0223);
    o[6] = d * -(e(0, 0) * a2313 - e(0, 2) * a0313 + e(0, 3) * a0213);
    o[7] = d * (e(0, 0) * a2312 - e(0, 2) * a0312 + e(0, 3) * a0212);
    o[8] = d * (e(1, 0) * a1323 - e(1, 1) * a0323 + e(1, 3) * a0123);
    o[9] = d * -(e(0, 0) * a1323 - e(0, 1) * a0323 + e(0, 3) * a0123);
    o[10] = d * (e(0, 0) * a1313 - e(0, 1) * a0313 + e(0, 3) * a0113);
    o[11] = d * -(e(0, 0) * a1312 - e(0, 1) * a0312 + e(0, 3) * a0112);
    o[12] = d * -(e(1, 0) * a1223 - e(1, 1) * a0223 + e(1, 2) * a0123);
    o[13] = d * (e(0, 0) * a1223 - e(0, 1) * a0223 + e(0, 2) * a0123);
    o[14] = d * -(e(0, 0) * a1213 - e(0, 1) * a0213 + e(0, 2) * a0113);
    o[15] = d * (e(0, 0) * a1212 - e(0, 1) * a0212 + e(0, 2) * a0112);
    return true;
}

// tuber-math/src/quaternion.rs:8-15
struct Quat {
    float w, x, y, z;
};

// quaternion.rs:39-59 — roll = angles.x, pitch = angles.y, yaw = angles.z; half() is
// `* 0.5` (number_traits.rs:144-146); sin/cos are platform libm sinf/cosf
// (number_traits.rs:137-143 → Rust std → LLVM intrinsic → glibc on linux-gnu).
inline Quat from_euler(const Vec3& a) {
    float half_roll = a.x * 0.5f;
    float half_pitch = a.y * 0.5f;
    float half_yaw = a.z * 0.5f;
    float cy = cosf(half_yaw);
    float sy = sinf(half_yaw);
    float cp = cosf(half_pitch);
    float sp = sinf(half_pitch);
    float cr = cosf(half_roll);
    float sr = sinf(half_roll);
    Quat q;
    q.w = cr * cp * cy + sr * sp * sy;
    q.x = sr * cp * cy - cr * sp * sy;
    q.y = cr * sp * cy + sr * cp * sy;
    q.z = cr * cp * sy - sr * sp * cy;
    return q;
}

// quaternion.rs:63-90 — the quaternion is NOT normalised first.
inline Mat4f rotation_matrix(const Quat& q) {
    float w = q.w, x = q.x, y = q.y, z = q.z;
    float x2 = x + x, y2 = y + y, z2 = z + z, w2 = w + w;
    float xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z;
    float yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
    float wx2 = w2 * x, wy2 = w2 * y, wz2 = w2 * z;
    Mat4f m;
    float* o = m.v;
    o[0] = 1.0f - yy2 - zz2;
    o[1] = xy2 - wz2;
    o[2] = xz2 + wy2;
    o[3] = 0.0f;
    o[4] = xy2 + wz2;
    o[5] = 1.0f - xx2 - zz2;
    o[6] = yz2 - wx2;
    o[7] = 0.0f;
    o[8] = xz2 - wy2;
    o[9] = yz2 + wx2;
    o[10] = 1.0f - xx2 - yy2;
    o[11] = 0.0f;
    o[12] = 0.0f;
    o[13] = 0.0f;
    o[14] = 0.0f;
    o[15] = 1.0f;
    return m;
}

// quaternion.rs:132-158 — Hamilton product, every component summed in the reference's own term order
// (x1*w2 + y1*z2 - z1*y2 + w1*x2, ...); off the per-frame path, pinned by the reference's test (quaternion.rs:167-177).
inline Quat quat_mul(const Quat& a, const Quat& b) {
    const float x1 = a.x, y1 = a.y, z1 = a.z, w1 = a.w;
    const float x2 = b.x, y2 = b.y, z2 = b.z, w2 = b.w;
    Quat r;
    r.w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
    r.x = x1 * w2 + y1 * z2 - z1 * y2 + w1 * x2;
    r.y = y1 * w2 + z1 * x2 + w1 * y2 - x1 * z2;
    r.z = z1 * w2 + w1 * z2 + x1 * y2 - y1 * x2;
    return r;
}

// quaternion.rs:28-37
inline Quat from_axis_angle(const Vec3& axis, float angle) {
    const float half_angle = angle * 0.5f;
    const float s = sinf(half_angle);
    Quat q;
    q.x = axis.x * s;
    q.y = axis.y * s;
    q.z = axis.z * s;
    q.w = cosf(half_angle);
    return q;
}

// quaternion.rs:106-117 (norm) and :92-96 (normalize): vector part and scalar part divided by the norm
// (vector.rs:132-140: component / rhs).
inline float quat_norm(const Quat& q) {
    const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, ww = q.w * q.w;
    return std::sqrt(ww + xx + yy + zz);
}
inline Quat quat_normalized(const Quat& q) {
    const float norm = quat_norm(q);
    return Quat{q.w / norm, q.x / norm, q.y / norm, q.z / norm};
}

// tuber-core/src/transform.rs:5-11 — 4 × Vector3<f32> = 48 bytes, this field order.
struct Transform {
    Vec3 translation;
    Vec3 angle;
    Vec3 rotation_center;
    Vec3 scale;
};
static_assert(sizeof(Transform) == 48, "Transform must be 48 bytes like the reference");

// transform.rs:13-22
inline Transform transform_default() {
    return Transform{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {1, 1, 1}};
}

// transform.rs:28-38 — (((S · T(t)) · T(c)) · R) · T(−c), four dense 4×4 products,
// left-associative.
inline Mat4f as_matrix4(const Transform& t) {
    Mat4f m = mul(new_scale(t.scale), new_translation(t.translation));
    m = mul(m, new_translation(t.rotation_center));
    m = mul(m, rotation_matrix(from_euler(t.angle)));
    m = mul(m, new_translation(neg(t.rotation_center)));
    return m;
}

// The closed form the device kernels evaluate: value-identical to
// as_matrix4 for finite inputs; only the sign of zeros may differ. Kept here so the CPU
// test-suite can prove that equivalence without a GPU (tests/test_oracle_math.py).
inline Mat4f as_matrix4_closed(const Transform& t) {
    Mat4f R = rotation_matrix(from_euler(t.angle));
    const float s[3] = {t.scale.x, t.scale.y, t.scale.z};
    const float c[3] = {t.rotation_center.x, t.rotation_center.y, t.rotation_center.z};
    const float tr[3] = {t.translation.x, t.translation.y, t.translation.z};
    Mat4f m;
    for (int j = 0; j < 3; ++j) {
        float r0 = s[j] * R.v[j * 4 + 0];
        float r1 = s[j] * R.v[j * 4 + 1];
        float r2 = s[j] * R.v[j * 4 + 2];
        float sc = s[j] * c[j];
        float st = s[j] * tr[j];
        float trj = sc + st;
        float a = r0 * (-c[0]);
        float b = r1 * (-c[1]);
        float d = r2 * (-c[2]);
        float acc = a + b;
        acc = acc + d;
        acc = acc + trj;
        m.v[j * 4 + 0] = r0;
        m.v[j * 4 + 1] = r1;
        m.v[j * 4 + 2] = r2;
        m.v[j * 4 + 3] = acc;
    }
    m.v[12] = m.v[13] = m.v[14] = 0.0f;
    m.v[15] = 1.0f;
    return m;
}

}  // namespace ref
