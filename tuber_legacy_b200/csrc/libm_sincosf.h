// sinf / cosf exactly as the reference computes them.
//
// The reference calls f32::sin / f32::cos (crates/tuber-math/src/number_traits.rs:137-143), which rustc lowers
// to the platform libm: on linux-gnu that is glibc's sinf / cosf, i.e. the ARM optimized-routines algorithm glibc
// adopted in 2.28 (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h, s_sincosf_data.c; pinned here: glibc
// 2.39-0ubuntu8.5, the libm of this image): argument reduction and a degree-7/8 polynomial evaluated in DOUBLE,
// rounded to float once. Every step is a deterministic IEEE operation, so the device can reproduce the host's
// result bit for bit — which makes every float the frame emits (and therefore world z, the sort key and the draw
// order under 3-D rotations and rotated parents) identical to the reference's, not merely within 1e-5.
//
// glibc ships two builds of these functions and picks one at load time (sysdeps/x86_64/fpu/multiarch/s_sinf.c):
// the generic build rounds every product and sum on its own; the FMA build (CPUs with FMA + AVX2: every x86
// server since 2013) is the same source compiled with -mfma -mavx2, where gcc contracts `a + b * c` into one
// fused operation. The contraction pattern below was read off the disassembly of this image's libm.so.6
// (__sinf_fma at 0x7e800, __cosf_fma at 0x7e330). `FMA` selects the build; the host picks the one its own libm
// uses (tb_ctx_create, same CPU-feature test as glibc's selector).
//
// This header is shared by the device code (compose.cuh) and by a host build (tests/helpers/libm_check.c) that
// checks it against the real sinf / cosf of the box, so the restatement is pinned on the CPU as well.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TBM_HD __host__ __device__ __forceinline__
#else
#define TBM_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define TBM_MUL(a, b) __dmul_rn((a), (b))
#define TBM_ADD(a, b) __dadd_rn((a), (b))
#define TBM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define TBM_D2F(a) __double2float_rn(a)
#define TBM_D2I_RZ(a) __double2int_rz(a)
#define TBM_ASUINT(f) __float_as_uint(f)
#else
#include <math.h>
#include <string.h>
// host build: compiled with -ffp-contract=off, so plain operators round individually
#define TBM_MUL(a, b) ((a) * (b))
#define TBM_ADD(a, b) ((a) + (b))
#define TBM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define TBM_D2F(a) ((float)(a))
#define TBM_D2I_RZ(a) ((int32_t)(a))
static inline uint32_t tbm_asuint_host(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
#define TBM_ASUINT(f) tbm_asuint_host(f)
#endif

namespace tbm {

// __sincosf_table[0] (s_sincosf_data.c); table[1] is the same with the cosine coefficients negated.
#define TBM_HPI_INV 0x1.45F306DC9C883p+23 /* 2/pi * 2^24 */
#define TBM_HPI 0x1.921FB54442D18p0       /* pi/2 */
#define TBM_PI63 0x1.921FB54442D18p-62    /* pi * 2^-63 ... as glibc's pi63 */
#define TBM_C0 0x1p0
#define TBM_C1 -0x1.ffffffd0c621cp-2
#define TBM_C2 0x1.55553e1068f19p-5
#define TBM_C3 -0x1.6c087e89a359dp-10
#define TBM_C4 0x1.99343027bf8c3p-16
#define TBM_S1 -0x1.555545995a603p-3
#define TBM_S2 0x1.1107605230bc4p-7
#define TBM_S3 -0x1.994eb3774cf24p-13

template <bool FMA>
TBM_HD double mad(double a, double b, double c) {  // c + a * b as the selected libm build evaluates it
    return FMA ? TBM_FMA(a, b, c) : TBM_ADD(c, TBM_MUL(a, b));
}

// sinf_poly, even n (sincosf.h): x + x^3 s1 + x^7 (s2 + x^2 s3), double.
template <bool FMA>
TBM_HD double sin_poly(double x, double x2) {
    const double x3 = TBM_MUL(x, x2);
    const double s1 = mad<FMA>(x2, TBM_S3, TBM_S2);
    const double x7 = TBM_MUL(x3, x2);
    const double s = mad<FMA>(x3, TBM_S1, x);
    return mad<FMA>(x7, s1, s);
}
// sinf_poly, odd n: (c0 + x^2 c1) + x^4 c2 + x^6 (c3 + x^2 c4); `neg` = table[1] (quadrants 2, 3).
template <bool FMA>
TBM_HD double cos_poly(double x2, bool neg) {
    const double c0 = neg ? -TBM_C0 : TBM_C0, c1k = neg ? -TBM_C1 : TBM_C1, c2k = neg ? -TBM_C2 : TBM_C2;
    const double c3 = neg ? -TBM_C3 : TBM_C3, c4 = neg ? -TBM_C4 : TBM_C4;
    const double x4 = TBM_MUL(x2, x2);
    const double c2 = mad<FMA>(x2, c4, c3);
    const double c1 = mad<FMA>(x2, c1k, c0);
    const double x6 = TBM_MUL(x4, x2);
    const double c = mad<FMA>(x4, c2k, c1);
    return mad<FMA>(x6, c2, c);
}

TBM_HD uint32_t abstop12(float x) { return (TBM_ASUINT(x) >> 20) & 0x7ffu; }

// __inv_pio4[k] (s_sincosf_data.c): the bits of 4/pi as overlapping 32-bit windows, window k ending at byte k of
// a2f9836e 4e441529 fc2757d1 f534ddc0 db629599 3c439041 (kept in registers: no table in local memory).
TBM_HD uint32_t inv_pio4(uint32_t k) {
    const uint32_t p = 8u * (23u - k);  // bit position of the window's least significant bit
    const uint32_t j = p >> 6, off = p & 63u;
    const uint64_t w0 = 0xdb6295993c439041ull, w1 = 0xfc2757d1f534ddc0ull, w2 = 0xa2f9836e4e441529ull;
    const uint64_t lo = j == 0 ? w0 : j == 1 ? w1 : w2;
    const uint64_t hi = j == 0 ? w1 : j == 1 ? w2 : 0ull;
    return (uint32_t)(off ? (lo >> off) | (hi << (64u - off)) : lo);
}

// reduce_large (sincosf.h): |x| >= 120, 4/pi from the table, result in [-pi/4, pi/4] and the quadrant.
TBM_HD double reduce_large(uint32_t xi, int* np) {
    const uint32_t i = (xi >> 26) & 15u;
    const uint32_t arr[9] = {inv_pio4(i), 0, 0, 0, inv_pio4(i + 4), 0, 0, 0, inv_pio4(i + 8)};
    const int shift = (int)((xi >> 23) & 7u);
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    res0 = (uint64_t)(uint32_t)(xi * arr[0]);
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ull << 61)) >> 62;
    res0 -= n << 62;
    const double x = (double)(int64_t)res0;
    *np = (int)n;
    return TBM_MUL(x, TBM_PI63);
}

// sinf(y) and cosf(y) of the same argument (the reference needs both of every angle: quaternion.rs:39-59). The
// two glibc functions share their reduction, so it is done once; each output is exactly what the separate call
// returns. Non-finite input gives NaN (glibc: __math_invalidf).
template <bool FMA>
TBM_HD void sincosf_ref(float y, float* sin_out, float* cos_out) {
    double x = (double)y;
    const uint32_t top = abstop12(y);
    if (top < 0x3f4u) {  // |y| < pi/4 (abstop12(pio4) == 0x3f4)
        if (top < 0x398u) {  // |y| < 2^-12
            *sin_out = y;
            *cos_out = 1.0f;
            return;
        }
        const double x2 = TBM_MUL(x, x);
        *sin_out = TBM_D2F(sin_poly<FMA>(x, x2));
        *cos_out = TBM_D2F(cos_poly<FMA>(x2, false));
        return;
    }
    int n;
    int sign = 0;
    if (top < 0x42fu) {  // |y| < 120: reduce_fast
        const double r = TBM_MUL(x, TBM_HPI_INV);
        n = (TBM_D2I_RZ(r) + 0x800000) >> 24;
        x = mad<FMA>(-(double)n, TBM_HPI, x);
    } else if (top < 0x7f8u) {
        const uint32_t xi = TBM_ASUINT(y);
        sign = (int)(xi >> 31);
        x = reduce_large(xi, &n);
    } else {
        const float nan = TBM_D2F(TBM_MUL(0.0, (double)(y - y)));  // inf - inf / nan - nan
        *sin_out = *cos_out = nan;
        return;
    }
    const int q = n + sign;              // quadrant incl. the original sign (reduce_large works on |y|)
    const double s = ((q & 3) == 1 || (q & 3) == 2) ? -1.0 : 1.0;  // p->sign[q & 3] = {1, -1, -1, 1}
    const bool neg = (q & 2) != 0;       // table[1]
    const double xs = TBM_MUL(x, s);
    const double x2 = TBM_MUL(x, x);
    const double ps = sin_poly<FMA>(xs, x2);
    const double pc = cos_poly<FMA>(x2, neg);
    // sinf: sinf_poly(xs, x2, p, n); cosf: sinf_poly(xs, x2, p, n ^ 1)
    *sin_out = TBM_D2F((n & 1) ? pc : ps);
    *cos_out = TBM_D2F((n & 1) ? ps : pc);
}

}  // namespace tbm
