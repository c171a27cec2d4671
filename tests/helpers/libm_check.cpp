// Test helper (not product): the restated glibc sinf / cosf of csrc/libm_sincosf.h, built for the host with
// -ffp-contract=off, compared bit for bit with the real sinf / cosf of this box's libm — the functions the
// reference reaches through f32::sin / f32::cos (crates/tuber-math/src/number_traits.rs:137-143).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../tuber_legacy_b200/csrc/libm_sincosf.h"

static inline float as_float(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint32_t as_uint(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

extern "C" {

// glibc's own selector (sysdeps/x86_64/fpu/multiarch/ifunc-fma.h): the FMA build when the CPU has FMA and AVX2.
int libm_uses_fma(void) {
    __builtin_cpu_init();
    return __builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2");
}

// Bit patterns first, first + stride, ... (count of them): how many differ from libm, for sin and cos.
// NaN results compare equal to NaN results. *first_bad receives the first differing input pattern.
void libm_compare(uint32_t first, uint32_t stride, uint64_t count, int fma, uint64_t* bad_sin, uint64_t* bad_cos,
                  uint32_t* first_bad) {
    uint64_t bs = 0, bc = 0;
    uint32_t u = first;
    for (uint64_t k = 0; k < count; ++k, u += stride) {
        const float y = as_float(u);
        float s, c;
        if (fma) tbm::sincosf_ref<true>(y, &s, &c);
        else tbm::sincosf_ref<false>(y, &s, &c);
        const float rs = sinf(y), rc = cosf(y);
        const bool es = as_uint(s) == as_uint(rs) || (s != s && rs != rs);
        const bool ec = as_uint(c) == as_uint(rc) || (c != c && rc != rc);
        if (!(es && ec) && bs + bc == 0 && first_bad) *first_bad = u;
        bs += !es;
        bc += !ec;
    }
    *bad_sin = bs;
    *bad_cos = bc;
}

// The restatement and libm on an array (used to check the device against both).
void libm_eval(const float* y, uint64_t n, int fma, float* s_ref, float* c_ref, float* s_libm, float* c_libm) {
    for (uint64_t k = 0; k < n; ++k) {
        if (fma) tbm::sincosf_ref<true>(y[k], &s_ref[k], &c_ref[k]);
        else tbm::sincosf_ref<false>(y[k], &s_ref[k], &c_ref[k]);
        s_libm[k] = sinf(y[k]);
        c_libm[k] = cosf(y[k]);
    }
}
}
