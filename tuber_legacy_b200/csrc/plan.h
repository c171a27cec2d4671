// Frame plan: which bits of the 64-bit sort key actually vary in this scene, how they are
// packed into a short key, and how that short key is split into radix digits.
// Shared by host code (plan construction, tests of the planner) and device code.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TB_HD __host__ __device__
#else
#define TB_HD
#endif

namespace tb {

constexpr int kMaxRuns = 16;
constexpr int kMaxPasses = 8;
constexpr int kMaxDigitBits = 8;
constexpr uint32_t kMaxLtBits = 16;  // (layer, texture) bins kept as a direct histogram up to 2^16
constexpr int kMaxZDict = 64;        // z dictionary: up to 64 distinct depths are sorted by their rank (6 bits)
constexpr int kZSetSlots = 512;      // open-addressing set that collects the distinct depths at plan time

// Bit-gather (pext) of the varying key bits, as up to 16 contiguous runs.
struct PextPlan {
    uint32_t nruns;
    uint32_t nbits;  // total extracted bits
    uint8_t src_shift[kMaxRuns];
    uint8_t width[kMaxRuns];
    uint8_t dst_shift[kMaxRuns];
};

struct FramePlan {
    PextPlan pext;
    uint64_t varying;   // key bits that differ between at least two renderables (possibly widened)
    uint64_t constant;  // value of the non-varying key bits
    uint32_t npasses;   // >= 1 radix passes, least significant digit first
    uint32_t shift[kMaxPasses];
    uint32_t bits[kMaxPasses];
    uint32_t ltbits;  // varying bits among key[63:32] (layer, texture)
    uint32_t zbits;   // varying bits among key[31:0]  (orderable world z)
    uint32_t key64;   // short key needs more than 32 bits
    // z dictionary (zdict_n > 0): the scene uses few distinct depths, so the z part of the short key
    // is the RANK of the orderable z among them (ceil(log2 n) bits) instead of its varying float bits
    // (0.0, 1.0, 2.0, 3.0 differ in 9 bits but are 4 values). `pext` then covers key[63:32] only.
    uint32_t zdict_n;
    uint32_t zdict[kMaxZDict];  // ascending orderable z values
};

TB_HD inline uint64_t pext_runs(uint64_t key, const PextPlan& p) {
    uint64_t out = 0;
    for (uint32_t r = 0; r < p.nruns; ++r) {
        uint64_t m = (p.width[r] >= 64) ? ~0ull : ((1ull << p.width[r]) - 1ull);
        out |= ((key >> p.src_shift[r]) & m) << p.dst_shift[r];
    }
    return out;
}

// Inverse of pext_runs on the varying bits (pdep).
TB_HD inline uint64_t pdep_runs(uint64_t shortkey, const PextPlan& p) {
    uint64_t out = 0;
    for (uint32_t r = 0; r < p.nruns; ++r) {
        uint64_t m = (p.width[r] >= 64) ? ~0ull : ((1ull << p.width[r]) - 1ull);
        out |= ((shortkey >> p.dst_shift[r]) & m) << p.src_shift[r];
    }
    return out;
}

// Short key of a full key under a plan. `zdict` is a copy of f.zdict the caller can index cheaply
// (shared memory on the device). A depth missing from the dictionary sets *miss (stale plan).
TB_HD inline uint64_t plan_short_key(const FramePlan& f, const uint32_t* zdict, uint64_t key, bool* miss) {
    const uint64_t sk = pext_runs(key, f.pext);
    if (f.zdict_n == 0) return sk;
    const uint32_t z = (uint32_t)key;
    uint32_t lo = 0, hi = f.zdict_n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (zdict[mid] < z) lo = mid + 1;
        else hi = mid;
    }
    const bool found = lo < f.zdict_n && zdict[lo] == z;
    if (!found) *miss = true;
    return (sk << f.zbits) | (found ? lo : 0u);
}

inline uint32_t popcount64(uint64_t v) {
    uint32_t c = 0;
    while (v) {
        v &= v - 1;
        ++c;
    }
    return c;
}

inline void split_plan_digits(FramePlan& f);

// Builds the plan from the OR and AND of every renderable's key. Runs of varying bits are
// kept as they are; when there are more than kMaxRuns of them the smallest gaps are filled
// (a filled gap bit is constant, so it only costs sort width, never correctness). Runs never
// straddle bit 32, so the short key is always [ (layer,texture) bits | z bits ].
inline void build_frame_plan(uint64_t key_or, uint64_t key_and, FramePlan* out) {
    FramePlan& f = *out;
    uint64_t varying = key_or & ~key_and;
    for (;;) {
        // collect runs, split at bit 32
        uint32_t nruns = 0;
        uint32_t starts[64], ends[64];
        uint32_t bit = 0;
        while (bit < 64) {
            if (!((varying >> bit) & 1ull)) {
                ++bit;
                continue;
            }
            uint32_t s = bit;
            while (bit < 64 && ((varying >> bit) & 1ull) && !(bit == 32 && s < 32)) ++bit;
            starts[nruns] = s;
            ends[nruns] = bit;
            ++nruns;
        }
        if (nruns <= (uint32_t)kMaxRuns) {
            f.pext.nruns = nruns;
            f.pext.nbits = 0;
            for (uint32_t r = 0; r < (uint32_t)kMaxRuns; ++r) f.pext.src_shift[r] = f.pext.width[r] = f.pext.dst_shift[r] = 0;
            for (uint32_t r = 0; r < nruns; ++r) {
                f.pext.src_shift[r] = (uint8_t)starts[r];
                f.pext.width[r] = (uint8_t)(ends[r] - starts[r]);
                f.pext.dst_shift[r] = (uint8_t)f.pext.nbits;
                f.pext.nbits += ends[r] - starts[r];
            }
            break;
        }
        // fill the smallest gap between two runs on the same side of bit 32 (a gap that
        // straddles bit 32 would be split again, so filling it cannot reduce the run count)
        uint32_t best = 0, best_gap = 65;
        for (uint32_t r = 0; r + 1 < nruns; ++r) {
            if (ends[r] <= 32 && starts[r + 1] >= 32) continue;
            uint32_t gap = starts[r + 1] - ends[r];
            if (gap < best_gap) {
                best_gap = gap;
                best = r;
            }
        }
        for (uint32_t b = ends[best]; b < starts[best + 1]; ++b) varying |= 1ull << b;
    }
    f.varying = varying;
    f.constant = key_and & ~varying;
    f.zbits = popcount64(varying & 0xFFFFFFFFull);
    f.ltbits = popcount64(varying >> 32);
    f.zdict_n = 0;
    split_plan_digits(f);
}

// The scatter passes move 8..12-byte (key, id) pairs and take full 8-bit digits; whatever is left
// goes to the LAST pass, the one fused with the emit: the fewer buckets it has, the longer the runs
// of consecutive output positions a tile produces and the better its 80/64-byte payload stores coalesce.
inline void split_plan_digits(FramePlan& f) {
    const uint32_t nbits = f.ltbits + f.zbits;
    f.key64 = nbits > 32 ? 1u : 0u;
    uint32_t np = (nbits + kMaxDigitBits - 1) / kMaxDigitBits;
    if (np == 0) np = 1;
    f.npasses = np;
    uint32_t sh = 0;
    for (uint32_t p = 0; p < (uint32_t)kMaxPasses; ++p) f.shift[p] = f.bits[p] = 0;
    for (uint32_t p = 0; p < np; ++p) {
        f.shift[p] = sh;
        f.bits[p] = (p + 1 < np) ? (uint32_t)kMaxDigitBits : nbits - sh;
        sh += f.bits[p];
    }
}

// Same plan, but the z part of the short key is the rank of z among the `n` distinct depths
// `zvals` (ascending) the scene uses — taken only when that is narrower than the varying z bits.
inline void build_frame_plan_dict(uint64_t key_or, uint64_t key_and, const uint32_t* zvals, uint32_t n, FramePlan* out) {
    build_frame_plan(key_or, key_and, out);
    FramePlan& f = *out;
    f.zdict_n = 0;
    if (n == 0 || n > (uint32_t)kMaxZDict) return;
    uint32_t rank_bits = 0;
    while ((1u << rank_bits) < n) ++rank_bits;
    if (rank_bits >= f.zbits) return;
    // re-plan the (layer, texture) half alone, then append the rank bits below it
    FramePlan hi;
    build_frame_plan(key_or & 0xFFFFFFFF00000000ull, key_and | 0x00000000FFFFFFFFull, &hi);
    f.pext = hi.pext;  // runs inside key[63:32] only, packed from bit 0
    f.pext.nbits = hi.ltbits + rank_bits;  // ... but nbits stays the width of the whole short key
    f.varying = (hi.varying & 0xFFFFFFFF00000000ull) | (f.varying & 0xFFFFFFFFull);
    f.constant = (hi.constant & 0xFFFFFFFF00000000ull) | (f.constant & 0xFFFFFFFFull);
    f.ltbits = hi.ltbits;
    f.zbits = rank_bits;
    f.zdict_n = n;
    for (uint32_t k = 0; k < n; ++k) f.zdict[k] = zvals[k];
    split_plan_digits(f);
}

// A plan stays valid for a frame whose keys vary only inside plan.varying and agree with
// plan.constant elsewhere (with a z dictionary the low half is validated by dictionary misses).
inline bool plan_covers(const FramePlan& f, uint64_t key_or, uint64_t key_and) {
    const uint64_t m = f.zdict_n ? 0xFFFFFFFF00000000ull : ~0ull;
    uint64_t varying = (key_or & ~key_and) & m;
    if (varying & ~f.varying) return false;
    return (key_and & ~f.varying & m) == (f.constant & m) && (key_or & ~f.varying & m) == (f.constant & m);
}

// Field-wise equality (struct padding and unused dictionary slots do not count).
inline bool plan_equal(const FramePlan& a, const FramePlan& b) {
    if (a.pext.nruns != b.pext.nruns || a.pext.nbits != b.pext.nbits) return false;
    for (int r = 0; r < kMaxRuns; ++r)
        if (a.pext.src_shift[r] != b.pext.src_shift[r] || a.pext.width[r] != b.pext.width[r] ||
            a.pext.dst_shift[r] != b.pext.dst_shift[r])
            return false;
    if (a.varying != b.varying || a.constant != b.constant || a.npasses != b.npasses || a.ltbits != b.ltbits ||
        a.zbits != b.zbits || a.key64 != b.key64 || a.zdict_n != b.zdict_n)
        return false;
    for (int p = 0; p < kMaxPasses; ++p)
        if (a.shift[p] != b.shift[p] || a.bits[p] != b.bits[p]) return false;
    for (uint32_t k = 0; k < a.zdict_n; ++k)
        if (a.zdict[k] != b.zdict[k]) return false;
    return true;
}

}  // namespace tb
