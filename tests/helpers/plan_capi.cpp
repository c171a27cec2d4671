// Host-only C surface over tuber_legacy_b200/csrc/plan.h (the frame planner is plain C++ shared by
// host and device code) so its logic can be tested on the CPU.
#include <cstring>

#include "../../tuber_legacy_b200/csrc/plan.h"

extern "C" {
void plan_build(uint64_t key_or, uint64_t key_and, tb::FramePlan* out) { tb::build_frame_plan(key_or, key_and, out); }
uint64_t plan_pext(const tb::FramePlan* f, uint64_t key) { return tb::pext_runs(key, f->pext); }
uint64_t plan_pdep(const tb::FramePlan* f, uint64_t sk) { return tb::pdep_runs(sk, f->pext); }
int plan_covers(const tb::FramePlan* f, uint64_t key_or, uint64_t key_and) { return tb::plan_covers(*f, key_or, key_and) ? 1 : 0; }
void plan_build_dict(uint64_t key_or, uint64_t key_and, const uint32_t* zvals, uint32_t n, tb::FramePlan* out) {
    tb::build_frame_plan_dict(key_or, key_and, zvals, n, out);
}
uint64_t plan_short_key(const tb::FramePlan* f, uint64_t key, int* miss) {
    bool m = false;
    const uint64_t sk = tb::plan_short_key(*f, f->zdict, key, &m);
    *miss = m ? 1 : 0;
    return sk;
}
int plan_equal(const tb::FramePlan* a, const tb::FramePlan* b) { return tb::plan_equal(*a, *b) ? 1 : 0; }
uint32_t plan_sizeof() { return (uint32_t)sizeof(tb::FramePlan); }
}
