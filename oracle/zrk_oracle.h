/*
 * zrk_oracle.h — CPU restatement of the reference tick.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this.  The product path (zrk_modulation_b200/, libzrk_b200.so) never does.
 *
 * Parity status: PINNED.  The reference is pure Python (no golden vectors of its own, SURVEY
 * §8c), so the pin is the reference itself imported in the build container:
 * tests/golden/make_golden.py runs /root/reference unmodified except for the noise injection
 * of SURVEY Appendix C and stores its canonical message log; tests/test_oracle_golden.py
 * checks this restatement against every stored log bit for bit (fp64).
 */
#ifndef ZRK_ORACLE_H
#define ZRK_ORACLE_H

#include "../include/zrk_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* numpy evaluates `b ** 2` (Missile.py:81) and `(...) ** 0.5` (CCP.py:297) on float64 scalars
 * through libm pow(), which is not always the correctly rounded square / square root.  With
 * this flag the oracle calls pow() too (bit-faithful to the reference run in this container);
 * without it it uses b*b and sqrt(), which is what the CUDA path computes. */
#define ZO_NUMPY_POW 1
#define ZO_KEEP_MSGS 2

typedef struct zo_run {
    zrk_msg *msgs;
    size_t n_msgs, cap_msgs;
    int32_t *ids; /* id pool of ACTIVE/ALL/FOUND lists */
    size_t n_ids, cap_ids;
    int32_t fault_code, fault_tick;
    int64_t counts[ZRK_MSG_NTYPES];
} zo_run;

/* Run scenario `scenario` of the batch for n_ticks ticks (Manager.py:111-140).  `state`, if
 * not NULL, receives the final state with the layout of zrk_read_state for a batch of ONE
 * scenario (every pointer addresses this scenario's rows only). */
int zrk_oracle_run(const zrk_batch_desc *d, int scenario, int n_ticks, int flags, zo_run *out,
                   const zrk_state_view *state);
void zrk_oracle_free(zo_run *r);

/* Run scenarios [s0, s1) on n_threads host threads, keeping only counters.
 * counts[(s - s0) * ZRK_MSG_NTYPES + type]; fault arrays [s1 - s0] (may be NULL). */
int zrk_oracle_run_batch(const zrk_batch_desc *d, int s0, int s1, int n_ticks, int flags,
                         int n_threads, int64_t *counts, int32_t *fault_code, int32_t *fault_tick);

#ifdef __cplusplus
}
#endif
#endif
