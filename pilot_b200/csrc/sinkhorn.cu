// C-ABI entry of kernel (3): dispatch between the batched shared-Gibbs-kernel solver and
// the reference-form kernel (reference call site: pilotpy/tools/Trajectory.py:513-515).
#include <stdlib.h>
#include "sinkhorn.cuh"

namespace pilot {

struct SkWs {
    unsigned long long *counter_fast, *n_redo, *counter_slow;
    double *setup, *scratch;
    long long *redo;
};

// every slot of the panel kernel could in principle be handed over to the tail kernel
static size_t sk_tail_slots() { return (size_t)sm_count() * skb_slots_per_cta(); }
static size_t sk_tail_rec_bytes() { return (sk_tail_slots() * sizeof(SkTailRec) + 255) / 256 * 256; }
static size_t sk_tail_uv_bytes(int KP) { return sk_tail_slots() * 2 * KP * sizeof(double); }

static size_t sk_ws_bytes(int K)
{
    const int KP = skb_pad(K <= 64 ? K : 64);
    return 256 + skb_setup_bytes(KP) + skb_scratch_bytes(KP, sm_count()) + (size_t)SK_REDO_CAP * sizeof(long long) +
           sk_tail_rec_bytes() + sk_tail_uv_bytes(KP);
}

size_t sinkhorn_ws_bytes(int K) { return sk_ws_bytes(K); }

}  // namespace pilot

extern "C" int pilot_sinkhorn_pairs(const double *props, int S, int K, const double *cost, double reg,
                                    int num_iter_max, double stop_thr, double tau, int check_every,
                                    const pilot_pair_range *range, int algo, int precision, double *out,
                                    int32_t *iters,
                                    int32_t *absorptions, int32_t *status, void *workspace,
                                    size_t workspace_bytes, void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(props && cost && out && workspace, "pilot_sinkhorn_pairs: NULL pointer");
    PILOT_CHECK_ARG(S >= 1 && K >= 1, "pilot_sinkhorn_pairs: S=%d K=%d", S, K);
    PILOT_CHECK_ARG(reg > 0.0, "pilot_sinkhorn_pairs: reg must be > 0");
    PILOT_CHECK_ARG(num_iter_max >= 1 && check_every >= 1, "pilot_sinkhorn_pairs: bad iteration parameters");
    PILOT_CHECK_ARG(algo == 0 || algo == 1 || algo == 3, "pilot_sinkhorn_pairs: algo %d", algo);
    PILOT_CHECK_ARG(precision == PILOT_F64 || precision == PILOT_F32, "pilot_sinkhorn_pairs: precision=%d", precision);
    PILOT_CHECK_ARG(sinkhorn_ref_smem(K) <= 200 * 1024, "pilot_sinkhorn_pairs: K=%d too large (max ~150)", K);
    PILOT_CHECK_ARG(workspace_bytes >= sk_ws_bytes(K), "pilot_sinkhorn_pairs: workspace %zu < %zu bytes",
                    workspace_bytes, sk_ws_bytes(K));
    PairMap pm;
    int rc = make_pair_map(range, S, &pm);
    if (rc) return rc;
    if (pm.n_local == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SkParams prm{reg, stop_thr, tau, num_iter_max, check_every};

    unsigned char *p = (unsigned char *)workspace;
    SkWs ws;
    ws.counter_fast = (unsigned long long *)p;
    ws.n_redo = ws.counter_fast + 1;
    ws.counter_slow = ws.counter_fast + 2;
    PILOT_CUDA(cudaMemsetAsync(p, 0, 256, st));
    if (algo == 1 || K > 64)
        return sinkhorn_ref_launch(props, K, cost, prm, pm, nullptr, nullptr, 0, out, iters, absorptions, status,
                                   ws.counter_slow, st);
    const int KP = skb_pad(K);
    ws.redo = (long long *)(p + 256 + skb_setup_bytes(KP) + skb_scratch_bytes(KP, sm_count()));
    long long redo_cap = SK_REDO_CAP;
    // (PILOT_SK_REDO_CAP, tests only: a smaller list capacity, to exercise the overflow scan)
    if (const char *e = getenv("PILOT_SK_REDO_CAP")) {
        const long long v = atoll(e);
        if (v >= 0 && v < redo_cap) redo_cap = v;
    }
    if (precision == PILOT_F32) {
        // single precision: per-problem kernel in registers (sinkhorn_f32.cu); what it cannot carry (NaN/Inf,
        // zero masses) goes to the FP64 reference-form kernel like in the FP64 mode
        rc = skf_launch(props, K, cost, prm, pm, out, iters, absorptions, status, ws.counter_fast, ws.redo,
                        ws.n_redo, st);
        if (rc) return rc;
        return sinkhorn_ref_launch(props, K, cost, prm, pm, ws.redo, ws.n_redo, redo_cap, out, iters, absorptions,
                                   status, ws.counter_slow, st);
    }
    ws.setup = (double *)(p + 256);
    ws.scratch = (double *)(p + 256 + skb_setup_bytes(KP));
    unsigned char *ptail = (unsigned char *)ws.redo + (size_t)SK_REDO_CAP * sizeof(long long);
    SkTail tail;
    tail.rec = (SkTailRec *)ptail;
    tail.uv = (double *)(ptail + sk_tail_rec_bytes());
    tail.n_tail = ws.counter_fast + 3;
    tail.evict_max = 2;  // measured: 0 (no hand-over) 2.23 ms, 2: 1.45 ms, 4: 1.44 ms, 7: 1.56 ms for 10^4 problems at K = 64
    unsigned long long *tail_counter = ws.counter_fast + 4;
    rc = skb_setup(cost, K, prm, ws.setup, st);
    if (rc) return rc;
    // The cost is symmetric in PILOT (a pdist matrix) but the API takes any matrix.  Whether it is is known
    // only on the device (setup kernel), so the symmetric and the general variant of each solver are both
    // enqueued and the one that does not apply returns at once: no host read-back, the call is fully
    // asynchronous on `stream`.
    // persistent grid: one CTA per SM.  With little work (latency-, not throughput-bound) spread it
    // over all SMs and run only as many warps / slot sets per CTA as there are slot-loads of
    // problems: fewer of them share the FP64 tensor pipe, so every iteration returns sooner.
    const int slot_cap = 2;
    // few cell types: one warp per problem with K0 in registers (lowest latency per iteration, and
    // for 17..32 types also the higher throughput).  With <= 16 types half of its lanes idle, so a
    // large batch goes to the DMMA panels instead (measured: 400 K problems, K = 12: 5.6 vs 6.9 ms).
    // The choice depends on the COHORT size only (not on this rank's share or on the window), so that a problem is
    // solved by the same arithmetic however the pair space is partitioned: results are partition-invariant bit for bit.
    const bool warp_form = algo == 0 && K <= swk_max_k() && (K > 16 || (long long)S * S < 200000);
    if (warp_form) {  // symmetric cost only; an asymmetric one falls through to the general panels below
        rc = swk_launch(props, K, prm, pm, ws.setup, out, iters, absorptions, status, ws.counter_fast, ws.redo,
                        ws.n_redo, st);
        if (rc) return rc;
    }
    const long long spw = skb_slots_per_warp();
    long long ctas = (pm.n_local + spw - 1) / spw;
    if (ctas > sm_count()) ctas = sm_count();
    long long warp_cap = (pm.n_local + spw * ctas - 1) / (spw * ctas);
    if (warp_cap > skb_warps()) warp_cap = skb_warps();
    if (warp_cap < 1) warp_cap = 1;
    for (int sym = warp_form ? 0 : 1; sym >= 0; --sym) {
        rc = skb_launch(props, K, prm, pm, ws.setup, ws.scratch, (int)ctas, slot_cap, (int)warp_cap, sym != 0,
                        tail, out, iters, absorptions, status, ws.counter_fast, ws.redo, ws.n_redo, st);
        if (rc) return rc;
    }
    // the stragglers the panels handed over continue in warp form
    for (int sym = warp_form ? 0 : 1; sym >= 0; --sym) {
        rc = skt_launch(props, K, prm, pm, ws.setup, ws.scratch, sym != 0, tail, tail_counter, out, iters,
                        absorptions, status, ws.redo, ws.n_redo, st);
        if (rc) return rc;
    }
    if (rc) return rc;
    // problems the scaled form could not represent (normally none): reference-form kernel
    return sinkhorn_ref_launch(props, K, cost, prm, pm, ws.redo, ws.n_redo, redo_cap, out, iters, absorptions,
                               status, ws.counter_slow, st);
}
