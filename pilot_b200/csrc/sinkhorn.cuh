// Shared declarations of the Sinkhorn kernels (sinkhorn_ref.cu, sinkhorn_batched.cu, sinkhorn_tail.cu, sinkhorn_warp.cu).
#pragma once
#include "common.cuh"

namespace pilot {

struct SkParams {
    double reg, stop_thr, tau;
    int num_iter_max, check_every;
};

constexpr long long SK_REDO_CAP = 1LL << 20;
// out[] of a problem queued for the reference-form kernel: a NaN with this payload.  When more than
// SK_REDO_CAP problems are queued the list overflows and the reference-form kernel finds the rest by
// scanning out[] for the marker (sinkhorn_ref.cu, mode 2), so no problem is ever left unsolved.
constexpr long long SK_REDO_MARK = 0x7ff8b200b200b200LL;

// hand-over of straggler problems from the DMMA-panel kernel to the warp-form tail kernel
struct SkTailRec {
    long long prob;  // local problem index
    int ii, nabs, hasabs;
    int sslot;       // index of the problem's rea / reb block in the panel kernel's scratch
};
struct SkTail {
    SkTailRec *rec;  // nullptr: no hand-over
    double *uv;      // [slot][ut | vt][KP]
    unsigned long long *n_tail;
    int evict_max;   // a warp hands its problems over when the pool is empty and <= this many are left
};

size_t sinkhorn_ref_smem(int K);
int sinkhorn_ref_launch(const double *props, int K, const double *cost, const SkParams &prm, const PairMap &pm,
                        const long long *list, const unsigned long long *n_list_dev, long long max_list,
                        double *out, int *iters, int *absorptions, int *status, unsigned long long *counter,
                        cudaStream_t st);

int skb_pad(int K);
size_t skb_setup_bytes(int KP);
size_t skb_smem_bytes(int KP);
size_t skb_scratch_bytes(int KP, int ctas);
int skb_slots_per_cta();
int skb_slots_per_warp();
int skb_warps();
int skb_setup(const double *cost, int K, const SkParams &prm, double *setup, cudaStream_t st);
int skb_launch(const double *props, int K, const SkParams &prm, const PairMap &pm, double *setup, double *scratch,
               int ctas, int slot_cap, int warp_cap, bool symmetric, const SkTail &tail, double *out, int *iters,
               int *absn, int *status, unsigned long long *counter, long long *redo, unsigned long long *n_redo,
               cudaStream_t st);

// warp-form continuation of the handed-over problems (sinkhorn_tail.cu)
int skt_launch(const double *props, int K, const SkParams &prm, const PairMap &pm, const double *setup,
               const double *scratch, bool symmetric, const SkTail &tail, unsigned long long *tail_counter,
               double *out, int *iters, int *absn, int *status, long long *redo, unsigned long long *n_redo,
               cudaStream_t st);

// one warp per problem, K0 in registers (sinkhorn_warp.cu): K <= swk_max_k(), symmetric cost
int swk_max_k();
int swk_launch(const double *props, int K, const SkParams &prm, const PairMap &pm, const double *setup, double *out,
               int *iters, int *absn, int *status, unsigned long long *counter, long long *redo,
               unsigned long long *n_redo, cudaStream_t st);

// FP32 mode: one warp per problem, per-problem kernel in registers (sinkhorn_f32.cu), K <= 64
int skf_launch(const double *props, int K, const double *cost, const SkParams &prm, const PairMap &pm, double *out,
               int *iters, int *absn, int *status, unsigned long long *counter, long long *redo,
               unsigned long long *n_redo, cudaStream_t st);

}  // namespace pilot
