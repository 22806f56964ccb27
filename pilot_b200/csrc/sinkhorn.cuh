// Shared declarations of the Sinkhorn kernels (sinkhorn_ref.cu, sinkhorn_batched.cu, sinkhorn_ws.cu, sinkhorn_warp.cu).
#pragma once
#include "common.cuh"

namespace pilot {

struct SkParams {
    double reg, stop_thr, tau;
    int num_iter_max, check_every;
};

constexpr long long SK_REDO_CAP = 1LL << 20;

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
int skb_setup(const double *cost, int K, const SkParams &prm, double *setup, bool *symmetric, cudaStream_t st);
int skb_launch(const double *props, int K, const SkParams &prm, const PairMap &pm, double *setup, double *scratch,
               int ctas, int slot_cap, int warp_cap, bool symmetric, double *out, int *iters, int *absn, int *status,
               unsigned long long *counter, long long *redo, unsigned long long *n_redo, cudaStream_t st);

// warp-specialised variant (sinkhorn_ws.cu)
size_t skw_smem_bytes(int KP);
size_t skw_scratch_bytes(int KP, int ctas);
int skw_slots_per_set();
int skw_sets();
int skw_launch(const double *props, int K, const SkParams &prm, const PairMap &pm, double *setup, double *scratch,
               int ctas, int slot_cap, int set_cap, bool symmetric, double *out, int *iters, int *absn, int *status,
               unsigned long long *counter, long long *redo, unsigned long long *n_redo, cudaStream_t st);

// one warp per problem, K0 in registers (sinkhorn_warp.cu): K <= swk_max_k(), symmetric cost
int swk_max_k();
int swk_launch(const double *props, int K, const SkParams &prm, const PairMap &pm, const double *setup, double *out,
               int *iters, int *absn, int *status, unsigned long long *counter, long long *redo,
               unsigned long long *n_redo, cudaStream_t st);

}  // namespace pilot
