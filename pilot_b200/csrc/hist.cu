// Kernel (1): per-(sample, cell-type) counting + first-appearance index, and the
// Dirichlet-smoothed proportions.  Replaces the pandas unique()/value_counts()/
// boolean-mask scans of Cluster_Representations (reference
// pilotpy/tools/Trajectory.py:402-430).
//
// HBM-bound integer work: 8 B read per cell (two int32 codes), S*K*8 B written.
// Layout: persistent grid (SM count x resident CTAs), 128-bit vectorised loads,
// warp-aggregated (match.any) atomics into a per-CTA shared-memory histogram
// when S*K fits, else straight into L2.
#include "common.cuh"

namespace pilot {

constexpr int HIST_THREADS = 512;

__device__ __forceinline__ void first_min(unsigned long long *slot, unsigned long long idx)
{
    // cheap guard: after the first few thousand cells nothing passes it
    if (idx < *reinterpret_cast<volatile unsigned long long *>(slot)) atomicMin(slot, idx);
}

template <bool SMEM_COUNTS>
__global__ void __launch_bounds__(HIST_THREADS)
hist_kernel(const int *__restrict__ ct, const int *__restrict__ smp, long long n, int K, int S,
            unsigned long long *__restrict__ counts, unsigned long long *__restrict__ first_ct,
            unsigned long long *__restrict__ first_smp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *s_fct = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *s_fsm = s_fct + K;
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(s_fsm + S);
    const int SK = S * K;
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_fct[i] = (unsigned long long)n;
    for (int i = threadIdx.x; i < S; i += blockDim.x) s_fsm[i] = (unsigned long long)n;
    if (SMEM_COUNTS)
        for (int i = threadIdx.x; i < SK; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();

    const long long nvec = n >> 2;  // groups of 4 cells
    const int4 *ct4 = reinterpret_cast<const int4 *>(ct);
    const int4 *sm4 = reinterpret_cast<const int4 *>(smp);
    // every lane of a warp runs the same number of iterations (match.any needs the full warp)
    const long long warp_stride = (long long)gridDim.x * (blockDim.x >> 5) * 32;
    const long long warp_first = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    const int lane = threadIdx.x & 31;
    for (long long base = warp_first; base < nvec; base += warp_stride) {
        const long long v = base + lane;
        const bool valid = v < nvec;
        int4 c = make_int4(0, 0, 0, 0), s = make_int4(0, 0, 0, 0);
        if (valid) {
            c = __ldg(ct4 + v);
            s = __ldg(sm4 + v);
        }
        const int cc[4] = {c.x, c.y, c.z, c.w};
        const int ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // out-of-range codes are dropped; the host detects them as sum(counts) != n_cells
            const bool ok = valid && (unsigned)cc[q] < (unsigned)K && (unsigned)ss[q] < (unsigned)S;
            if (ok) {
                const int key = ss[q] * K + cc[q];
                const unsigned long long idx = (unsigned long long)(v * 4 + q);
                first_min(&s_fct[cc[q]], idx);
                first_min(&s_fsm[ss[q]], idx);
                if (SMEM_COUNTS) {
                    // ATOMS.POPC.INC: the hardware merges the lanes of the warp that hit the same counter
                    atomicAdd(&s_cnt[key], 1u);
                } else {
                    // L2 atomics: merge equal keys of the warp first (match.any), one atomic per group
                    const unsigned grp = __match_any_sync(__activemask(), key);
                    if (lane == __ffs(grp) - 1) atomicAdd(&counts[key], (unsigned long long)__popc(grp));
                }
            }
        }
    }
    // scalar tail (n % 4 cells), first CTA only
    if (blockIdx.x == 0) {
        for (long long i = (nvec << 2) + threadIdx.x; i < n; i += blockDim.x) {
            const int k = ct[i], s = smp[i];
            if ((unsigned)k < (unsigned)K && (unsigned)s < (unsigned)S) {
                first_min(&s_fct[k], (unsigned long long)i);
                first_min(&s_fsm[s], (unsigned long long)i);
                if (SMEM_COUNTS) atomicAdd(&s_cnt[s * K + k], 1u);
                else atomicAdd(&counts[s * K + k], 1ULL);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x)
        if (s_fct[i] < (unsigned long long)n) atomicMin(&first_ct[i], s_fct[i]);
    for (int i = threadIdx.x; i < S; i += blockDim.x)
        if (s_fsm[i] < (unsigned long long)n) atomicMin(&first_smp[i], s_fsm[i]);
    if (SMEM_COUNTS)
        for (int i = threadIdx.x; i < SK; i += blockDim.x) {
            const unsigned c = s_cnt[i];
            if (c) atomicAdd(&counts[i], (unsigned long long)c);
        }
}

__global__ void hist_init_kernel(unsigned long long *counts, long long sk, unsigned long long *first_ct,
                                 int K, unsigned long long *first_smp, int S, unsigned long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < sk) counts[i] = 0ULL;
    if (i < K) first_ct[i] = n;
    if (i < S) first_smp[i] = n;
}

// prior_k = n_k/(N-1)*regulizer and its sequential sum (Trajectory.py:405-409, :430).  One CTA of
// K x R threads: thread (r, k) adds the rows s = r (mod R) of column k (integers: any order is
// exact), a shared-memory pass folds the R partials, thread 0 takes the left-to-right FP64 sum.
__global__ void props_prior_kernel(const long long *__restrict__ counts_raw, int K_raw,
                                   const int *__restrict__ perm_s, const int *__restrict__ perm_k,
                                   int K, int S, int R, long long n_cells, double regulizer,
                                   double *__restrict__ prior)
{
    extern __shared__ unsigned long long s_part[];  // R x K
    const int k = threadIdx.x % K, r = threadIdx.x / K;
    if (r < R) {
        const int kr = perm_k ? perm_k[k] : k;
        unsigned long long acc = 0;
        for (int s = r; s < S; s += R) acc += (unsigned long long)counts_raw[(long long)(perm_s ? perm_s[s] : s) * K_raw + kr];
        s_part[r * K + k] = acc;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        unsigned long long nk = 0;
        for (int q = 0; q < R; ++q) nk += s_part[q * K + threadIdx.x];
        prior[threadIdx.x] = __dmul_rn(__ddiv_rn((double)nk, (double)(n_cells - 1)), regulizer);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sp = 0.0;
        for (int q = 0; q < K; ++q) sp = __dadd_rn(sp, prior[q]);
        prior[K] = sp;
    }
}

// prior_k = n_k/(N-1)*regulizer (Trajectory.py:405-409); every FP sum is a sequential
// left-to-right FP64 sum like Python's builtin sum() (:428-430); explicit
// round-to-nearest intrinsics so nvcc cannot contract into FMA.  One thread per sample row.
__global__ void props_finalize_kernel(const long long *__restrict__ counts_raw, int K_raw,
                                      const int *__restrict__ perm_s, const int *__restrict__ perm_k,
                                      int K, int S, int normalization, const double *__restrict__ prior,
                                      double *__restrict__ props, long long *__restrict__ counts_out)
{
    extern __shared__ double s_prior[];  // K + 1
    for (int k = threadIdx.x; k <= K; k += blockDim.x) s_prior[k] = prior[k];
    __syncthreads();
    const double sp = s_prior[K];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const long long *row = counts_raw + (long long)(perm_s ? perm_s[s] : s) * K_raw;
    double *out = props + (long long)s * K;
    double sc = 0.0;
    for (int k = 0; k < K; ++k) {
        const long long c = row[perm_k ? perm_k[k] : k];
        if (counts_out) counts_out[(long long)s * K + k] = c;
        sc = __dadd_rn(sc, (double)c);
        out[k] = (double)c;
    }
    if (normalization) {
        const double den = __dadd_rn(sc, sp);
        for (int k = 0; k < K; ++k) out[k] = __ddiv_rn(__dadd_rn(out[k], s_prior[k]), den);
    }
}

}  // namespace pilot

extern "C" int pilot_hist(const int32_t *ct_code, const int32_t *smp_code, int64_t n_cells, int K, int S,
                          int64_t *counts, int64_t *first_ct, int64_t *first_smp, void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(K >= 1 && S >= 1, "pilot_hist: K=%d S=%d must be >= 1", K, S);
    PILOT_CHECK_ARG(n_cells >= 0, "pilot_hist: n_cells=%lld", (long long)n_cells);
    PILOT_CHECK_ARG((long long)K * S < (1LL << 31), "pilot_hist: S*K too large");
    PILOT_CHECK_ARG(counts && first_ct && first_smp, "pilot_hist: NULL output");
    PILOT_CHECK_ARG(n_cells == 0 || (ct_code && smp_code), "pilot_hist: NULL input");
    PILOT_CHECK_ARG((((uintptr_t)ct_code | (uintptr_t)smp_code) & 15) == 0,
                    "pilot_hist: code arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const long long sk = (long long)S * K;
    {
        long long m = sk > K ? sk : K;
        if (S > m) m = S;
        const int th = 256;
        hist_init_kernel<<<(unsigned)((m + th - 1) / th), th, 0, st>>>(
            (unsigned long long *)counts, sk, (unsigned long long *)first_ct, K,
            (unsigned long long *)first_smp, S, (unsigned long long)n_cells);
        PILOT_LAUNCH_CHECK();
    }
    if (n_cells == 0) return 0;
    const size_t first_bytes = (size_t)(K + S) * sizeof(unsigned long long);
    const size_t smem_counts = first_bytes + (size_t)sk * sizeof(unsigned int);
    const bool use_smem = smem_counts <= 200 * 1024;  // up to 110 KB two CTAs per SM stay resident, beyond that one
    const size_t smem = use_smem ? smem_counts : first_bytes;
    PILOT_CHECK_ARG(smem <= 200 * 1024, "pilot_hist: K+S=%d too large for shared first-index table", K + S);
    const long long warps_needed = ((n_cells >> 2) + 31) / 32;
    long long ctas = (warps_needed + (HIST_THREADS / 32) - 1) / (HIST_THREADS / 32);
    const long long max_ctas = (long long)sm_count() * (use_smem && smem > 110 * 1024 ? 1 : (use_smem && smem > 48 * 1024 ? 2 : 4));
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    if (use_smem) {
        PILOT_CUDA(cudaFuncSetAttribute(hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hist_kernel<true><<<(unsigned)ctas, HIST_THREADS, smem, st>>>(
            ct_code, smp_code, n_cells, K, S, (unsigned long long *)counts,
            (unsigned long long *)first_ct, (unsigned long long *)first_smp);
    } else {
        PILOT_CUDA(cudaFuncSetAttribute(hist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hist_kernel<false><<<(unsigned)ctas, HIST_THREADS, smem, st>>>(
            ct_code, smp_code, n_cells, K, S, (unsigned long long *)counts,
            (unsigned long long *)first_ct, (unsigned long long *)first_smp);
    }
    PILOT_LAUNCH_CHECK();
    return 0;
}

extern "C" int pilot_props_finalize(const int64_t *counts_raw, int K_raw, int S_raw, const int32_t *perm_k,
                                    const int32_t *perm_s, int K, int S, int64_t n_cells, double regulizer,
                                    int normalization, double *props, int64_t *counts_out, double *prior_out,
                                    void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(K >= 1 && S >= 1 && counts_raw && props && prior_out, "pilot_props_finalize: bad argument");
    PILOT_CHECK_ARG(K <= K_raw && S <= S_raw, "pilot_props_finalize: K=%d S=%d exceed raw dims %d %d", K, S, K_raw, S_raw);
    PILOT_CHECK_ARG(K <= 1024, "pilot_props_finalize: K=%d too large (max 1024)", K);
    cudaStream_t st = (cudaStream_t)stream;
    int R = 1024 / K;
    if (R > S) R = S;
    if (R < 1) R = 1;
    props_prior_kernel<<<1, K * R, (size_t)R * K * sizeof(unsigned long long), st>>>(
        (const long long *)counts_raw, K_raw, perm_s, perm_k, K, S, R, n_cells, regulizer, prior_out);
    PILOT_LAUNCH_CHECK();
    const int th = 128;
    props_finalize_kernel<<<(S + th - 1) / th, th, (K + 1) * sizeof(double), st>>>(
        (const long long *)counts_raw, K_raw, perm_s, perm_k, K, S, normalization, prior_out, props,
        (long long *)counts_out);
    PILOT_LAUNCH_CHECK();
    return 0;
}
