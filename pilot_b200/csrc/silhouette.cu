// Silhouette coefficients of the samples on the resident S x S matrix -- the quality score of the second consumer
// of the hot path: Sil_computing(EMD, labels, metric='cosine') =
// sklearn.metrics.silhouette_score(EMD, labels, metric='cosine') (reference pilotpy/tools/Trajectory.py:593-612,
// called from wasserstein_distance :107-113, and once per resolution from the Leiden sweeps
// pilotpy/plot/ploting.py:310-324, :420-439), with the ROWS of the matrix as S-dimensional points.
//
// scikit-learn materialises the S x S pairwise distance matrix and reduces it per (sample, cluster).  Here the Gram
// matrix G = X X^T is one library DGEMM on the caller's side (as for pilot_knn_rows); this kernel does the rest in
// ONE pass over G, one CTA per sample i:
//   d_ij   = clip(1 - G_ij / (|x_i| |x_j|), 0, 2)          (cosine; sklearn's cosine_distances, d_ii = 0)
//          = sqrt(max(G_ii + G_jj - 2 G_ij, 0))             (euclidean)
//          = M_ij                                           (precomputed: the matrix IS the distance)
//   staged in shared memory (S <= 24 576), then one warp per cluster sums the cluster's members -- the samples
//   arrive sorted by label (perm / seg from the caller), lanes stride the segment, fixed shuffle tree: the sums
//   are deterministic --, a_i = intra mean over the n_c - 1 others, b_i = smallest mean to another cluster,
//   s_i = (b_i - a_i) / max(a_i, b_i), 0 for singleton clusters (sklearn's convention).
// HBM-bound: 8 S^2 bytes read once.
#include "common.cuh"

namespace pilot {

constexpr int SIL_THREADS = 1024;  // one CTA per SM (the staged row fills shared memory): all 32 warps of it
constexpr int SIL_MAXL = 4096;

__global__ void sil_prep_kernel(const double *__restrict__ M, int S, int metric, double *__restrict__ nrm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const double g = M[(size_t)i * S + i];
    // cosine: sklearn normalises a zero row to itself (norm 0 -> divide by 1)
    nrm[i] = metric == PILOT_METRIC_COSINE ? (g > 0.0 ? 1.0 / sqrt(g) : 1.0) : g;
}

template <bool STAGED>
__global__ void __launch_bounds__(SIL_THREADS)
sil_rows_kernel(const double *__restrict__ M, const double *__restrict__ nrm, int S, int metric,
                const int *__restrict__ perm, const int *__restrict__ seg, const int *__restrict__ label, int L,
                double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ssum = reinterpret_cast<double *>(smem_raw);  // [L]
    double *srow = ssum + L;                              // [S] when STAGED
    const int i = blockIdx.x;
    const double *row = M + (size_t)i * S;
    const double ni = metric == PILOT_SIL_PRECOMPUTED ? 0.0 : nrm[i];
    auto dist = [&](int j) -> double {
        if (metric == PILOT_SIL_PRECOMPUTED) return row[j];
        if (j == i) return 0.0;
        if (metric == PILOT_METRIC_COSINE) {
            const double v = 1.0 - row[j] * ni * nrm[j];
            return v < 0.0 ? 0.0 : (v > 2.0 ? 2.0 : v);
        }
        const double v = (ni + nrm[j]) - 2.0 * row[j];
        return v > 0.0 ? sqrt(v) : 0.0;
    };
    constexpr int UB = 8;  // independent loads in flight per thread: both loops are chains of L2 round trips otherwise
    if (STAGED) {
        for (int j0 = threadIdx.x; j0 < S; j0 += UB * SIL_THREADS) {
            double g[UB], nj[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int j = j0 + u * SIL_THREADS;
                g[u] = j < S ? row[j] : 0.0;
                nj[u] = (j < S && metric != PILOT_SIL_PRECOMPUTED) ? nrm[j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int j = j0 + u * SIL_THREADS;
                if (j >= S) continue;
                double v;
                if (metric == PILOT_SIL_PRECOMPUTED) v = g[u];
                else if (j == i) v = 0.0;
                else if (metric == PILOT_METRIC_COSINE) {
                    v = 1.0 - g[u] * ni * nj[u];
                    v = v < 0.0 ? 0.0 : (v > 2.0 ? 2.0 : v);
                } else {
                    v = (ni + nj[u]) - 2.0 * g[u];
                    v = v > 0.0 ? sqrt(v) : 0.0;
                }
                srow[j] = v;
            }
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = warp; c < L; c += SIL_THREADS / 32) {
        double s = 0.0;
        const int p1 = seg[c + 1];
        if (STAGED) {
            // lane-strided partial sums in a fixed order (batches of UB members: the perm loads are independent)
            for (int p0 = seg[c] + lane; p0 < p1; p0 += 32 * UB) {
                int q[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) q[u] = p0 + 32 * u < p1 ? perm[p0 + 32 * u] : -1;
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (q[u] >= 0) s += srow[q[u]];
            }
        } else {
            for (int p = seg[c] + lane; p < p1; p += 32) s += dist(perm[p]);
        }
        s = warp_sum_d(s);
        if (lane == 0) ssum[c] = s;
    }
    __syncthreads();
    if (warp == 0) {
        const int ci = label[i];
        const int nci = seg[ci + 1] - seg[ci];
        double b = __longlong_as_double(0x7ff0000000000000LL);
        for (int c = lane; c < L; c += 32) {
            const int nc = seg[c + 1] - seg[c];
            if (c != ci && nc > 0) b = fmin(b, ssum[c] / (double)nc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = fmin(b, __shfl_xor_sync(0xffffffffu, b, o));
        if (lane == 0) {
            double s = 0.0;
            if (nci > 1) {
                const double a = ssum[ci] / (double)(nci - 1);
                const double m = fmax(a, b);
                s = m > 0.0 ? (b - a) / m : 0.0;  // sklearn: nan_to_num(0 / 0) = 0
            }
            out[i] = s;
        }
    }
}

}  // namespace pilot

extern "C" int pilot_silhouette_rows(const double *matrix, int S, int metric, const int32_t *perm, const int32_t *seg,
                                     const int32_t *label, int n_labels, double *sil, void *workspace,
                                     size_t workspace_bytes, void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(matrix && perm && seg && label && sil && workspace, "pilot_silhouette_rows: NULL pointer");
    PILOT_CHECK_ARG(metric == PILOT_SIL_PRECOMPUTED || metric == PILOT_METRIC_COSINE || metric == PILOT_METRIC_EUCLIDEAN,
                    "pilot_silhouette_rows: metric %d (cosine, euclidean or precomputed)", metric);
    PILOT_CHECK_ARG(S >= 2 && n_labels >= 2 && n_labels <= S - 1 && n_labels <= SIL_MAXL,
                    "pilot_silhouette_rows: S=%d n_labels=%d (2 <= n_labels <= min(S - 1, %d))", S, n_labels, SIL_MAXL);
    PILOT_CHECK_ARG(workspace_bytes >= (size_t)S * sizeof(double), "pilot_silhouette_rows: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double *nrm = (double *)workspace;
    if (metric != PILOT_SIL_PRECOMPUTED) {
        sil_prep_kernel<<<(S + 255) / 256, 256, 0, st>>>(matrix, S, metric, nrm);
        PILOT_LAUNCH_CHECK();
    }
    const size_t base = (size_t)n_labels * sizeof(double);
    const size_t staged = base + (size_t)S * sizeof(double);
    if (staged <= 200 * 1024) {
        PILOT_CUDA(cudaFuncSetAttribute(sil_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
        sil_rows_kernel<true><<<S, SIL_THREADS, staged, st>>>(matrix, nrm, S, metric, perm, seg, label, n_labels, sil);
    } else {
        sil_rows_kernel<false><<<S, SIL_THREADS, base, st>>>(matrix, nrm, S, metric, perm, seg, label, n_labels, sil);
    }
    PILOT_LAUNCH_CHECK();
    return 0;
}
