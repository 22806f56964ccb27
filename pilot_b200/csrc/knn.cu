// Row k-nearest-neighbour search on the resident S x S distance matrix -- the first consumer of the hot path:
// pilotpy.pl.trajectory feeds EMD / EMD.max() to pydiffmap's DiffusionMap.from_sklearn(k=64), whose first step is
// sklearn NearestNeighbors(n_neighbors=k).kneighbors_graph(X, mode='distance') with the ROWS of the matrix as
// S-dimensional feature vectors (reference pilotpy/plot/ploting.py:95-110).
//
// The S x S Gram matrix G = X X^T is a plain library GEMM (cuBLAS DGEMM through torch.mm on the host side); this
// kernel does what follows it, one CTA per query row i:
//   d2_j = |x_i|^2 + |x_j|^2 - 2 G_ij (clamped at 0, exactly 0 for j == i, as sklearn's euclidean_distances does)
//   staged once in shared memory (S <= 24 576 doubles = 192 KB; larger S re-reads the Gram row from L2),
//   k-th smallest by an 8-pass radix select on the bit patterns (non-negative doubles order like integers),
//   the k winners (ties at the threshold by ascending index) sorted by (distance, index) with a bitonic network,
//   written as int32 indices + sqrt distances.
// HBM-bound: one read of the Gram row per query (8 S^2 bytes in total).
#include "common.cuh"

namespace pilot {

constexpr int KNN_THREADS = 1024;  // one CTA per SM (the staged row fills shared memory): all 32 warps of it
constexpr int KNN_MAXK = 1024;

__global__ void knn_sqnorm_kernel(const double *__restrict__ G, int S, double *__restrict__ sq)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) sq[i] = G[(size_t)i * S + i];
}

template <bool STAGED>
__global__ void __launch_bounds__(KNN_THREADS)
knn_rows_kernel(const double *__restrict__ G, const double *__restrict__ sq, int S, int k, int kp2,
                int *__restrict__ idx_out, double *__restrict__ dist_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(smem_raw);         // [kp2] sort keys (d2 bits)
    int *sidx = reinterpret_cast<int *>(skey + kp2);                                      // [kp2]
    double *srow = reinterpret_cast<double *>(sidx + kp2 + (kp2 & 1));                    // [S] when STAGED
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_rank, s_nlt, s_neq;
    const int i = blockIdx.x;
    const double ni = sq[i];
    const double *grow = G + (size_t)i * S;
    auto d2 = [&](int j) -> double {
        if (j == i) return 0.0;
        const double v = (ni + sq[j]) - 2.0 * grow[j];
        return v > 0.0 ? v : 0.0;
    };
    if (STAGED) {
        constexpr int UB = 8;  // independent loads in flight per thread
        for (int j0 = threadIdx.x; j0 < S; j0 += UB * KNN_THREADS) {
            double g[UB], nj[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int j = j0 + u * KNN_THREADS;
                g[u] = j < S ? grow[j] : 0.0;
                nj[u] = j < S ? sq[j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int j = j0 + u * KNN_THREADS;
                if (j >= S) continue;
                const double v = (ni + nj[u]) - 2.0 * g[u];
                srow[j] = j == i ? 0.0 : (v > 0.0 ? v : 0.0);
            }
        }
        __syncthreads();
    }
    auto val = [&](int j) -> double { return STAGED ? srow[j] : d2(j); };

    // ---- radix select of the (k-1)-th smallest (0-based): 8 passes of 8 bits, most significant first ----
    unsigned long long prefix = 0ULL;
    unsigned int rank = (unsigned)(k - 1);
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0u;
        __syncthreads();
        const unsigned long long himask = shift == 56 ? 0ULL : (~0ULL << (shift + 8));
        for (int j = threadIdx.x; j < S; j += blockDim.x) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(val(j));
            if ((key & himask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int run = 0, r = rank;
            int digit = 255;
            for (int b = 0; b < 256; ++b) {
                if (r < run + hist[b]) { digit = b; break; }
                run += hist[b];
            }
            s_prefix = prefix | ((unsigned long long)digit << shift);
            s_rank = r - run;
        }
        __syncthreads();
        prefix = s_prefix;
        rank = s_rank;
        __syncthreads();
    }
    const unsigned long long tkey = prefix;  // bit pattern of the k-th smallest squared distance

    // ---- collect: everything below the threshold, then ties at the threshold by ascending index ----
    if (threadIdx.x == 0) { s_nlt = 0u; s_neq = 0u; }
    for (int t = threadIdx.x; t < kp2; t += blockDim.x) { skey[t] = ~0ULL; sidx[t] = 0x7fffffff; }
    __syncthreads();
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(val(j));
        if (key < tkey) {
            const unsigned p = atomicAdd(&s_nlt, 1u);
            skey[p] = key;
            sidx[p] = j;
        }
    }
    __syncthreads();
    const unsigned nlt = s_nlt, need = (unsigned)k - nlt;  // >= 1 ties are needed
    // ties: ranks by index (a CTA-wide ordered pass in chunks of blockDim)
    for (int base = 0; base < S && s_neq < need; base += blockDim.x) {
        const int j = base + threadIdx.x;
        const bool eq = j < S && (unsigned long long)__double_as_longlong(val(j)) == tkey;
        // ordered compaction inside the chunk: warp ballots + per-warp offsets
        __shared__ unsigned int woff[KNN_THREADS / 32 + 1];
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) woff[warp + 1] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            woff[0] = s_neq;
            for (int w = 0; w < KNN_THREADS / 32; ++w) woff[w + 1] += woff[w];
        }
        __syncthreads();
        if (eq) {
            const unsigned p = woff[warp] + __popc(bal & ((1u << lane) - 1u));
            if (p < need) { skey[nlt + p] = tkey; sidx[nlt + p] = j; }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_neq = woff[KNN_THREADS / 32];
        __syncthreads();
    }

    // ---- sort the k winners by (d2, index): bitonic network over kp2 slots (padding sorts last) ----
    for (int size = 2; size <= kp2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < kp2; t += blockDim.x) {
                const int p = t ^ stride;
                if (p > t) {
                    const unsigned long long ka = skey[t], kb = skey[p];
                    const int ia = sidx[t], ib = sidx[p];
                    const bool gt = ka > kb || (ka == kb && ia > ib);
                    const bool up = (t & size) == 0;
                    if (gt == up) { skey[t] = kb; skey[p] = ka; sidx[t] = ib; sidx[p] = ia; }
                }
            }
            __syncthreads();
        }
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        idx_out[(size_t)i * k + t] = sidx[t];
        dist_out[(size_t)i * k + t] = sqrt(__longlong_as_double((long long)skey[t]));
    }
}

}  // namespace pilot

extern "C" int pilot_knn_rows(const double *gram, int S, int k, int32_t *idx, double *dist, void *workspace,
                              size_t workspace_bytes, void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(gram && idx && dist && workspace, "pilot_knn_rows: NULL pointer");
    PILOT_CHECK_ARG(S >= 1 && k >= 1 && k <= S && k <= KNN_MAXK, "pilot_knn_rows: S=%d k=%d (1 <= k <= min(S, %d))", S, k,
                    KNN_MAXK);
    PILOT_CHECK_ARG(workspace_bytes >= (size_t)S * sizeof(double), "pilot_knn_rows: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double *sq = (double *)workspace;
    knn_sqnorm_kernel<<<(S + 255) / 256, 256, 0, st>>>(gram, S, sq);
    PILOT_LAUNCH_CHECK();
    int kp2 = 2;
    while (kp2 < k) kp2 <<= 1;
    const size_t base = (size_t)kp2 * 12 + 8;
    const size_t staged_bytes = base + (size_t)S * sizeof(double);
    if (staged_bytes <= 200 * 1024) {
        PILOT_CUDA(cudaFuncSetAttribute(knn_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)staged_bytes));
        knn_rows_kernel<true><<<S, KNN_THREADS, staged_bytes, st>>>(gram, sq, S, k, kp2, idx, dist);
    } else {
        knn_rows_kernel<false><<<S, KNN_THREADS, base, st>>>(gram, sq, S, k, kp2, idx, dist);
    }
    PILOT_LAUNCH_CHECK();
    return 0;
}
