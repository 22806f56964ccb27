// Kernel (2a): per-(cell type, dimension) MEDIAN of the embedding rows, in the
// input dtype.  Replaces data[annot.cell_type == k].median(axis=0)
// (reference pilotpy/tools/Trajectory.py:465-466; pandas nanmedian semantics:
// NaNs ignored, even counts -> (lo + hi) / 2 rounded in the input dtype).
//
// Exact selection by most-significant-digit radix select on the order-preserving
// integer image of the floats: every pass streams X once (coalesced, row-major),
// histograms the current 8-bit digit of the elements whose higher digits match
// the running prefix of their (type, dim) query, then a one-warp-per-query scan
// picks the digit that contains the wanted rank.  Two queries per (type, dim)
// (ranks (n-1)/2 and n/2) so even counts need no second selection.
//
// Algorithmic bytes: one read of X + codes; this v1 reads X
// once per digit pass (4 for f32, 8 for f64) -- see DESIGN.md for the roofline
// accounting and the planned single-pass variant.
#include "common.cuh"

namespace pilot {

template <typename T> struct KeyOf;
template <> struct KeyOf<float> {
    using type = unsigned int;
    static constexpr int PASSES = 4;
    __device__ static unsigned int key(float x)
    {
        unsigned int u = __float_as_uint(x);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
    __device__ static float value(unsigned int k)
    {
        unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
        return __uint_as_float(u);
    }
};
template <> struct KeyOf<double> {
    using type = unsigned long long;
    static constexpr int PASSES = 8;
    __device__ static unsigned long long key(double x)
    {
        unsigned long long u = (unsigned long long)__double_as_longlong(x);
        return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
    }
    __device__ static double value(unsigned long long k)
    {
        unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
        return __longlong_as_double((long long)u);
    }
};

// workspace layout (all 8-byte aligned):
//   hist   : KD * 2 * 256 u32
//   prefix : KD * 2 key (stored as u64)
//   rank   : KD * 2 u64
//   nvalid : KD u64
struct MedianWs {
    unsigned int *hist;
    unsigned long long *prefix, *rank, *nvalid;
};

static size_t median_ws_bytes_impl(int K, int D)
{
    size_t kd = (size_t)K * D;
    return kd * 2 * 256 * sizeof(unsigned int) + kd * 2 * 8 + kd * 2 * 8 + kd * 8;
}

template <typename T, int PASS>
__global__ void __launch_bounds__(256)
median_hist_kernel(const T *__restrict__ X, long long n, int D, long long ldx, const int *__restrict__ code,
                   int K, MedianWs ws)
{
    using KO = KeyOf<T>;
    using Key = typename KO::type;
    constexpr int BITS = sizeof(Key) * 8;
    constexpr int SHIFT = BITS - 8 * (PASS + 1);
    const long long total = n * D;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long r = e / D;
        const int d = (int)(e - r * D);
        const int k = __ldg(code + r);
        if ((unsigned)k >= (unsigned)K) continue;
        const T x = X[r * ldx + d];
        if (x != x) continue;  // nanmedian ignores NaN
        const Key key = KO::key(x);
        const int kd = k * D + d;
        const unsigned digit = (unsigned)(key >> SHIFT) & 0xffu;
        if constexpr (PASS == 0) {
            atomicAdd(&ws.hist[((size_t)kd * 2) * 256 + digit], 1u);
        } else {
            const Key hi = (Key)(key >> (SHIFT + 8));
            const Key p0 = (Key)ws.prefix[kd * 2] >> (SHIFT + 8);
            const Key p1 = (Key)ws.prefix[kd * 2 + 1] >> (SHIFT + 8);
            if (hi == p0) atomicAdd(&ws.hist[((size_t)kd * 2) * 256 + digit], 1u);
            if (hi == p1) atomicAdd(&ws.hist[((size_t)kd * 2 + 1) * 256 + digit], 1u);
        }
    }
}

// one warp per (kd, q)
template <typename T, int PASS>
__global__ void median_scan_kernel(int KD, MedianWs ws)
{
    using Key = typename KeyOf<T>::type;
    constexpr int BITS = sizeof(Key) * 8;
    constexpr int SHIFT = BITS - 8 * (PASS + 1);
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= KD * 2) return;
    const int kd = w >> 1, q = w & 1;
    unsigned int *h = ws.hist + ((size_t)kd * 2 + (PASS == 0 ? 0 : q)) * 256;
    unsigned int c[8];
    unsigned int mine = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        c[t] = h[lane * 8 + t];
        mine += c[t];
    }
    // inclusive warp scan of per-lane totals
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long rank;
    if (PASS == 0) {
        if (lane == 0 && q == 0) ws.nvalid[kd] = total;
        rank = total == 0 ? 0ULL : (q == 0 ? (unsigned long long)(total - 1) / 2 : (unsigned long long)total / 2);
    } else {
        rank = ws.rank[kd * 2 + q];
    }
    const unsigned int excl = incl - mine;
    const bool here = total > 0 && rank >= excl && rank < incl;
    if (here) {
        unsigned long long run = excl, below = 0;
        int digit = -1;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (digit < 0 && rank < run + c[t]) {
                digit = lane * 8 + t;
                below = run;
            }
            run += c[t];
        }
        Key p = PASS == 0 ? (Key)0 : (Key)ws.prefix[kd * 2 + q];
        p |= (Key)digit << SHIFT;
        ws.prefix[kd * 2 + q] = (unsigned long long)p;
        ws.rank[kd * 2 + q] = rank - below;
    }
    if (total == 0 && lane == 0) {
        ws.prefix[kd * 2 + q] = 0ULL;
        ws.rank[kd * 2 + q] = 0ULL;
    }
}

template <typename T>
__global__ void median_final_kernel(int KD, MedianWs ws, T *__restrict__ cent, double *__restrict__ cent64)
{
    using KO = KeyOf<T>;
    using Key = typename KO::type;
    const int kd = blockIdx.x * blockDim.x + threadIdx.x;
    if (kd >= KD) return;
    T med;
    if (ws.nvalid[kd] == 0) {
        med = (T)NAN;
    } else {
        const T lo = KO::value((Key)ws.prefix[kd * 2]);
        const T hi = KO::value((Key)ws.prefix[kd * 2 + 1]);
        if (lo == hi) med = lo;
        else if (sizeof(T) == 4) med = (T)__fdiv_rn(__fadd_rn((float)lo, (float)hi), 2.0f);
        else med = (T)__ddiv_rn(__dadd_rn((double)lo, (double)hi), 2.0);
    }
    cent[kd] = med;
    cent64[kd] = (double)med;
}

template <typename T, int PASS>
static int median_pass(const T *X, long long n, int D, long long ldx, const int *code, int K, MedianWs ws,
                       size_t hist_bytes, cudaStream_t st)
{
    PILOT_CUDA(cudaMemsetAsync(ws.hist, 0, hist_bytes, st));
    const long long total = n * D;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    median_hist_kernel<T, PASS><<<(unsigned)blocks, 256, 0, st>>>(X, n, D, ldx, code, K, ws);
    PILOT_LAUNCH_CHECK();
    const int KD = K * D;
    const int warps = KD * 2;
    median_scan_kernel<T, PASS><<<(warps * 32 + 255) / 256, 256, 0, st>>>(KD, ws);
    PILOT_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int median_run(const T *X, long long n, int D, long long ldx, const int *code, int K, T *cent,
                      double *cent64, void *workspace, cudaStream_t st)
{
    const size_t kd = (size_t)K * D;
    MedianWs ws;
    unsigned char *p = (unsigned char *)workspace;
    ws.hist = (unsigned int *)p;
    const size_t hist_bytes = kd * 2 * 256 * sizeof(unsigned int);
    p += hist_bytes;
    ws.prefix = (unsigned long long *)p; p += kd * 2 * 8;
    ws.rank = (unsigned long long *)p;   p += kd * 2 * 8;
    ws.nvalid = (unsigned long long *)p;
    int rc;
#define PILOT_MEDIAN_PASS(P) \
    if (KeyOf<T>::PASSES > P) { rc = median_pass<T, (P < KeyOf<T>::PASSES ? P : 0)>(X, n, D, ldx, code, K, ws, hist_bytes, st); if (rc) return rc; }
    PILOT_MEDIAN_PASS(0) PILOT_MEDIAN_PASS(1) PILOT_MEDIAN_PASS(2) PILOT_MEDIAN_PASS(3)
    PILOT_MEDIAN_PASS(4) PILOT_MEDIAN_PASS(5) PILOT_MEDIAN_PASS(6) PILOT_MEDIAN_PASS(7)
#undef PILOT_MEDIAN_PASS
    median_final_kernel<T><<<(unsigned)((kd + 127) / 128), 128, 0, st>>>((int)kd, ws, cent, cent64);
    PILOT_LAUNCH_CHECK();
    return 0;
}

size_t median_ws_bytes(int K, int D) { return median_ws_bytes_impl(K, D); }

}  // namespace pilot

extern "C" int pilot_centroid_median(const void *X, int dtype, int64_t n_cells, int D, int64_t ldx,
                                     const int32_t *ct_code, int K, void *centroids, double *centroids_f64,
                                     void *workspace, size_t workspace_bytes, void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(dtype == PILOT_F32 || dtype == PILOT_F64, "pilot_centroid_median: dtype %d", dtype);
    PILOT_CHECK_ARG(K >= 1 && D >= 1 && n_cells >= 1 && ldx >= D, "pilot_centroid_median: bad shape");
    PILOT_CHECK_ARG(X && ct_code && centroids && centroids_f64 && workspace, "pilot_centroid_median: NULL pointer");
    PILOT_CHECK_ARG((long long)K * D < (1LL << 30), "pilot_centroid_median: K*D too large");
    PILOT_CHECK_ARG(n_cells < (1LL << 32), "pilot_centroid_median: n_cells must be < 2^32");
    PILOT_CHECK_ARG(workspace_bytes >= median_ws_bytes(K, D),
                    "pilot_centroid_median: workspace %zu < %zu bytes", workspace_bytes, median_ws_bytes(K, D));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PILOT_F32)
        return median_run<float>((const float *)X, n_cells, D, ldx, ct_code, K, (float *)centroids,
                                 centroids_f64, workspace, st);
    return median_run<double>((const double *)X, n_cells, D, ldx, ct_code, K, (double *)centroids,
                              centroids_f64, workspace, st);
}
