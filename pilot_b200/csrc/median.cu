// Kernel (2a): per-(cell type, dimension) MEDIAN of the embedding rows, in the
// input dtype.  Replaces data[annot.cell_type == k].median(axis=0)
// (reference pilotpy/tools/Trajectory.py:465-466; pandas nanmedian semantics:
// NaNs ignored, even counts -> (lo + hi) / 2 rounded in the input dtype).
//
// Two exact paths:
//  * sampled-pivot streaming path (default for n >= 64K): a block-strided sample of rows gives,
//    per (type, dim), two pivots lo <= hi that bracket the median with ~1e-7 failure odds;
//    ONE coalesced streaming pass over X then counts x < lo, x == lo, x == hi in warp-private
//    shared-memory tables (no atomics: lane = dimension, so a warp never collides with itself)
//    and appends the few elements with lo < x < hi (~10 %) to per-pair candidate lists; a
//    one-CTA-per-pair kernel finishes the selection inside the candidates.  If a pair's rank
//    falls outside its bracket (or its list overflows) the same CTA falls back to an exact
//    radix select over the full column -- no host round trip, results are always exact.
//    HBM traffic: one read of X + ~12 % for the sample + the candidate lists (L2 resident).
//  * most-significant-digit radix select (small inputs, huge K*D): every pass streams X once
//    and histograms the current 8-bit digit of the elements whose higher digits match the
//    running prefix of their (type, dim) query; 4 passes for f32, 8 for f64.
// Two queries per (type, dim) (ranks (n-1)/2 and n/2) so even counts need no second selection.
//
// Algorithmic bytes: one read of X + codes.
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


// =====================================================================================
// sampled-pivot streaming path over TYPE-SORTED row tiles
// =====================================================================================
constexpr int MED_SCAP = 2048;          // samples per (type, dim) for the pivots
constexpr double MED_SIGMAS = 5.5;      // half-width of the bracket in binomial sigmas
constexpr int MS_PROC = 4;              // tile processors (64 threads) per stream CTA
constexpr int MS_MAXCH = 4;             // column chunks of 64 per processor -> D <= 256
constexpr int MS_STAGE_BYTES = 32768;   // candidates staged in shared memory by the finish kernel
constexpr int MS_MAXITEMS_PER_TYPE = 2048;
constexpr int MS_FIN_THREADS = 256;

// shared-memory reduction without the compiler's warp-aggregation collective
__device__ __forceinline__ void red_shared_inc(unsigned int *p)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u) : "memory");
}

template <typename T> struct Inf;
template <> struct Inf<float> { __device__ static float pos() { return __int_as_float(0x7f800000); } };
template <> struct Inf<double> { __device__ static double pos() { return __longlong_as_double(0x7ff0000000000000LL); } };

// workspace header (first 64 bytes, zeroed per call)
struct MsHeader {
    unsigned int fail;          // pairs that took the exact full-column fallback (diagnostic)
    unsigned int item_counter;  // work queue of the stream kernel
    unsigned int n_items;
    unsigned int pad[13];
};

template <typename T> struct MsWs {
    MsHeader *hdr;
    unsigned int *cursor;      // K: scatter reservation cursors
    unsigned long long *type_cnt;  // K
    unsigned int *type_off;    // K + 1
    int *item_first;           // K + 1
    int *item_k;               // NI
    unsigned int *item_r0, *item_r1;  // NI: range inside sorted_rows
    unsigned int *sorted_rows; // n
    T *piv;                    // KD x 2
    unsigned int *cnt;         // NI x D x 5: below, eq_lo, eq_hi, nan, ncand
    T *cand;                   // NI x D x capi
    unsigned int item_rows, capi, ni_max;
};

__global__ void msort_count_kernel(const int *__restrict__ code, long long n, int K,
                                   unsigned long long *__restrict__ type_cnt)
{
    extern __shared__ unsigned int s_cnt[];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int k = __ldg(code + i);
        if ((unsigned)k < (unsigned)K) atomicAdd(&s_cnt[k], 1u);  // ATOMS.POPC.INC: hardware warp aggregation
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(&type_cnt[i], (unsigned long long)s_cnt[i]);
}

// one CTA: type offsets and the work items (type, row range) of the stream kernel
template <typename T>
__global__ void __launch_bounds__(256) msort_plan_kernel(int K, MsWs<T> ws)
{
    if (threadIdx.x == 0) {
        unsigned off = 0;
        int ni = 0;
        for (int k = 0; k < K; ++k) {
            const unsigned nk = (unsigned)ws.type_cnt[k];
            ws.type_off[k] = off;
            ws.item_first[k] = ni;
            ni += (int)((nk + ws.item_rows - 1) / ws.item_rows);
            off += nk;
        }
        ws.type_off[K] = off;
        ws.item_first[K] = ni;
        ws.hdr->n_items = (unsigned)ni;
    }
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const unsigned off = ws.type_off[k], nk = ws.type_off[k + 1] - off;
        const int first = ws.item_first[k], cnt = ws.item_first[k + 1] - first;
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
            const unsigned r = (unsigned)j * ws.item_rows;
            ws.item_k[first + j] = k;
            ws.item_r0[first + j] = off + r;
            ws.item_r1[first + j] = off + (r + ws.item_rows < nk ? r + ws.item_rows : nk);
        }
    }
}

// counting-sort scatter of the row ids by type: CTA-local histogram, one global reservation per
// (CTA, type), shared-memory cursors for the positions inside the reservation
template <typename T>
__global__ void __launch_bounds__(256)
msort_scatter_kernel(const int *__restrict__ code, long long n, int K, MsWs<T> ws)
{
    extern __shared__ unsigned int s_u[];  // hist[K], base[K]
    unsigned int *hist = s_u, *base = s_u + K;
    for (int i = threadIdx.x; i < K; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * per, c1 = c0 + per < n ? c0 + per : n;
    for (long long i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
        const int k = __ldg(code + i);
        if ((unsigned)k < (unsigned)K) atomicAdd(&hist[k], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        base[i] = hist[i] ? ws.type_off[i] + atomicAdd(&ws.cursor[i], hist[i]) : 0u;
        hist[i] = 0u;
    }
    __syncthreads();
    for (long long i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
        const int k = __ldg(code + i);
        if ((unsigned)k < (unsigned)K) {
            const unsigned pos = atomicAdd(&hist[k], 1u);
            ws.sorted_rows[base[k] + pos] = (unsigned)i;
        }
    }
}

// CTA-level radix select of two ranks over an arbitrary key source.  `known_prefix`/`first_shift`:
// digits above first_shift are already known to equal known_prefix for every key of interest.
template <typename T, typename Src>
__device__ void cta_select2_raw(Src src, long long r0, long long r1, unsigned int *hist /*[512]*/,
                                long long *sh /*[4]*/, typename KeyOf<T>::type known_prefix, int first_shift,
                                T &out0, T &out1)
{
    using KO = KeyOf<T>;
    using Key = typename KO::type;
    constexpr int BITS = sizeof(Key) * 8;
    Key p0 = known_prefix, p1 = known_prefix;
    long long q0 = r0, q1 = r1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int shift = first_shift; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 512; b += blockDim.x) hist[b] = 0u;
        __syncthreads();
        const bool top = shift + 8 >= BITS;
        const Key h0 = top ? (Key)0 : (Key)(p0 >> (shift + 8)), h1 = top ? (Key)0 : (Key)(p1 >> (shift + 8));
        src([&](T x) {
            const Key k = KO::key(x);
            const Key hi = top ? (Key)0 : (Key)(k >> (shift + 8));
            const unsigned dg = (unsigned)(k >> shift) & 0xffu;
            if (hi == h0) red_shared_inc(&hist[dg]);
            if (hi == h1) red_shared_inc(&hist[256 + dg]);
        });
        __syncthreads();
        if (warp < 2) {
            const unsigned int *h = hist + warp * 256;
            unsigned c[8];
            unsigned long long mine = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) { c[t] = h[lane * 8 + t]; mine += c[t]; }
            unsigned long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const unsigned long long excl = incl - mine;
            const unsigned long long rank = (unsigned long long)(warp ? q1 : q0);
            if (rank >= excl && rank < incl) {
                unsigned long long run = excl, below = excl;
                int digit = lane * 8 + 7;
                bool found = false;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (!found && rank < run + c[t]) { digit = lane * 8 + t; below = run; found = true; }
                    run += c[t];
                }
                sh[warp * 2] = digit;
                sh[warp * 2 + 1] = (long long)below;
            }
            const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0 && rank >= total) { sh[warp * 2] = 255; sh[warp * 2 + 1] = (long long)total; }  // n == 0
        }
        __syncthreads();
        p0 |= (Key)sh[0] << shift; q0 -= sh[1];
        p1 |= (Key)sh[2] << shift; q1 -= sh[3];
        __syncthreads();
    }
    out0 = KO::value(p0);
    out1 = KO::value(p1);
}

// one CTA per (type, dim): pivots lo <= hi bracketing the median, from <= MED_SCAP evenly spaced rows
// of the type's sorted segment, gathered once into shared memory
template <typename T>
__global__ void __launch_bounds__(256)
msort_pivot_kernel(const T *__restrict__ X, int D, long long ldx, MsWs<T> ws)
{
    using KO = KeyOf<T>;
    using Key = typename KO::type;
    constexpr int BITS = sizeof(Key) * 8;
    __shared__ unsigned int hist[512];
    __shared__ long long sh[4];
    __shared__ T vals[MED_SCAP];
    const int kd = blockIdx.x;
    const int k = kd / D, d = kd - k * D;
    const unsigned off = ws.type_off[k];
    const unsigned long long Nk = ws.type_off[k + 1] - off;
    const int ns = (int)(Nk < (unsigned long long)MED_SCAP ? Nk : (unsigned long long)MED_SCAP);
    T lo = -Inf<T>::pos(), hi = Inf<T>::pos();
    if (ns >= 64) {
        const int delta = (int)ceil(0.5 * MED_SIGMAS * sqrt((double)ns)) + 1;
        const int jlo = (ns - 1) / 2 - delta, jhi = ns / 2 + delta;
        if (jlo > 0 && jhi < ns - 1) {
            const T *Xd = X + d;
            for (int i0 = threadIdx.x; i0 < ns; i0 += 4 * 256) {
                unsigned rid[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 256 * u;
                    rid[u] = i < ns ? ws.sorted_rows[off + (unsigned)(((unsigned long long)i * Nk) / (unsigned)ns)] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 256 * u;
                    if (i < ns) vals[i] = Xd[(long long)rid[u] * ldx];
                }
            }
            __syncthreads();
            auto smp = [&](auto f) {
                for (int i = threadIdx.x; i < ns; i += 256) f(vals[i]);
            };
            // NaN samples carry the largest key, i.e. they sort last, as in np.sort
            cta_select2_raw<T>(smp, jlo, jhi, hist, sh, (Key)0, BITS - 8, lo, hi);
            if (hi != hi) hi = Inf<T>::pos();
            if (lo != lo) lo = -Inf<T>::pos();
        }
    }
    if (threadIdx.x == 0) {
        ws.piv[2 * kd] = lo;
        ws.piv[2 * kd + 1] = hi;
    }
}

template <int BYTES> __device__ __forceinline__ void cp_async_bytes(unsigned smem_addr, const void *src)
{
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(src) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(src) : "memory");
}

// The streaming pass.  A tile processor (64 threads) takes work items = (type, <= item_rows rows of
// that type); rows are gathered with cp.async into a double-buffered shared-memory tile (each row is
// one contiguous D-element read), then THREAD = COLUMN: the pivots of (type, d) and three counters
// (below, above, listed) live in registers; everything inside the closed bracket, and every NaN,
// goes to the thread's private list through a predicated store.  No atomics, no ballots, no tables,
// two compares per element.
template <typename T, int VB>  // VB = bytes per cp.async (row starts and D * sizeof(T) are multiples of it)
__global__ void __launch_bounds__(MS_PROC * 64)
msort_stream_kernel(const T *__restrict__ X, int D, long long ldx, int TR, MsWs<T> ws)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int s_item[MS_PROC];
    const int proc = threadIdx.x >> 6, ptid = threadIdx.x & 63, pw = ptid >> 5, lane = threadIdx.x & 31;
    T *tile = reinterpret_cast<T *>(smem_raw) + (size_t)proc * 2 * TR * D;  // [2][TR][D]
    const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
    const unsigned n_items = ws.hdr->n_items;
    const unsigned capi = ws.capi;
    const int row_bytes = D * (int)sizeof(T);
    const int nvec = row_bytes / VB;  // cp.async chunks per row
    for (;;) {
        if (ptid == 0) s_item[proc] = atomicAdd(&ws.hdr->item_counter, 1u);
        asm volatile("bar.sync %0, 64;" ::"r"(proc + 1) : "memory");
        const unsigned it = s_item[proc];
        asm volatile("bar.sync %0, 64;" ::"r"(proc + 1) : "memory");
        if (it >= n_items) break;
        const int k = ws.item_k[it];
        const unsigned r0 = ws.item_r0[it], r1 = ws.item_r1[it];
        T lo[MS_MAXCH], hi[MS_MAXCH];
        unsigned c_below[MS_MAXCH], c_above[MS_MAXCH], c_cand[MS_MAXCH];
#pragma unroll
        for (int ch = 0; ch < MS_MAXCH; ++ch) {
            const int d = ch * 64 + ptid;
            lo[ch] = d < D ? ws.piv[2 * ((size_t)k * D + d)] : (T)0;
            hi[ch] = d < D ? ws.piv[2 * ((size_t)k * D + d) + 1] : (T)0;
            c_below[ch] = c_above[ch] = c_cand[ch] = 0u;
        }
        const int ntiles = (int)((r1 - r0 + TR - 1) / TR);
        auto load_tile = [&](int t, int buf) {
            const unsigned rb = r0 + (unsigned)t * TR;
            const unsigned rid_mine = (rb + lane < r1 && lane < TR) ? ws.sorted_rows[rb + lane] : 0u;
            const unsigned dst = tile_s + (unsigned)(buf * TR) * row_bytes;
            const int nrow = (int)(r1 - rb < (unsigned)TR ? r1 - rb : (unsigned)TR);
            for (int row = pw; row < nrow; row += 2) {
                const unsigned rid = __shfl_sync(0xffffffffu, rid_mine, row);
                const char *src = reinterpret_cast<const char *>(X + (long long)rid * ldx);
                for (int c = lane; c < nvec; c += 32) cp_async_bytes<VB>(dst + row * row_bytes + c * VB, src + c * VB);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        load_tile(0, 0);
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            if (t + 1 < ntiles) {
                load_tile(t + 1, buf ^ 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            asm volatile("bar.sync %0, 64;" ::"r"(proc + 1) : "memory");
            const unsigned rb = r0 + (unsigned)t * TR;
            const int rows = (int)(r1 - rb < (unsigned)TR ? r1 - rb : (unsigned)TR);
            const T *src = tile + (size_t)buf * TR * D;
#pragma unroll
            for (int ch = 0; ch < MS_MAXCH; ++ch) {
                const int d = ch * 64 + ptid;
                if (d < D) {
                    T *cl = ws.cand + ((size_t)it * D + d) * capi;
                    const T l = lo[ch], h = hi[ch];
                    // lo == hi: the sample saw a plateau of ties -- they are counted (everything that is
                    // neither below nor above nor stored), not stored; x != NaN is true for every x
                    const T tie = l == h ? l : (T)NAN;
                    unsigned nb = c_below[ch], na = c_above[ch], nc = c_cand[ch];
                    const T *col = src + d;
#pragma unroll 4
                    for (int row = 0; row < rows; ++row) {
                        const T x = col[row * D];
                        const bool lt = x < l, gt = x > h;
                        nb += lt ? 1u : 0u;
                        na += gt ? 1u : 0u;
                        // the closed bracket [lo, hi] and every NaN go to the list
                        const bool st = !lt && !gt && x != tie;
                        if (st && nc < capi) cl[nc] = x;
                        nc += st ? 1u : 0u;
                    }
                    c_below[ch] = nb; c_above[ch] = na; c_cand[ch] = nc;
                }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(proc + 1) : "memory");  // tile buffer free again
        }
#pragma unroll
        for (int ch = 0; ch < MS_MAXCH; ++ch) {
            const int d = ch * 64 + ptid;
            if (d < D) {
                unsigned int *o = ws.cnt + ((size_t)it * D + d) * 5;
                o[0] = c_below[ch]; o[1] = c_above[ch]; o[4] = c_cand[ch];
            }
        }
    }
}

// one CTA per (type, dim): rank bookkeeping over the type's items (below / tie plateau / lists /
// above), NaN count and selection inside the lists, exact fallback over the type's own rows when the
// bracket missed or a list overflowed
template <typename T>
__global__ void __launch_bounds__(MS_FIN_THREADS)
msort_finish_kernel(const T *__restrict__ X, int D, long long ldx, MsWs<T> ws, T *__restrict__ cent,
                    double *__restrict__ cent64)
{
    using KO = KeyOf<T>;
    using Key = typename KO::type;
    constexpr int BITS = sizeof(Key) * 8;
    constexpr int STAGE = MS_STAGE_BYTES / (int)sizeof(T);
    __shared__ unsigned int hist[512];
    __shared__ long long sh[4];
    __shared__ unsigned long long s_sum[5];
    __shared__ unsigned int s_off[MS_MAXITEMS_PER_TYPE + 1];
    __shared__ T s_stage[STAGE];
    __shared__ int s_over;
    const int kd = blockIdx.x;
    const int k = kd / D, d = kd - k * D;
    const int i0 = ws.item_first[k], i1 = ws.item_first[k + 1];
    const int nit = i1 - i0;
    if (threadIdx.x < 5) s_sum[threadIdx.x] = 0ULL;
    if (threadIdx.x == 0) s_over = nit > MS_MAXITEMS_PER_TYPE ? 1 : 0;
    __syncthreads();
    {
        unsigned long long a[2] = {0, 0};
        for (int q = threadIdx.x; q < nit; q += blockDim.x) {
            const unsigned int *c = ws.cnt + ((size_t)(i0 + q) * D + d) * 5;
            a[0] += c[0]; a[1] += c[1];
            const unsigned nc = c[4];
            if (nc > ws.capi) s_over = 1;
            if (q < MS_MAXITEMS_PER_TYPE) s_off[q + 1] = nc < ws.capi ? nc : ws.capi;
        }
        // one shared 64-bit atomic per warp (they are CAS loops: 256 contending threads would serialise)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
            if ((threadIdx.x & 31) == 0 && a[j]) atomicAdd(&s_sum[j], a[j]);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // inclusive scan of s_off[1..m] by one warp, 32 entries per step
        const int m = nit < MS_MAXITEMS_PER_TYPE ? nit : MS_MAXITEMS_PER_TYPE;
        unsigned carry = 0;
        for (int base = 0; base < m; base += 32) {
            const int q = base + threadIdx.x;
            unsigned v = q < m ? s_off[q + 1] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xffffffffu, v, o);
                if ((int)threadIdx.x >= o) v += u;
            }
            if (q < m) s_off[q + 1] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (threadIdx.x == 0) { s_off[0] = 0; s_sum[4] = carry; }
    }
    __syncthreads();
    const bool overflow = s_over != 0;
    const long long below = (long long)s_sum[0], above = (long long)s_sum[1], nlist = (long long)s_sum[4];
    const unsigned toff = ws.type_off[k];
    const long long Nk = (long long)ws.type_off[k + 1] - toff;
    const T lo = ws.piv[2 * kd], hi = ws.piv[2 * kd + 1];
    // the lists hold the closed bracket [lo, hi] and every NaN; with lo == hi the ties were counted
    // instead of listed: ties = Nk - below - above - listed
    const long long ties = lo == hi ? Nk - below - above - nlist : 0;
    const bool staged = !overflow && nlist <= STAGE;
    if (staged) {
        // warp w copies the lists of items w, w + 8, ... (each a short contiguous run)
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        // four lists per round, so that four independent global loads are in flight per lane
        constexpr int NW = MS_FIN_THREADS / 32;
        for (int q = warp; q < nit; q += 4 * NW) {
            const T *cl[4];
            unsigned o[4], c[4], cmax = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int qq = q + u * NW;
                const bool ok = qq < nit;
                cl[u] = ws.cand + ((size_t)(i0 + (ok ? qq : q)) * D + d) * ws.capi;
                o[u] = ok ? s_off[qq] : 0u;
                c[u] = ok ? s_off[qq + 1] - o[u] : 0u;
                cmax = c[u] > cmax ? c[u] : cmax;
            }
            for (unsigned i = lane; i < cmax; i += 32) {
                T v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = i < c[u] ? cl[u][i] : (T)0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i < c[u]) s_stage[o[u] + i] = v[u];
            }
        }
        __syncthreads();
    }
    // every list element, NaN included
    auto lst_all = [&](auto f) {
        if (staged) {
            for (unsigned i = threadIdx.x; i < (unsigned)nlist; i += blockDim.x) f(s_stage[i]);
        } else {
            for (int q = 0; q < nit; ++q) {
                const T *cl = ws.cand + ((size_t)(i0 + q) * D + d) * ws.capi;
                const unsigned c = s_off[q + 1] - s_off[q];
                for (unsigned i = threadIdx.x; i < c; i += blockDim.x) f(cl[i]);
            }
        }
    };
    long long n_nan = 0;
    if (!overflow) {
        unsigned my = 0;
        lst_all([&](T x) { my += x != x ? 1u : 0u; });
        if (my) atomicAdd(&s_sum[3], (unsigned long long)my);
        __syncthreads();
        n_nan = (long long)s_sum[3];
    }
    const long long nmid = nlist - n_nan;  // listed, orderable
    const long long nvalid = Nk - n_nan;
    T med;
    if (nvalid <= 0 && !overflow) {
        med = (T)NAN;
    } else {
        const long long r0 = (nvalid - 1) / 2, r1 = nvalid / 2;
        // where does each rank land?  0: < lo (fail) 1: the tie plateau lo == hi 2: the lists 4: beyond (fail)
        auto region = [&](long long r, long long &rin) -> int {
            if (r < below) return 0;
            r -= below;
            if (r < ties) return 1;
            r -= ties;
            if (r < nmid) { rin = r; return 2; }
            return 4;
        };
        long long j0 = 0, j1 = 0;
        const int g0 = overflow ? 0 : region(r0, j0), g1 = overflow ? 0 : region(r1, j1);
        T v0, v1;
        if (overflow || g0 == 0 || g0 == 4 || g1 == 0 || g1 == 4) {
            // exact fallback: radix select over this type's own rows
            if (threadIdx.x == 0) atomicAdd(&ws.hdr->fail, 1u);
            __shared__ unsigned long long s_valid;
            if (threadIdx.x == 0) s_valid = 0ULL;
            __syncthreads();
            {
                unsigned long long my = 0;
                for (long long i = threadIdx.x; i < Nk; i += blockDim.x) {
                    const T x = X[(long long)ws.sorted_rows[toff + i] * ldx + d];
                    my += x == x ? 1ULL : 0ULL;
                }
                if (my) atomicAdd(&s_valid, my);
            }
            __syncthreads();
            const long long nv = (long long)s_valid;
            if (nv <= 0) {
                v0 = v1 = (T)NAN;
            } else {
                auto col = [&](auto f) {
                    for (long long i = threadIdx.x; i < Nk; i += blockDim.x) {
                        const T x = X[(long long)ws.sorted_rows[toff + i] * ldx + d];
                        if (x == x) f(x);
                    }
                };
                cta_select2_raw<T>(col, (nv - 1) / 2, nv / 2, hist, sh, (Key)0, BITS - 8, v0, v1);
            }
        } else {
            T m0 = lo, m1 = lo;
            if (g0 == 2 || g1 == 2) {
                auto lst = [&](auto f) { lst_all([&](T x) { if (x == x) f(x); }); };
                // every orderable list element lies in [lo, hi]: the digits above the first one in
                // which key(lo) and key(hi) differ are common to all of them
                Key prefix = 0;
                int first = BITS - 8;
                if (lo > -Inf<T>::pos() && hi < Inf<T>::pos()) {
                    const Key kl = KO::key(lo), kh = KO::key(hi);
                    while (first > 0 && (kl >> first) == (kh >> first)) first -= 8;
                    if (first < BITS - 8) prefix = (Key)((kl >> (first + 8)) << (first + 8));
                }
                cta_select2_raw<T>(lst, g0 == 2 ? j0 : 0, g1 == 2 ? j1 : 0, hist, sh, prefix, first, m0, m1);
            }
            v0 = g0 == 1 ? lo : m0;
            v1 = g1 == 1 ? lo : m1;
        }
        if (v0 == v1) med = v0;
        else if (sizeof(T) == 4) med = (T)__fdiv_rn(__fadd_rn((float)v0, (float)v1), 2.0f);
        else med = (T)__ddiv_rn(__dadd_rn((double)v0, (double)v1), 2.0);
    }
    if (threadIdx.x == 0) {
        cent[kd] = med;
        cent64[kd] = (double)med;
    }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct MsPlan {
    unsigned item_rows, capi, ni_max;
    int TR;
};

static MsPlan ms_plan(long long n, int K, int D, size_t elt)
{
    MsPlan p;
    long long ir = (n + (long long)sm_count() * 16 - 1) / ((long long)sm_count() * 16);
    ir = (ir + 31) / 32 * 32;
    if (ir < 128) ir = 128;
    p.item_rows = (unsigned)ir;
    p.capi = (unsigned)(0.35 * (double)ir) + 32;
    p.ni_max = (unsigned)(n / ir) + (unsigned)K + 2;
    long long tr = (48 * 1024) / ((long long)MS_PROC * 2 * D * (long long)elt);  // ~48 KB of tiles per CTA
    if (tr > 32) tr = 32;
    p.TR = (int)(tr < 4 ? 4 : tr);
    return p;
}

template <typename T>
static size_t ms_ws_bytes(long long n, int K, int D)
{
    const MsPlan p = ms_plan(n, K, D, sizeof(T));
    const size_t kd = (size_t)K * D;
    size_t b = 256;                                   // header
    b += align256((size_t)K * 4);                     // cursor
    b += align256((size_t)K * 8);                     // type_cnt
    b += align256((size_t)(K + 1) * 4) * 2;           // type_off, item_first
    b += align256((size_t)p.ni_max * 4) * 3;          // item_k, item_r0, item_r1
    b += align256((size_t)n * 4);                     // sorted_rows
    b += align256(kd * 2 * sizeof(T));                // piv
    b += align256((size_t)p.ni_max * D * 5 * 4);      // cnt
    b += align256((size_t)p.ni_max * D * p.capi * sizeof(T));  // cand
    return b;
}

template <typename T>
static int median_run_sorted(const T *X, long long n, int D, long long ldx, const int *code, int K, T *cent,
                             double *cent64, void *workspace, cudaStream_t st)
{
    const MsPlan pl = ms_plan(n, K, D, sizeof(T));
    const size_t kd = (size_t)K * D;
    MsWs<T> ws;
    unsigned char *p = (unsigned char *)workspace;
    ws.hdr = (MsHeader *)p; p += 256;
    ws.cursor = (unsigned int *)p; p += align256((size_t)K * 4);
    ws.type_cnt = (unsigned long long *)p; p += align256((size_t)K * 8);
    const size_t zero_bytes = (size_t)(p - (unsigned char *)workspace);
    ws.type_off = (unsigned int *)p; p += align256((size_t)(K + 1) * 4);
    ws.item_first = (int *)p; p += align256((size_t)(K + 1) * 4);
    ws.item_k = (int *)p; p += align256((size_t)pl.ni_max * 4);
    ws.item_r0 = (unsigned int *)p; p += align256((size_t)pl.ni_max * 4);
    ws.item_r1 = (unsigned int *)p; p += align256((size_t)pl.ni_max * 4);
    ws.sorted_rows = (unsigned int *)p; p += align256((size_t)n * 4);
    ws.piv = (T *)p; p += align256(kd * 2 * sizeof(T));
    ws.cnt = (unsigned int *)p; p += align256((size_t)pl.ni_max * D * 5 * 4);
    ws.cand = (T *)p;
    ws.item_rows = pl.item_rows; ws.capi = pl.capi; ws.ni_max = pl.ni_max;
    PILOT_CUDA(cudaMemsetAsync(workspace, 0, zero_bytes, st));
    long long blocks = (n + 2047) / 2048;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    msort_count_kernel<<<(unsigned)blocks, 256, K * sizeof(unsigned int), st>>>(code, n, K, ws.type_cnt);
    PILOT_LAUNCH_CHECK();
    msort_plan_kernel<T><<<1, 256, 0, st>>>(K, ws);
    PILOT_LAUNCH_CHECK();
    msort_scatter_kernel<T><<<(unsigned)blocks, 256, 2 * K * sizeof(unsigned int), st>>>(code, n, K, ws);
    PILOT_LAUNCH_CHECK();
    msort_pivot_kernel<T><<<(unsigned)kd, 256, 0, st>>>(X, D, ldx, ws);
    PILOT_LAUNCH_CHECK();
    {
        const size_t smem = (size_t)MS_PROC * 2 * pl.TR * D * sizeof(T);
        // widest cp.async every row start and row length allow (the tile rows inherit the alignment)
        const size_t rowb = (size_t)D * sizeof(T), strideb = (size_t)ldx * sizeof(T);
        int vb = (int)sizeof(T);
        if (rowb % 16 == 0 && strideb % 16 == 0 && ((uintptr_t)X % 16) == 0) vb = 16;
        else if (rowb % 8 == 0 && strideb % 8 == 0 && ((uintptr_t)X % 8) == 0) vb = 8;
#define MS_LAUNCH(VBV)                                                                                          \
        do {                                                                                                    \
            PILOT_CUDA(cudaFuncSetAttribute(msort_stream_kernel<T, VBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            (int)smem));                                                        \
            int per_sm = 1;                                                                                     \
            PILOT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msort_stream_kernel<T, VBV>,      \
                                                                     MS_PROC * 64, smem));                      \
            if (per_sm < 1) per_sm = 1;                                                                         \
            if (per_sm > 4) per_sm = 4;                                                                         \
            msort_stream_kernel<T, VBV><<<sm_count() * per_sm, MS_PROC * 64, smem, st>>>(X, D, ldx, pl.TR, ws);  \
        } while (0)
        if (vb == 16) MS_LAUNCH(16);
        else if (vb == 8) MS_LAUNCH(8);
        else MS_LAUNCH((int)sizeof(T));
#undef MS_LAUNCH
        PILOT_LAUNCH_CHECK();
    }
    msort_finish_kernel<T><<<(unsigned)kd, MS_FIN_THREADS, 0, st>>>(X, D, ldx, ws, cent, cent64);
    PILOT_LAUNCH_CHECK();
    return 0;
}

static bool median_use_sorted(long long n, int K, int D)
{
    return n >= 65536 && n < (1LL << 32) && D <= 64 * MS_MAXCH && K <= 4096;
}

size_t median_ws_bytes(long long n, int K, int D)
{
    size_t a = median_ws_bytes_impl(K, D);
    if (median_use_sorted(n, K, D)) {
        const size_t b = ms_ws_bytes<double>(n, K, D);
        if (b > a) a = b;
    }
    return a;
}
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
    PILOT_CHECK_ARG(workspace_bytes >= median_ws_bytes(n_cells, K, D),
                    "pilot_centroid_median: workspace %zu < %zu bytes", workspace_bytes,
                    median_ws_bytes(n_cells, K, D));
    cudaStream_t st = (cudaStream_t)stream;
    if (median_use_sorted(n_cells, K, D)) {
        if (dtype == PILOT_F32)
            return median_run_sorted<float>((const float *)X, n_cells, D, ldx, ct_code, K, (float *)centroids,
                                            centroids_f64, workspace, st);
        return median_run_sorted<double>((const double *)X, n_cells, D, ldx, ct_code, K, (double *)centroids,
                                         centroids_f64, workspace, st);
    }
    if (dtype == PILOT_F32)
        return median_run<float>((const float *)X, n_cells, D, ldx, ct_code, K, (float *)centroids,
                                 centroids_f64, workspace, st);
    return median_run<double>((const double *)X, n_cells, D, ldx, ct_code, K, (double *)centroids,
                              centroids_f64, workspace, st);
}
