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
// sampled-pivot streaming path
// =====================================================================================
constexpr int MED_SCAP = 4096;          // samples kept per type
template <typename T> struct MedChunk { static constexpr int V = sizeof(T) == 4 ? 8 : 4; };  // dims per pivot CTA
constexpr int MED_PIV_THREADS = 256;    // 8 warps: one per dim of the chunk
constexpr int MED_SAMPLE_BLOCK = 32;    // consecutive rows per sample block
constexpr double MED_SIGMAS = 5.5;      // half-width of the bracket in binomial sigmas
constexpr int MED_STAGE_BYTES = 32768;   // candidates staged in shared memory by the finish kernel
constexpr int MED_G = 32;               // candidate sub-lists per (type, dim): spreads the binning atomics

// shared-memory reduction without the compiler's warp-aggregation collective (which costs far
// more than the conflict it avoids when the bins are spread)
__device__ __forceinline__ void red_shared_inc(unsigned int *p)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u) : "memory");
}

template <typename T> struct Inf;
template <> struct Inf<float> { __device__ static float pos() { return __int_as_float(0x7f800000); } };
template <> struct Inf<double> { __device__ static double pos() { return __longlong_as_double(0x7ff0000000000000LL); } };

struct MedianSampling {
    long long m_s;      // sampled rows
    long long nblocks;  // sample blocks
    long long bstride;  // rows between block starts
};

static MedianSampling median_sampling(long long n, int K)
{
    MedianSampling sp;
    long long target = (long long)MED_SCAP * K;
    if (target > n) target = n;
    sp.nblocks = (target + MED_SAMPLE_BLOCK - 1) / MED_SAMPLE_BLOCK;
    if (sp.nblocks < 1) sp.nblocks = 1;
    sp.bstride = (n / sp.nblocks) / MED_SAMPLE_BLOCK * MED_SAMPLE_BLOCK;  // warp-aligned sample blocks
    if (sp.bstride < MED_SAMPLE_BLOCK) sp.bstride = MED_SAMPLE_BLOCK;
    while (sp.nblocks > 1 && (sp.nblocks - 1) * sp.bstride + MED_SAMPLE_BLOCK > n) --sp.nblocks;
    sp.m_s = sp.nblocks * MED_SAMPLE_BLOCK;
    if (sp.m_s > n) sp.m_s = n;
    return sp;
}

template <typename T> struct MedianWs2 {
    unsigned long long *type_cnt;  // K
    unsigned long long *cand_total;  // 1: running offset allocator
    unsigned int *ccnt;            // KD: candidates appended
    unsigned int *fail;            // 1: pairs that needed the full-column fallback (diagnostic)
    unsigned int *cap;             // KD
    unsigned long long *coff;      // KD
    T *piv;                        // KD * 2
    unsigned int *partial;         // ctasB * KD * 2
    T *cand;                       // cand_capacity
    unsigned long long cand_capacity;
    unsigned int *scnt;            // K: sampled rows seen per type
    int *slist;                    // K x MED_SCAP sampled row ids
    unsigned int *log_cnt;         // TW: records in each warp's private candidate log
    unsigned int *log_flag;        // 1: some log overflowed -> every pair takes the exact fallback
    int *log_kd;                   // TW x log_cap
    T *log_val;                    // TW x log_cap
    unsigned int log_cap;
};

__global__ void median_count_kernel(const int *__restrict__ code, long long n, int K, MedianSampling sp,
                                    unsigned long long *__restrict__ type_cnt, unsigned int *__restrict__ scnt,
                                    int *__restrict__ slist)
{
    extern __shared__ unsigned int s_cnt[];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nround = (n + stride - 1) / stride;
    const int lane = threadIdx.x & 31;
    for (long long it = 0; it < nround; ++it) {
        const long long i = it * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const int k = i < n ? __ldg(code + i) : -1;
        const bool ok = (unsigned)k < (unsigned)K;
        const unsigned grp = __match_any_sync(0xffffffffu, ok ? k : -1 - lane);
        if (ok && lane == __ffs(grp) - 1) atomicAdd(&s_cnt[k], (unsigned)__popc(grp));
        // block-strided row sample: rows [b*bstride, b*bstride + 32) for b < nblocks
        const long long b = i / sp.bstride;
        const bool smp = ok && b < sp.nblocks && (i - b * sp.bstride) < MED_SAMPLE_BLOCK;
        if (!__any_sync(0xffffffffu, smp)) continue;  // warp-aligned blocks: ~1 warp-step in 8 gets here
        const unsigned sg = __match_any_sync(0xffffffffu, smp ? k : -1 - lane);
        if (smp) {
            const int leader = __ffs(sg) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&scnt[k], (unsigned)__popc(sg));
            base = __shfl_sync(sg, base, leader);
            const unsigned pos = base + __popc(sg & ((1u << lane) - 1u));
            if (pos < (unsigned)MED_SCAP) slist[(size_t)k * MED_SCAP + pos] = (int)i;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(&type_cnt[i], (unsigned long long)s_cnt[i]);
}

// the streaming pass: one coalesced read of X.  Warp w of the grid owns a contiguous row range;
// lanes walk consecutive elements (lane = dimension, D >= 32, so the 32 lanes of one step never
// share a (type, dim) counter and the shared-memory read-modify-write needs no atomics); MED_U
// independent loads per lane are in flight before the first is consumed.  Elements inside the
// bracket go to the warp's PRIVATE append log (ballot prefix, plain stores, no atomics).
constexpr int MED_U = 8;

template <typename T>
__global__ void __launch_bounds__(1024)
median_stream_kernel(const T *__restrict__ X, long long n, int D, long long ldx, const int *__restrict__ code,
                     int K, int W, MedianWs2<T> ws)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int KD = K * D;
    T *spiv = reinterpret_cast<T *>(smem_raw);                                   // KD x 2 (lo, hi)
    unsigned int *rare = reinterpret_cast<unsigned int *>(spiv + 2 * (size_t)KD);  // KD: x == hi | NaN << 16 (atomics)
    unsigned int *cnt = rare + KD;                                               // W x KD: x < lo | x == lo << 16
    for (int i = threadIdx.x; i < 2 * KD; i += blockDim.x) spiv[i] = ws.piv[i];
    for (int i = threadIdx.x; i < (W + 1) * KD; i += blockDim.x) rare[i] = 0u;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        unsigned int *mycnt = cnt + (size_t)warp * KD;
        const long long TW = (long long)gridDim.x * W;
        const long long gw = (long long)blockIdx.x * W + warp;
        const long long r_begin = gw * n / TW, r_end = (gw + 1) * n / TW;
        int *lkd = ws.log_kd + (size_t)gw * ws.log_cap;
        T *lval = ws.log_val + (size_t)gw * ws.log_cap;
        unsigned nlog = 0;
        // all index arithmetic below is 32-bit and local to my row range (rows_per_warp * ldx < 2^31)
        const int nrows = (int)(r_end - r_begin);
        const int e_end = nrows * D;  // elements of my range, flat (row-major, d fastest)
        const T *Xw = X + r_begin * ldx;
        const int *cw = code + r_begin;
        const int ldx32 = (int)ldx;
        const unsigned lt_mask = (1u << lane) - 1u;
        auto consume = [&](T x, int kd) {
            bool cand = false;
            if (kd >= 0) {
                const T lo = spiv[2 * kd], hi = spiv[2 * kd + 1];
                const unsigned inc0 = x < lo ? 1u : (x == lo ? 0x10000u : 0u);
                if (inc0) mycnt[kd] += inc0;
                if ((x == hi && hi != lo) || x != x) {
                    // plain shared-memory reduction (inline PTX keeps the compiler from wrapping this
                    // rare path in its warp-aggregation collective)
                    const unsigned addr = (unsigned)__cvta_generic_to_shared(rare + kd);
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(x != x ? 0x10000u : 1u) : "memory");
                }
                cand = x > lo && x < hi;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, cand);
            if (cand) {
                const unsigned pos = nlog + __popc(bal & lt_mask);
                if (pos < ws.log_cap) { lkd[pos] = kd; lval[pos] = x; }
            }
            nlog += __popc(bal);
        };

        // The warp walks its contiguous range in blocks of 32 * MED_U elements.  (rowb, remb) =
        // (eb / D, eb % D) is kept incrementally; the <= 256/D + 2 type codes a block touches are
        // loaded by the first lanes and handed out by shuffle; X is read through one running
        // pointer with compile-time offsets (the matrix is contiguous: ldx == D).
        constexpr int BLK = 32 * MED_U;
        const int blk_rows = BLK / D, blk_rem = BLK - blk_rows * D;
        int rowb = 0, remb = 0;
        const T *p = Xw + lane;
        int eb = 0;
        for (; eb < e_end; eb += BLK, p += BLK) {
            const bool full = eb + BLK <= e_end;
            const int cval = (rowb + lane < nrows) ? __ldg(cw + rowb + lane) : -1;
            T xv[MED_U];
#pragma unroll
            for (int j = 0; j < MED_U; ++j) xv[j] = (full || eb + 32 * j + lane < e_end) ? p[32 * j] : (T)0;
            int d = remb + lane, ro = 0;
            if (d >= D) { d -= D; ro = 1; }
#pragma unroll
            for (int j = 0; j < MED_U; ++j) {
                const int k = __shfl_sync(0xffffffffu, cval, ro);
                const bool ok = full || eb + 32 * j + lane < e_end;
                consume(xv[j], (ok && (unsigned)k < (unsigned)K) ? k * D + d : -1);
                d += 32;
                if (d >= D) { d -= D; ++ro; }  // D >= 32: at most one wrap per step
            }
            rowb += blk_rows;
            remb += blk_rem;
            if (remb >= D) { remb -= D; ++rowb; }
        }
        if (lane == 0) {
            ws.log_cnt[gw] = nlog < ws.log_cap ? nlog : ws.log_cap;
            if (nlog > ws.log_cap) atomicOr(ws.log_flag, 1u);
        }
    }
    __syncthreads();
    // per-CTA partial table: [cta][kd][4] = x < lo, x == lo, x == hi, NaN
    unsigned int *out = ws.partial + (size_t)blockIdx.x * KD * 4;
    for (int i = threadIdx.x; i < KD; i += blockDim.x) {
        unsigned a = 0, b = 0;
        for (int w = 0; w < W; ++w) {
            const unsigned v = cnt[(size_t)w * KD + i];
            a += v & 0xffffu;
            b += v >> 16;
        }
        const unsigned rr = rare[i];
        *reinterpret_cast<uint4 *>(out + 4 * (size_t)i) = make_uint4(a, b, rr & 0xffffu, rr >> 16);
    }
}

// regroup the warp logs by (type, dim): one CTA per log, one returning atomic per record -- here
// thousands of independent records are in flight, so the atomics are throughput- not latency-bound
template <typename T>
__global__ void __launch_bounds__(256)
median_bin_kernel(MedianWs2<T> ws)
{
    const size_t gw = blockIdx.x;
    const unsigned nrec = ws.log_cnt[gw];
    const int *lkd = ws.log_kd + gw * ws.log_cap;
    const T *lval = ws.log_val + gw * ws.log_cap;
    const unsigned g = (unsigned)(gw % MED_G);
    for (unsigned i0 = threadIdx.x; i0 < nrec; i0 += 4 * blockDim.x) {
        int kd[4];
        T v[4];
        unsigned capv[4], slot[4];
        unsigned long long off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned i = i0 + u * blockDim.x;
            kd[u] = i < nrec ? lkd[i] : -1;
            v[u] = i < nrec ? lval[i] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (kd[u] >= 0) {
                capv[u] = ws.cap[kd[u]];
                off[u] = ws.coff[kd[u]];
                slot[u] = atomicAdd(&ws.ccnt[(size_t)kd[u] * MED_G + g], 1u);
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (kd[u] >= 0 && slot[u] < capv[u]) ws.cand[off[u] + (unsigned long long)g * capv[u] + slot[u]] = v[u];
    }
}

// CTA-level radix select of two ranks over an arbitrary key source (candidate lists or a full
// column).  `known_prefix`/`first_shift`: digits above first_shift are already known to equal
// known_prefix for every key of interest (all candidates lie between the two pivots).
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
                unsigned long long run = excl;
                int digit = lane * 8 + 7;
                unsigned long long below = excl;
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

// one CTA per (type, dim): pivots lo/hi bracketing the median, from the type's sampled rows
template <typename T>
__global__ void __launch_bounds__(128)
median_pivot_kernel(const T *__restrict__ X, long long n, int D, long long ldx, const int *__restrict__ code,
                    int K, MedianSampling sp, MedianWs2<T> ws)
{
    using KO = KeyOf<T>;
    using Key = typename KO::type;
    constexpr int BITS = sizeof(Key) * 8;
    __shared__ unsigned int hist[512];
    __shared__ long long sh[4];
    const int kd = blockIdx.x;
    const int k = kd / D, d = kd - k * D;
    const int ns = (int)min(ws.scnt[k], (unsigned)MED_SCAP);
    const int *rows = ws.slist + (size_t)k * MED_SCAP;
    const unsigned long long Nk = ws.type_cnt[k];
    T lo = -Inf<T>::pos(), hi = Inf<T>::pos();
    double frac = 1.0;
    if (ns >= 64) {
        const int delta = (int)ceil(0.5 * MED_SIGMAS * sqrt((double)ns)) + 1;
        const int jlo = (ns - 1) / 2 - delta, jhi = ns / 2 + delta;
        if (jlo > 0 && jhi < ns - 1) {
            const T *Xd = X + d;
            auto smp = [&](auto f) {
                for (int i0 = threadIdx.x; i0 < ns; i0 += 4 * 128) {
                    T xv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + 128 * u;
                        xv[u] = i < ns ? Xd[(long long)__ldg(rows + i) * ldx] : (T)0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + 128 * u < ns) f(xv[u]);
                }
            };
            // NaN samples carry the largest key, i.e. they sort last, as in np.sort
            cta_select2_raw<T>(smp, jlo, jhi, hist, sh, (Key)0, BITS - 8, lo, hi);
            if (hi != hi) hi = Inf<T>::pos();
            if (lo != lo) lo = -Inf<T>::pos();
            frac = (double)(jhi - jlo + 1) / ns;
        }
    }
    if (threadIdx.x == 0) {
        ws.piv[2 * kd] = lo;
        ws.piv[2 * kd + 1] = hi;
        // MED_G sub-lists per pair, each sized for its share (x2: the split is binomial) plus slack
        unsigned long long want = (unsigned long long)(frac * (double)Nk * 2.0 / MED_G) + 128ULL;
        if (want > Nk) want = Nk;
        if (want > 0x0fffffffULL) want = 0x0fffffffULL;
        const unsigned long long off = atomicAdd(ws.cand_total, want * MED_G);
        unsigned int cap = (unsigned int)want;
        if (off + want * MED_G > ws.cand_capacity) cap = 0;  // out of candidate space -> exact fallback
        ws.cap[kd] = cap;   // per sub-list
        ws.coff[kd] = off;  // sub-list g starts at off + g * cap
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
median_finish_kernel(const T *__restrict__ X, long long n, int D, long long ldx, const int *__restrict__ code,
                     int K, int ctasB, MedianWs2<T> ws, T *__restrict__ cent, double *__restrict__ cent64)
{
    using KO = KeyOf<T>;
    using Key = typename KO::type;
    constexpr int BITS = sizeof(Key) * 8;
    __shared__ unsigned int hist[512];
    __shared__ long long sh[4];
    __shared__ unsigned long long s_sum[4];
    __shared__ unsigned int s_goff[MED_G + 1];
    constexpr int MED_STAGE = MED_STAGE_BYTES / (int)sizeof(T);
    __shared__ T s_stage[MED_STAGE];
    const int kd = blockIdx.x;
    const int k = kd / D, d = kd - k * D;
    if (threadIdx.x < 4) s_sum[threadIdx.x] = 0ULL;
    __syncthreads();
    {
        unsigned long long a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int c = threadIdx.x; c < ctasB; c += blockDim.x) {
            const uint4 v = *reinterpret_cast<const uint4 *>(ws.partial + ((size_t)c * K * D + kd) * 4);
            a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
        }
        atomicAdd(&s_sum[0], a0); atomicAdd(&s_sum[1], a1); atomicAdd(&s_sum[2], a2); atomicAdd(&s_sum[3], a3);
    }
    const unsigned int capv = ws.cap[kd];
    // sub-list sizes: one lane per sub-list, warp scan for the offsets
    __shared__ int s_over;
    if (threadIdx.x == 0) s_over = *ws.log_flag != 0u ? 1 : 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        static_assert(MED_G == 32, "one lane per sub-list");
        const unsigned c = ws.ccnt[(size_t)kd * MED_G + threadIdx.x];
        const unsigned cc = c < capv ? c : capv;
        if (c > capv) s_over = 1;
        unsigned incl = cc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (threadIdx.x >= o) incl += v;
        }
        s_goff[threadIdx.x] = incl - cc;
        if (threadIdx.x == 31) s_goff[32] = incl;
    }
    __syncthreads();
    const bool overflow = s_over != 0;
    const long long below = (long long)s_sum[0], eq_lo = (long long)s_sum[1], eq_hi = (long long)s_sum[2];
    const long long n_nan = (long long)s_sum[3];
    const long long Nk = (long long)ws.type_cnt[k];
    const long long nvalid = Nk - n_nan;
    const long long nmid = s_goff[MED_G];
    T med;
    if (nvalid <= 0) {
        med = (T)NAN;
    } else {
        const long long r0 = (nvalid - 1) / 2, r1 = nvalid / 2;
        const T lo = ws.piv[2 * kd], hi = ws.piv[2 * kd + 1];
        // where does each rank land?  0: < lo (fail) 1: == lo 2: candidates 3: == hi 4: beyond (fail)
        auto region = [&](long long r, long long &rin) -> int {
            if (r < below) return 0;
            r -= below;
            if (r < eq_lo) return 1;
            r -= eq_lo;
            if (r < nmid) { rin = r; return 2; }
            r -= nmid;
            if (r < eq_hi) return 3;
            return 4;
        };
        long long i0 = 0, i1 = 0;
        const int g0 = region(r0, i0), g1 = region(r1, i1);
        T v0, v1;
        if (overflow || g0 == 0 || g0 == 4 || g1 == 0 || g1 == 4) {
            // exact fallback: radix select over the whole column of this type
            if (threadIdx.x == 0) atomicAdd(ws.fail, 1u);
            auto col = [&](auto f) {
                for (long long i = threadIdx.x; i < n; i += blockDim.x)
                    if (__ldg(code + i) == k) {
                        const T x = X[i * ldx + d];
                        if (x == x) f(x);
                    }
            };
            cta_select2_raw<T>(col, r0, r1, hist, sh, (Key)0, BITS - 8, v0, v1);
        } else {
            T m0 = lo, m1 = lo;
            if (g0 == 2 || g1 == 2) {
                const T *cl = ws.cand + ws.coff[kd];
                // gather the MED_G sub-lists into shared memory once (all loads in flight together);
                // longer lists are read from global memory on every pass
                const bool staged = nmid <= MED_STAGE;
                if (staged) {
                    for (int idx = threadIdx.x; idx < MED_G * 32; idx += blockDim.x) {
                        const int g = idx >> 5, l5 = idx & 31;
                        const unsigned cg = s_goff[g + 1] - s_goff[g];
                        const T *seg = cl + (unsigned long long)g * capv;
                        for (unsigned i = l5; i < cg; i += 32) s_stage[s_goff[g] + i] = seg[i];
                    }
                    __syncthreads();
                }
                auto lst = [&](auto f) {
                    if (staged) {
                        for (unsigned i = threadIdx.x; i < (unsigned)nmid; i += blockDim.x) f(s_stage[i]);
                    } else {
                        for (int g = 0; g < MED_G; ++g) {
                            const unsigned cg = s_goff[g + 1] - s_goff[g];
                            const T *seg = cl + (unsigned long long)g * capv;
                            for (unsigned i = threadIdx.x; i < cg; i += blockDim.x) f(seg[i]);
                        }
                    }
                };
                // every candidate lies strictly between lo and hi: the digits above the first one in
                // which key(lo) and key(hi) differ are common to all of them
                Key prefix = 0;
                int first = BITS - 8;
                if (lo > -Inf<T>::pos() && hi < Inf<T>::pos()) {
                    const Key kl = KO::key(lo), kh = KO::key(hi);
                    while (first > 0 && (kl >> first) == (kh >> first)) first -= 8;
                    if (first < BITS - 8) prefix = (Key)((kl >> (first + 8)) << (first + 8));
                }
                cta_select2_raw<T>(lst, g0 == 2 ? i0 : 0, g1 == 2 ? i1 : 0, hist, sh, prefix, first, m0, m1);
            }
            v0 = g0 == 1 ? lo : (g0 == 3 ? hi : m0);
            v1 = g1 == 1 ? lo : (g1 == 3 ? hi : m1);
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

static int median_stream_warps(int K, int D, size_t elt)
{
    const size_t kd = (size_t)K * D;
    const size_t fixed = kd * 2 * elt + kd * sizeof(unsigned int), tab = kd * sizeof(unsigned int);
    const size_t budget = 208 * 1024;
    if (fixed + tab > budget) return 0;
    size_t w = (budget - fixed) / tab;
    return (int)(w > 32 ? 32 : w);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static unsigned int median_log_cap(long long n, int D, long long TW)
{
    // expected ~10 % of a warp's elements are candidates; allow 3.5x that plus slack
    const double per_warp = (double)n * D / (double)TW;
    double c = 0.35 * per_warp + 2048.0;
    if (c > 4.0e9) c = 4.0e9;
    return (unsigned int)c;
}

template <typename T>
static size_t median_ws2_bytes(long long n, int K, int D, int ctasB, int W)
{
    const size_t kd = (size_t)K * D;
    const long long TW = (long long)ctasB * W;
    const size_t lc = median_log_cap(n, D, TW);
    size_t b = 0;
    b += align256((size_t)K * 8 + 8);              // type_cnt + cand_total
    b += align256(kd * MED_G * 4 + 4);             // ccnt (MED_G sub-lists) + fail
    b += align256((size_t)K * 4 + 4);              // scnt + log_flag
    b += align256((size_t)TW * 4);                 // log_cnt
    b += align256(kd * 4);                         // cap
    b += align256(kd * 8);                         // coff
    b += align256(kd * 2 * sizeof(T));             // piv
    b += align256((size_t)K * MED_SCAP * 4);       // slist
    b += align256((size_t)ctasB * kd * 4 * 4);     // partial: 4 counters per pair and CTA
    b += align256((size_t)TW * lc * 4);            // log_kd
    b += align256((size_t)TW * lc * sizeof(T));    // log_val
    b += align256(((size_t)(0.3 * (double)n * D) + kd * 256 * MED_G) * sizeof(T));  // candidates
    return b;
}

template <typename T>
static int median_run_sampled(const T *X, long long n, int D, long long ldx, const int *code, int K, T *cent,
                              double *cent64, void *workspace, int W, cudaStream_t st)
{
    const size_t kd = (size_t)K * D;
    const int ctasB = sm_count();
    const long long TW = (long long)ctasB * W;
    MedianWs2<T> ws;
    unsigned char *p = (unsigned char *)workspace;
    ws.type_cnt = (unsigned long long *)p; ws.cand_total = ws.type_cnt + K; p += align256((size_t)K * 8 + 8);
    ws.ccnt = (unsigned int *)p; ws.fail = ws.ccnt + kd * MED_G; p += align256(kd * MED_G * 4 + 4);
    ws.scnt = (unsigned int *)p; ws.log_flag = ws.scnt + K; p += align256((size_t)K * 4 + 4);
    ws.log_cnt = (unsigned int *)p; p += align256((size_t)TW * 4);
    const size_t zero_bytes = (size_t)(p - (unsigned char *)workspace);
    ws.cap = (unsigned int *)p; p += align256(kd * 4);
    ws.coff = (unsigned long long *)p; p += align256(kd * 8);
    ws.piv = (T *)p; p += align256(kd * 2 * sizeof(T));
    ws.slist = (int *)p; p += align256((size_t)K * MED_SCAP * 4);
    ws.partial = (unsigned int *)p; p += align256((size_t)ctasB * kd * 4 * 4);
    ws.log_cap = median_log_cap(n, D, TW);
    ws.log_kd = (int *)p; p += align256((size_t)TW * ws.log_cap * 4);
    ws.log_val = (T *)p; p += align256((size_t)TW * ws.log_cap * sizeof(T));
    ws.cand = (T *)p;
    ws.cand_capacity = (unsigned long long)(0.3 * (double)n * D) + kd * 256 * MED_G;
    PILOT_CUDA(cudaMemsetAsync(workspace, 0, zero_bytes, st));
    const MedianSampling sp = median_sampling(n, K);
    {
        long long blocks = (n + 1023) / 1024;
        const long long cap = (long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        median_count_kernel<<<(unsigned)blocks, 256, K * sizeof(unsigned int), st>>>(code, n, K, sp, ws.type_cnt,
                                                                                    ws.scnt, ws.slist);
        PILOT_LAUNCH_CHECK();
    }
    median_pivot_kernel<T><<<(unsigned)kd, 128, 0, st>>>(X, n, D, ldx, code, K, sp, ws);
    PILOT_LAUNCH_CHECK();
    {
        const size_t smem = kd * 2 * sizeof(T) + (size_t)(W + 1) * kd * sizeof(unsigned int);
        PILOT_CUDA(cudaFuncSetAttribute(median_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        median_stream_kernel<T><<<ctasB, 32 * W, smem, st>>>(X, n, D, ldx, code, K, W, ws);
        PILOT_LAUNCH_CHECK();
    }
    median_bin_kernel<T><<<(unsigned)TW, 256, 0, st>>>(ws);
    PILOT_LAUNCH_CHECK();
    median_finish_kernel<T><<<(unsigned)kd, 128, 0, st>>>(X, n, D, ldx, code, K, ctasB, ws, cent, cent64);
    PILOT_LAUNCH_CHECK();
    return 0;
}

static bool median_use_sampled(long long n, int K, int D, size_t elt, int *W)
{
    *W = median_stream_warps(K, D, elt);
    // per-warp 16-bit counters: a warp must see fewer than 65536 rows
    const long long rows_per_warp = n / ((long long)sm_count() * (*W > 0 ? *W : 1)) + 1;
    return n >= 65536 && *W >= 2 && rows_per_warp < 60000 && D >= 32 && D <= 256;
}

size_t median_ws_bytes(long long n, int K, int D)
{
    size_t a = median_ws_bytes_impl(K, D);
    size_t b = 0;
    for (int elt = 4; elt <= 8; elt += 4) {
        const int W = median_stream_warps(K, D, elt);
        if (W >= 2) {
            const size_t c = elt == 4 ? median_ws2_bytes<float>(n, K, D, sm_count(), W)
                                      : median_ws2_bytes<double>(n, K, D, sm_count(), W);
            if (c > b) b = c;
        }
    }
    return a > b ? a : b;
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
    int W = 0;
    if (ldx == D && median_use_sampled(n_cells, K, D, dtype == PILOT_F32 ? 4 : 8, &W)) {
        if (dtype == PILOT_F32)
            return median_run_sampled<float>((const float *)X, n_cells, D, ldx, ct_code, K, (float *)centroids,
                                             centroids_f64, workspace, W, st);
        return median_run_sampled<double>((const double *)X, n_cells, D, ldx, ct_code, K, (double *)centroids,
                                          centroids_f64, workspace, W, st);
    }
    if (dtype == PILOT_F32)
        return median_run<float>((const float *)X, n_cells, D, ldx, ct_code, K, (float *)centroids,
                                 centroids_f64, workspace, st);
    return median_run<double>((const double *)X, n_cells, D, ldx, ct_code, K, (double *)centroids,
                              centroids_f64, workspace, st);
}
