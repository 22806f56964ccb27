// Kernel (3), warp-specialised variant of the batched shared-Gibbs-kernel Sinkhorn solver
// (see sinkhorn_batched.cu for the formulation; reference call site pilotpy/tools/Trajectory.py:513-515).
//
// The single-role kernel leaves the FP64 tensor pipe idle ~43 % of the time: its four warps per
// scheduler alternate between a DMMA phase and a long latency-bound element-wise phase, and all four
// are regularly in the element-wise phase at once.  Here the roles are split like in a Blackwell
// GEMM: 4 MMA warps (one per scheduler) do nothing but stream matvecs -- B fragments from a slot
// set's panel, A fragments from the shared K0 / K0^T, accumulators written to the set's result
// buffer -- while 12 epilogue warps each own one slot set (8 problems) and do the division, the
// reductions, absorption, convergence check, final cost and refill.  A set ping-pongs between its
// MMA warp and its epilogue warp through a shared-memory state word (U_READY -> T_READY -> V_READY
// -> S_READY -> U_READY ...); every MMA warp serves three sets, so the tensor pipe always has a
// matvec to run while the other sets are in their epilogues.
#include "sinkhorn.cuh"

namespace pilot {

constexpr int SKW_MMA_WARPS = 4;
constexpr int SKW_SETS = 12;                       // one epilogue warp per set
constexpr int SKW_WARPS = SKW_MMA_WARPS + SKW_SETS;
constexpr int SKW_SPW = 8;                         // slots per set

enum { SKW_U_READY = 1, SKW_T_READY = 2, SKW_V_READY = 3, SKW_S_READY = 4, SKW_DONE = 5 };

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// branch-free x / y for finite positive y.  FP64-pipe instructions are what the element-wise phases
// pay for (they queue behind the DMMAs): 4 here.
__device__ __forceinline__ double fast_div(double x, double y)
{
    // 20-bit seed r, e = 1 - y r, q = x r (1 + e + e^2): relative error e^3 ~ 2^-60 before the final
    // rounding; 4 FP64-pipe instructions in a dependent chain of 3
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    const double e = fma(-y, r, 1.0);
    const double q0 = x * r;
    return fma(q0, fma(e, e, e), q0);
}

// |x| as an ordered integer: for non-NaN doubles the bit pattern orders like the value, NaN / Inf
// sort above every finite value -- max / threshold tests without touching the FP64 pipe
__device__ __forceinline__ long long abs_bits(double x) { return __double_as_longlong(x) & 0x7fffffffffffffffLL; }

// swizzled matrix addressing: column ^ 4 * (row % 4).  An A-fragment load reads rows 4*ks + t
// (t = 0..3) at columns 8*m + g; 64-bit shared loads are served per half-warp (g = 0..3 / 4..7), and
// with this swizzle the four rows of a half-warp fall into four different 8-bank groups: 2
// wavefronts per load, the minimum (an 8-column swizzle on odd rows gave 4).
__device__ __forceinline__ int swz(int row, int col) { return col ^ ((row & 3) << 2); }

__device__ __forceinline__ int skw_load_flag(const volatile int *f) { return *f; }

template <int KP, bool FULL, bool SYM>
__global__ void __launch_bounds__(SKW_WARPS * 32, 1)
sinkhorn_ws_kernel(const double *__restrict__ props, int K, SkParams prm, PairMap pm, int slot_cap, int set_cap,
                   const double *__restrict__ gK0, const double *__restrict__ gK0T,
                   const double *__restrict__ gMK, const double *__restrict__ gc0,
                   double *__restrict__ scratch,  // [gridDim * SETS * 8][2][KP] rea / reb
                   double *__restrict__ out, int *__restrict__ iters_out, int *__restrict__ abs_out,
                   int *__restrict__ status_out, unsigned long long *__restrict__ counter,
                   long long *__restrict__ redo_list, unsigned long long *__restrict__ n_redo)
{
    if ((reinterpret_cast<const int *>(gc0 + KP)[0] != 0) == SYM) return;  // see sinkhorn_batched.cu
    constexpr int MT = KP / 8;   // accumulator row tiles
    constexpr int KS = KP / 4;   // k-steps of 4
    constexpr int PS = SKW_SPW;  // panel row stride (doubles)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool sym = SYM;
    double *sK0 = reinterpret_cast<double *>(smem_raw);  // [i][swz(i, j)]
    double *sSecond = sK0 + KP * KP;                     // K0^T (swizzled), or M o K0 (plain) when symmetric
    double *sK0T = sym ? sK0 : sSecond;
    double *sc0 = sSecond + KP * KP;
    double *sPan = sc0 + KP;                             // [set][U, V][KP][8]
    double *sRes = sPan + (size_t)SKW_SETS * 2 * KP * PS;  // [set][KP][8]
    constexpr bool STAGE_AB = KP <= 32;
    double *sAB = sRes + (size_t)SKW_SETS * KP * PS;     // [set][a, b][8][KP]  (small K only)
    volatile int *flags = reinterpret_cast<volatile int *>(sAB + (STAGE_AB ? (size_t)SKW_SETS * 2 * SKW_SPW * KP : 0));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
        const int r = e / KP, c = e - r * KP;
        sK0[r * KP + swz(r, c)] = gK0[e];
        if (sym) sSecond[e] = gMK[e];
        else sSecond[r * KP + swz(r, c)] = gK0T[e];
    }
    for (int e = threadIdx.x; e < KP; e += blockDim.x) sc0[e] = gc0[e];
    if (threadIdx.x < SKW_SETS) flags[threadIdx.x] = 0;
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int tsw = (t & 2) << 2, gsw = g ^ ((t & 1) << 2);  // K0 fragment rows 4*ks + t: swizzle 4 * t

    if (warp < SKW_MMA_WARPS) {
        // ============================ MMA warp ============================
        for (;;) {
            int n_done = 0;
            bool worked = false;
#pragma unroll 1
            for (int q = 0; q < SKW_SETS / SKW_MMA_WARPS; ++q) {
                const int s = warp + SKW_MMA_WARPS * q;
                const int stt = skw_load_flag(flags + s);
                if (stt == SKW_DONE) { ++n_done; continue; }
                if (stt != SKW_U_READY && stt != SKW_V_READY) continue;
                __threadfence_block();
                const bool phaseA = stt == SKW_U_READY;
                const double *B = sPan + ((size_t)s * 2 + (phaseA ? 0 : 1)) * KP * PS;
                const double *A = phaseA ? sK0 : sK0T;
                double acc[MT][2];
#pragma unroll
                for (int m = 0; m < MT; ++m) { acc[m][0] = 0.0; acc[m][1] = 0.0; }
                // software-pipelined: the fragments of k-step ks + 1 are in flight while the 8 DMMAs
                // of step ks run (this warp is alone on its scheduler's tensor pipe, nobody else
                // hides its shared-memory latency)
                const double *Bp = B + t * PS + g;
                const double *Ap = A + t * KP + gsw;
                double an[MT], bn = Bp[0];
#pragma unroll
                for (int m = 0; m < MT; ++m) an[m] = Ap[(8 * m) ^ tsw];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    double ac[MT];
                    const double bc = bn;
#pragma unroll
                    for (int m = 0; m < MT; ++m) ac[m] = an[m];
                    if (ks + 1 < KS) {
                        bn = Bp[4 * (ks + 1) * PS];
#pragma unroll
                        for (int m = 0; m < MT; ++m) an[m] = Ap[4 * (ks + 1) * KP + ((8 * m) ^ tsw)];
                    }
#pragma unroll
                    for (int m = 0; m < MT; ++m) dmma884(acc[m][0], acc[m][1], ac[m], bc);
                }
                double *R = sRes + (size_t)s * KP * PS + g * PS + 2 * t;
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    *reinterpret_cast<double2 *>(R + 8 * m * PS) = make_double2(acc[m][0], acc[m][1]);
                __threadfence_block();
                __syncwarp();
                if (lane == 0) flags[s] = phaseA ? SKW_T_READY : SKW_S_READY;
                worked = true;
            }
            if (n_done == SKW_SETS / SKW_MMA_WARPS) break;
            if (!worked) __nanosleep(40);
        }
        return;
    }

    // ============================ epilogue warp: owns slot set `set` ============================
    const int set = warp - SKW_MMA_WARPS;
    volatile int *myflag = flags + set;
    if (set >= set_cap) {
        if (lane == 0) *myflag = SKW_DONE;
        return;
    }
    double *U = sPan + (size_t)set * 2 * KP * PS;  // U[row][slot], then V
    double *V = U + KP * PS;
    const double *Rm = sRes + (size_t)set * KP * PS + g * PS + 2 * t;
    double *sA = sAB + (size_t)set * 2 * SKW_SPW * KP;  // [slot][KP], then b
    double *sB = sA + SKW_SPW * KP;
    const unsigned gmask = 0x11111111u << t;  // the 8 lanes that share my 2 slots
    double *wscr = scratch + ((size_t)(blockIdx.x * SKW_SETS + set) * SKW_SPW) * 2 * KP;
    double *rea0 = wscr + (size_t)(2 * t) * 2 * KP;  // slot 2t: rea, then reb; slot 2t+1 follows
    const double invK = 1.0 / K;
    double *Uc = U + g * PS + 2 * t;  // my first element of U; rows advance by 8 * PS
    double *Vc = V + g * PS + 2 * t;

    // per-slot state (h = 0, 1), replicated in the 8 lanes of the group
    long long sl[2];
    const double *pa[2], *pb[2];
    int s_ii[2], s_abs[2];
    bool s_act[2], s_hasabs[2], s_pend[2], s_force[2], s_fresh[2], s_bad[2];

#define ROW_OK(r) (FULL || (r) < K)
#define SKW_ASSIGN(h, w)                                                                    \
    do {                                                                                    \
        sl[h] = (long long)(w);                                                             \
        s_act[h] = (w) != ~0ULL && (long long)(w) < pm.n_local;                             \
        int si_ = 0, sj_ = 0;                                                               \
        if (s_act[h]) global_to_ij(pm, local_to_global(pm, sl[h]), si_, sj_);               \
        pa[h] = props + (long long)si_ * K + g;                                             \
        pb[h] = props + (long long)sj_ * K + g;                                             \
        s_ii[h] = 0; s_abs[h] = 0; s_hasabs[h] = false; s_pend[h] = false; s_force[h] = false; \
        s_fresh[h] = true; s_bad[h] = false;                                                \
        _Pragma("unroll") for (int m = 0; m < MT; ++m) {                                    \
            const bool rok = s_act[h] && ROW_OK(8 * m + g);                                 \
            const double v0 = rok ? invK : 0.0;                                             \
            Uc[8 * m * PS + h] = v0;                                                        \
            Vc[8 * m * PS + h] = v0;                                                        \
            if (STAGE_AB) {                                                                 \
                sA[(2 * t + h) * KP + 8 * m + g] = rok ? __ldg(pa[h] + 8 * m) : 0.0;        \
                sB[(2 * t + h) * KP + 8 * m + g] = rok ? __ldg(pb[h] + 8 * m) : 0.0;        \
            }                                                                               \
        }                                                                                   \
    } while (0)
#define SKW_POST(val)                                                                       \
    do {                                                                                    \
        __threadfence_block();                                                              \
        __syncwarp();                                                                       \
        if (lane == 0) *myflag = (val);                                                     \
    } while (0)
#define SKW_WAIT(val)                                                                       \
    do {                                                                                    \
        while (skw_load_flag(myflag) != (val)) __nanosleep(64);                             \
        __threadfence_block();                                                              \
    } while (0)

    // initial fill
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        unsigned long long w = ~0ULL;
        if (g == 0 && h < slot_cap) w = atomicAdd(counter, 1ULL);
        w = __shfl_sync(gmask, w, t);
        SKW_ASSIGN(h, w);
    }
    if (!__any_sync(0xffffffffu, s_act[0] || s_act[1])) {
        if (lane == 0) *myflag = SKW_DONE;
        return;
    }
    SKW_POST(SKW_U_READY);

    for (;;) {
        double acc[MT][2];
        double num[MT][2];
        // ======================= after phase A: T = K0^T Ut is in the result buffer =======================
        SKW_WAIT(SKW_T_READY);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const double2 r2 = *reinterpret_cast<const double2 *>(Rm + 8 * m * PS);
            acc[m][0] = r2.x; acc[m][1] = r2.y;
        }
        // numerators of the v-update: b of my two slots
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int m = 0; m < MT; ++m)
                num[m][h] = STAGE_AB ? sB[(2 * t + h) * KP + 8 * m + g]
                                     : ((s_act[h] && ROW_OK(8 * m + g)) ? __ldg(pb[h] + 8 * m) : 0.0);

        // ---- resolve the pending convergence check / iteration cap of the previous iteration ----
        bool stop[2];
        int stop_status[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            stop[h] = false; stop_status[h] = PILOT_ST_MAXITER;
            if (s_act[h] && !s_fresh[h]) {
                if (s_bad[h]) {
                    stop[h] = true; stop_status[h] = -1;  // handed to the reference-form kernel
                } else {
                    bool conv = false;
                    if (s_pend[h]) {
                        double e2 = 0.0;
#pragma unroll
                        for (int m = 0; m < MT; ++m)
                            if (ROW_OK(8 * m + g)) {
                                const double d = fma(Vc[8 * m * PS + h], acc[m][h], -num[m][h]);
                                e2 = fma(d, d, e2);
                            }
                        e2 += __shfl_xor_sync(gmask, e2, 4);
                        e2 += __shfl_xor_sync(gmask, e2, 8);
                        e2 += __shfl_xor_sync(gmask, e2, 16);
                        conv = sqrt(e2) <= prm.stop_thr;
                    }
                    if (conv) { stop[h] = true; stop_status[h] = PILOT_ST_CONVERGED; }
                    else if (s_force[h]) { stop[h] = true; stop_status[h] = PILOT_ST_MAXITER; }
                }
            }
        }
        // ---- warp-cooperative finalisation + refill of the stopped slots ----
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            unsigned mball = __ballot_sync(0xffffffffu, stop[h]) & 0xfu;  // one bit per group (lanes 0..3)
            while (mball) {
                const int tt = __ffs(mball) - 1;  // group id == its t
                mball &= mball - 1;
                const int col = 2 * tt + h;
                const int stt = __shfl_sync(0xffffffffu, stop_status[h], tt);
                double cost = 0.0;
                if (stt >= 0) {
                    // cost = sum_j Vt_j * sum_i (M o K0)_ij Ut_i ; lane = column j
                    const double *mk = sym ? sSecond : gMK;
                    for (int j = lane; j < KP; j += 32) {
                        double w0 = 0.0, w1 = 0.0;
                        for (int i = 0; i + 1 < K; i += 2) {
                            w0 = fma(mk[i * KP + j], U[i * PS + col], w0);
                            w1 = fma(mk[(i + 1) * KP + j], U[(i + 1) * PS + col], w1);
                        }
                        if (K & 1) w0 = fma(mk[(K - 1) * KP + j], U[(K - 1) * PS + col], w0);
                        cost = fma(V[j * PS + col], w0 + w1, cost);
                    }
                    cost = warp_sum_d(cost);
                }
                unsigned long long w = 0;
                if (lane == 0) w = atomicAdd(counter, 1ULL);
                w = __shfl_sync(0xffffffffu, w, 0);
                const bool mine = t == tt;
                __syncwarp();  // every lane has read the slot's U, V columns before its owners refill them
                if (mine && g == 0) {
                    if (stt >= 0) {
                        out[sl[h]] = cost;
                        if (iters_out) iters_out[sl[h]] = s_ii[h];
                        if (abs_out) abs_out[sl[h]] = s_abs[h];
                        if (status_out) status_out[sl[h]] = stt;
                    } else {
                        const unsigned long long slot = atomicAdd(n_redo, 1ULL);
                        if ((long long)slot < SK_REDO_CAP) redo_list[slot] = sl[h];
                        out[sl[h]] = __longlong_as_double(0x7ff8000000000000LL);
                        if (status_out) status_out[sl[h]] = -1;
                    }
                }
                if (mine) {
                    SKW_ASSIGN(h, w);
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        num[m][h] = (s_act[h] && ROW_OK(8 * m + g)) ? __ldg(pb[h] + 8 * m) : 0.0;
                }
            }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, s_act[0] || s_act[1])) {
            SKW_POST(SKW_DONE);
            break;
        }

        // ---- v-update: Vt = b / T (my two slots are adjacent: one 128-bit store per row) ----
        long long mxv[2] = {0, 0};
        if (s_act[0] || s_act[1]) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int row = 8 * m + g;
                double vv[2] = {0.0, 0.0};
                if (ROW_OK(row)) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (s_act[h]) {
                            const double tv = s_fresh[h] ? sc0[row] : acc[m][h];
                            vv[h] = fast_div(num[m][h], tv);
                            const double uu = s_hasabs[h] ? vv[h] * rea0[(2 * h + 1) * KP + row] : vv[h];
                            mxv[h] = max(mxv[h], abs_bits(uu));
                        }
                }
                *reinterpret_cast<double2 *>(Vc + 8 * m * PS) = make_double2(vv[0], vv[1]);
            }
        }
        SKW_POST(SKW_V_READY);

        // numerators of the u-update while the MMA warp runs phase B
        if (s_act[0] || s_act[1]) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    num[m][h] = STAGE_AB ? sA[(2 * t + h) * KP + 8 * m + g]
                                         : ((s_act[h] && ROW_OK(8 * m + g)) ? __ldg(pa[h] + 8 * m) : 0.0);
        }
        // ======================= after phase B: S = K0 Vt is in the result buffer =======================
        SKW_WAIT(SKW_S_READY);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const double2 r2 = *reinterpret_cast<const double2 *>(Rm + 8 * m * PS);
            acc[m][0] = r2.x; acc[m][1] = r2.y;
        }
        // ---- u-update: Ut = a / S, then the per-slot service (absorption, counters) ----
        if (s_act[0] || s_act[1]) {
            double un[MT][2];
            long long mxu[2] = {0, 0};
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const int row = 8 * m + g;
                    un[m][h] = 0.0;
                    if (s_act[h] && ROW_OK(row)) {
                        un[m][h] = fast_div(num[m][h], acc[m][h]);
                        const double uu = s_hasabs[h] ? un[m][h] * rea0[(2 * h) * KP + row] : un[m][h];
                        mxu[h] = max(mxu[h], abs_bits(uu));
                    }
                }
            bool absorb[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                long long mx = max(mxu[h], mxv[h]);
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) mx = max(mx, __shfl_xor_sync(gmask, mx, o));
                const bool bad = mx >= 0x7ff0000000000000LL;  // NaN or Inf anywhere in u, v
                absorb[h] = s_act[h] && !bad && mx > __double_as_longlong(prm.tau);
                if (s_act[h]) {
                    s_bad[h] = bad;
                    s_pend[h] = (s_ii[h] % prm.check_every) == 0;
                    ++s_ii[h];
                    s_force[h] = s_ii[h] >= prm.num_iter_max;
                    s_fresh[h] = false;
                }
            }
            if (absorb[0] || absorb[1]) {
                // u = v = 1/K in POT == divide the scaled iterates by K; remember 1/ut, 1/vt
                int rb[2] = {0, 0};
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const int row = 8 * m + g;
                    double2 vo = *reinterpret_cast<double2 *>(Vc + 8 * m * PS);
                    if (ROW_OK(row)) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if (absorb[h]) {
                                double &vref = h ? vo.y : vo.x;
                                rea0[(2 * h) * KP + row] = 1.0 / un[m][h];
                                rea0[(2 * h + 1) * KP + row] = 1.0 / vref;
                                // keep e^{+-alpha/reg} comfortably inside the FP64 range
                                rb[h] |= !(un[m][h] > 1e-250 && un[m][h] < 1e250 && vref > 1e-250 && vref < 1e250);
                                un[m][h] *= invK;
                                vref *= invK;
                            }
                    }
                    *reinterpret_cast<double2 *>(Vc + 8 * m * PS) = vo;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (absorb[h]) {
                        int r = rb[h];
                        r |= __shfl_xor_sync(gmask, r, 4);
                        r |= __shfl_xor_sync(gmask, r, 8);
                        r |= __shfl_xor_sync(gmask, r, 16);
                        s_hasabs[h] = true;
                        ++s_abs[h];
                        s_bad[h] = s_bad[h] || r;
                    }
            }
#pragma unroll
            for (int m = 0; m < MT; ++m)
                *reinterpret_cast<double2 *>(Uc + 8 * m * PS) = make_double2(un[m][0], un[m][1]);
        }
        SKW_POST(SKW_U_READY);
    }
#undef SKW_ASSIGN
#undef SKW_POST
#undef SKW_WAIT
#undef ROW_OK
}

size_t skw_smem_bytes(int KP)
{
    const size_t stage = KP <= 32 ? (size_t)SKW_SETS * 2 * SKW_SPW * KP : 0;
    return sizeof(double) * ((size_t)2 * KP * KP + KP + (size_t)SKW_SETS * 3 * KP * SKW_SPW + stage) + 64;
}
size_t skw_scratch_bytes(int KP, int ctas) { return sizeof(double) * (size_t)ctas * SKW_SETS * SKW_SPW * 2 * KP; }
int skw_slots_per_set() { return SKW_SPW; }
int skw_sets() { return SKW_SETS; }

template <int KP, bool FULL, bool SYM>
static int skw_launch_t(const double *props, int K, const SkParams &prm, const PairMap &pm, int slot_cap, int set_cap,
                        const double *setup, double *scratch, int ctas, double *out, int *iters, int *absn,
                        int *status, unsigned long long *counter, long long *redo, unsigned long long *n_redo,
                        cudaStream_t st)
{
    const size_t smem = skw_smem_bytes(KP);
    PILOT_CUDA(cudaFuncSetAttribute(sinkhorn_ws_kernel<KP, FULL, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    const double *K0 = setup, *K0T = K0 + KP * KP, *MK = K0T + KP * KP, *c0 = MK + KP * KP;
    sinkhorn_ws_kernel<KP, FULL, SYM><<<ctas, SKW_WARPS * 32, smem, st>>>(
        props, K, prm, pm, slot_cap, set_cap, K0, K0T, MK, c0, scratch, out, iters, absn, status, counter, redo, n_redo);
    PILOT_LAUNCH_CHECK();
    return 0;
}

// `setup` must already hold K0, K0^T, M o K0, c0 and the asymmetry flag (skb_setup_kernel)
int skw_launch(const double *props, int K, const SkParams &prm, const PairMap &pm, double *setup, double *scratch,
               int ctas, int slot_cap, int set_cap, bool symmetric, double *out, int *iters, int *absn, int *status,
               unsigned long long *counter, long long *redo, unsigned long long *n_redo, cudaStream_t st)
{
    const int KP = skb_pad(K);
#define SKW_GO(KPV, FULLV)                                                                                        \
    do {                                                                                                          \
        if (symmetric)                                                                                            \
            return skw_launch_t<KPV, FULLV, true>(props, K, prm, pm, slot_cap, set_cap, setup, scratch, ctas, out, \
                                                  iters, absn, status, counter, redo, n_redo, st);                 \
        return skw_launch_t<KPV, FULLV, false>(props, K, prm, pm, slot_cap, set_cap, setup, scratch, ctas, out,    \
                                               iters, absn, status, counter, redo, n_redo, st);                    \
    } while (0)
    if (KP == 16) { if (K == 16) SKW_GO(16, true); SKW_GO(16, false); }
    if (KP == 32) { if (K == 32) SKW_GO(32, true); SKW_GO(32, false); }
    if (KP == 48) { if (K == 48) SKW_GO(48, true); SKW_GO(48, false); }
    if (K == 64) SKW_GO(64, true);
    SKW_GO(64, false);
#undef SKW_GO
}

}  // namespace pilot
