// Kernel (3), batched variant: all-pairs stabilised Sinkhorn with ONE shared Gibbs kernel
// K0 = exp(-M/reg) resident in shared memory for every problem of the CTA.
// Replaces the loop over ot.sinkhorn2(a_i, a_j, cost, reg, method="sinkhorn_stabilized")
// (reference pilotpy/tools/Trajectory.py:513-515; schedule: SURVEY.md Appendix A.2).
//
// Formulation.  POT keeps K = diag(e^{alpha/reg}) K0 diag(e^{beta/reg}) per problem and
// iterates v = b/(K^T u), u = a/(K v).  With ut = e^{alpha/reg} o u and vt = e^{beta/reg} o v
// the same iteration is vt = b/(K0^T ut), ut = a/(K0 vt) on the SHARED K0; alpha/beta only
// matter for (i) the absorption trigger max|u|,|v| > tau, u = ut * rea with rea = e^{-alpha/reg}
// (= 1/ut at the last absorption), and (ii) the reset u = v = 1/K at an absorption, which
// in scaled variables is ut /= K, vt /= K.  The marginal error every `check_every` iterations
// is || vt o (K0^T ut) - b ||, i.e. the product the next iteration needs anyway, and the
// result is sum_ij M_ij K0_ij ut_i vt_j.  Rounding differs from the reference form at the
// 1e-16 level with identical iteration/absorption schedules.
//
// Mapping.  The two matvecs for many problems are K0^T [ut_1 .. ut_P] and K0 [vt_1 .. vt_P]:
// register-tiled FP64 GEMMs.  A group of NTJ = KP/8 adjacent lanes owns 4 problem "slots";
// each lane accumulates an 8 (rows) x 4 (slots) tile, A-fragments (K0 rows) and B-fragments
// (slot vectors) come from shared memory as 128-bit loads: 6 LDS.128 per 32 DFMA.  All
// traffic for one slot stays inside its warp, so warps are independent persistent workers
// (no CTA barrier in the loop); a finished slot is refilled at once from a global counter.
// Problems the scaled form cannot represent (NaN, |log ut| near the FP64 range) are queued
// for the reference-form kernel (sinkhorn_ref.cu).
#include "sinkhorn.cuh"

namespace pilot {

constexpr int SKB_WARPS = 8;

template <int KP> struct SkbCfg {
    static constexpr int NTJ = KP / 8;    // lanes per slot group
    static constexpr int GPW = 32 / NTJ;  // groups per warp
    static constexpr int SPW = GPW * 4;   // slots per warp
    static constexpr int PS = SPW;        // row stride (doubles) of the per-warp U / V panels
};

// K0, K0^T, M o K0 (all KP x KP, zero padded) and c0 = K0^T (1/K) into the workspace
__global__ void skb_setup_kernel(const double *__restrict__ M, int K, int KP, double reg,
                                 double *__restrict__ K0, double *__restrict__ K0T,
                                 double *__restrict__ MK, double *__restrict__ c0)
{
    const int n = KP * KP;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = e / KP, j = e - i * KP;
        double k0 = 0.0, mk = 0.0;
        if (i < K && j < K) {
            const double m = M[i * K + j];
            k0 = exp(-m / reg);
            mk = m * k0;
        }
        K0[i * KP + j] = k0;
        K0T[j * KP + i] = k0;
        MK[i * KP + j] = mk;
    }
    if (blockIdx.x == 0)
        for (int j = threadIdx.x; j < KP; j += blockDim.x) {
            double s = 0.0;
            const double u0 = 1.0 / K;
            if (j < K)
                for (int i = 0; i < K; ++i) s += exp(-M[i * K + j] / reg) * u0;
            c0[j] = s;
        }
}

template <int KP>
__global__ void __launch_bounds__(SKB_WARPS * 32, 1)
sinkhorn_batched_kernel(const double *__restrict__ props, int K, SkParams prm, PairMap pm,
                        const double *__restrict__ gK0, const double *__restrict__ gK0T,
                        const double *__restrict__ gMK, const double *__restrict__ gc0,
                        double *__restrict__ scratch,  // [gridDim * WARPS * SPW][2][KP] rea / reb
                        double *__restrict__ out, int *__restrict__ iters_out, int *__restrict__ abs_out,
                        int *__restrict__ status_out, unsigned long long *__restrict__ counter,
                        long long *__restrict__ redo_list, unsigned long long *__restrict__ n_redo)
{
    using C = SkbCfg<KP>;
    constexpr int NTJ = C::NTJ, SPW = C::SPW, PS = C::PS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sK0 = reinterpret_cast<double *>(smem_raw);
    double *sK0T = sK0 + KP * KP;
    double *sc0 = sK0T + KP * KP;
    double *sUV = sc0 + KP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *U = sUV + (size_t)warp * 2 * KP * PS;  // U[k][PS], then V[k][PS]
    double *V = U + KP * PS;
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) { sK0[e] = gK0[e]; sK0T[e] = gK0T[e]; }
    for (int e = threadIdx.x; e < KP; e += blockDim.x) sc0[e] = gc0[e];
    __syncthreads();

    const int tj = lane % NTJ, grp = lane / NTJ;
    const unsigned gmask = (NTJ == 32 ? 0xffffffffu : ((1u << NTJ) - 1u)) << (grp * NTJ);
    const int col0 = grp * 4;  // first slot column of this group inside the warp panel
    double *wscr = scratch + ((size_t)(blockIdx.x * SKB_WARPS + warp) * SPW) * 2 * KP;
    // rows owned by this lane: j(q,h) = 2*NTJ*q + 2*tj + h
    int rowj[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) { rowj[2 * q] = 2 * NTJ * q + 2 * tj; rowj[2 * q + 1] = rowj[2 * q] + 1; }
    const double invK = 1.0 / K;

    // per-slot state, replicated in the NTJ lanes of the group
    long long sl[4];
    int s_i[4], s_j[4], s_ii[4], s_abs[4];
    bool s_act[4], s_hasabs[4], s_pend[4], s_force[4], s_fresh[4], s_bad[4];

    // initial fill
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned long long w = 0;
        if (tj == 0) w = atomicAdd(counter, 1ULL);
        w = __shfl_sync(gmask, w, grp * NTJ);
        sl[c] = (long long)w; s_i[c] = 0; s_j[c] = 0;
        s_act[c] = (long long)w < pm.n_local;
        if (s_act[c]) global_to_ij(pm, local_to_global(pm, sl[c]), s_i[c], s_j[c]);
        s_ii[c] = 0; s_abs[c] = 0; s_hasabs[c] = false; s_pend[c] = false; s_force[c] = false;
        s_fresh[c] = true; s_bad[c] = false;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            U[rowj[r] * PS + col0 + c] = (s_act[c] && rowj[r] < K) ? invK : 0.0;
            V[rowj[r] * PS + col0 + c] = (s_act[c] && rowj[r] < K) ? invK : 0.0;
        }
    }
    __syncwarp();

    for (;;) {
        bool any_act = false;
#pragma unroll
        for (int c = 0; c < 4; ++c) any_act |= s_act[c];
        if (!__any_sync(0xffffffffu, any_act)) break;

        double acc[8][4];
        // ======================= phase A: T = K0^T Ut =======================
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll 4
        for (int k = 0; k < KP; ++k) {
            double a[8], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 t2 = *reinterpret_cast<const double2 *>(sK0 + k * KP + 2 * NTJ * q + 2 * tj);
                a[2 * q] = t2.x; a[2 * q + 1] = t2.y;
            }
            {
                const double2 b0 = *reinterpret_cast<const double2 *>(U + k * PS + col0);
                const double2 b1 = *reinterpret_cast<const double2 *>(U + k * PS + col0 + 2);
                b[0] = b0.x; b[1] = b0.y; b[2] = b1.x; b[3] = b1.y;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }

        // ---- resolve the pending convergence check / iteration cap of the previous iteration ----
        bool stop[4];
        int stop_status[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            stop[c] = false; stop_status[c] = PILOT_ST_MAXITER;
            if (s_act[c] && !s_fresh[c]) {
                if (s_bad[c]) {
                    stop[c] = true; stop_status[c] = -1;  // handed to the reference-form kernel
                } else {
                    bool conv = false;
                    if (s_pend[c]) {
                        double e2 = 0.0;
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if (rowj[r] < K) {
                                const double bj = __ldg(props + (long long)s_j[c] * K + rowj[r]);
                                const double d = fma(V[rowj[r] * PS + col0 + c], acc[r][c], -bj);
                                e2 = fma(d, d, e2);
                            }
#pragma unroll
                        for (int o = 1; o < NTJ; o <<= 1) e2 += __shfl_xor_sync(gmask, e2, o);
                        conv = sqrt(e2) <= prm.stop_thr;
                    }
                    if (conv) { stop[c] = true; stop_status[c] = PILOT_ST_CONVERGED; }
                    else if (s_force[c]) { stop[c] = true; stop_status[c] = PILOT_ST_MAXITER; }
                }
            }
        }
        // ---- warp-cooperative finalisation + refill of the stopped slots ----
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned m = __ballot_sync(0xffffffffu, stop[c]);
            while (m) {
                const int src = __ffs(m) - 1;              // first lane of a stopping group
                const int g = src / NTJ;
                m &= ~(((NTJ == 32) ? 0xffffffffu : ((1u << NTJ) - 1u)) << (g * NTJ));
                const int col = g * 4 + c;
                const int stt = __shfl_sync(0xffffffffu, stop_status[c], src);
                double cost = 0.0;
                if (stt >= 0) {
                    // cost = sum_j Vt_j * sum_i (M o K0)_ij Ut_i ; lane = column j
                    for (int j = lane; j < KP; j += 32) {
                        double wj = 0.0;
                        for (int i = 0; i < K; ++i) wj = fma(__ldg(gMK + i * KP + j), U[i * PS + col], wj);
                        cost = fma(V[j * PS + col], wj, cost);
                    }
                    cost = warp_sum_d(cost);
                }
                unsigned long long w = 0;
                if (lane == 0) w = atomicAdd(counter, 1ULL);
                w = __shfl_sync(0xffffffffu, w, 0);
                const bool mine = grp == g;
                if (mine && tj == 0) {
                    if (stt >= 0) {
                        out[sl[c]] = cost;
                        if (iters_out) iters_out[sl[c]] = s_ii[c];
                        if (abs_out) abs_out[sl[c]] = s_abs[c];
                        if (status_out) status_out[sl[c]] = stt;
                    } else {
                        const unsigned long long slot = atomicAdd(n_redo, 1ULL);
                        if ((long long)slot < SK_REDO_CAP) redo_list[slot] = sl[c];
                        out[sl[c]] = __longlong_as_double(0x7ff8000000000000LL);
                        if (status_out) status_out[sl[c]] = -1;
                    }
                }
                if (mine) {
                    sl[c] = (long long)w;
                    s_act[c] = (long long)w < pm.n_local;
                    if (s_act[c]) global_to_ij(pm, local_to_global(pm, sl[c]), s_i[c], s_j[c]);
                    s_ii[c] = 0; s_abs[c] = 0; s_hasabs[c] = false; s_pend[c] = false; s_force[c] = false;
                    s_fresh[c] = true; s_bad[c] = false;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        U[rowj[r] * PS + col0 + c] = (s_act[c] && rowj[r] < K) ? invK : 0.0;
                        V[rowj[r] * PS + col0 + c] = (s_act[c] && rowj[r] < K) ? invK : 0.0;
                    }
                }
            }
        }
        __syncwarp();

        // ---- v-update: Vt = b / T ----
        double mxv[4];
        bool badv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            mxv[c] = 0.0; badv[c] = false;
            if (s_act[c]) {
                const double *reb = wscr + ((size_t)(col0 + c) * 2 + 1) * KP;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int j = rowj[r];
                    if (j < K) {
                        const double t = s_fresh[c] ? sc0[j] : acc[r][c];
                        const double bj = __ldg(props + (long long)s_j[c] * K + j);
                        const double vn = bj / t;
                        V[j * PS + col0 + c] = vn;
                        const double uu = s_hasabs[c] ? vn * reb[j] : vn;
                        badv[c] |= (vn != vn) || (uu != uu);
                        mxv[c] = fmax(mxv[c], fabs(uu));
                    }
                }
            }
        }
        __syncwarp();

        // ======================= phase B: S = K0 Vt =======================
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll 4
        for (int k = 0; k < KP; ++k) {
            double a[8], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 t2 = *reinterpret_cast<const double2 *>(sK0T + k * KP + 2 * NTJ * q + 2 * tj);
                a[2 * q] = t2.x; a[2 * q + 1] = t2.y;
            }
            {
                const double2 b0 = *reinterpret_cast<const double2 *>(V + k * PS + col0);
                const double2 b1 = *reinterpret_cast<const double2 *>(V + k * PS + col0 + 2);
                b[0] = b0.x; b[1] = b0.y; b[2] = b1.x; b[3] = b1.y;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
        // ---- u-update: Ut = a / S, then the per-slot service (absorption, counters) ----
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (s_act[c]) {
                double *rea = wscr + ((size_t)(col0 + c) * 2) * KP;
                double *reb = rea + KP;
                double mxu = 0.0;
                bool bad = badv[c];
                double un[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = rowj[r];
                    un[r] = 0.0;
                    if (i < K) {
                        const double ai = __ldg(props + (long long)s_i[c] * K + i);
                        un[r] = ai / acc[r][c];
                        const double uu = s_hasabs[c] ? un[r] * rea[i] : un[r];
                        bad |= (un[r] != un[r]) || (uu != uu);
                        mxu = fmax(mxu, fabs(uu));
                    }
                }
                double mxvv = mxv[c];
                int badi = bad ? 1 : 0;
#pragma unroll
                for (int o = 1; o < NTJ; o <<= 1) {
                    mxu = fmax(mxu, __shfl_xor_sync(gmask, mxu, o));
                    mxvv = fmax(mxvv, __shfl_xor_sync(gmask, mxvv, o));
                    badi |= __shfl_xor_sync(gmask, badi, o);
                }
                const bool absorb = !badi && (mxu > prm.tau || mxvv > prm.tau);
                bool range_bad = false;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = rowj[r];
                    if (i < K) {
                        if (absorb) {
                            const double vo = V[i * PS + col0 + c];
                            rea[i] = 1.0 / un[r];
                            reb[i] = 1.0 / vo;
                            // keep e^{+-alpha/reg} comfortably inside the FP64 range
                            range_bad |= !(un[r] > 1e-250 && un[r] < 1e250 && vo > 1e-250 && vo < 1e250);
                            un[r] *= invK;
                            V[i * PS + col0 + c] = vo * invK;
                        }
                        U[i * PS + col0 + c] = un[r];
                    }
                }
                int rb = range_bad ? 1 : 0;
#pragma unroll
                for (int o = 1; o < NTJ; o <<= 1) rb |= __shfl_xor_sync(gmask, rb, o);
                if (absorb) { s_hasabs[c] = true; ++s_abs[c]; }
                s_bad[c] = badi || rb;
                s_pend[c] = (s_ii[c] % prm.check_every) == 0;
                ++s_ii[c];
                s_force[c] = s_ii[c] >= prm.num_iter_max;
                s_fresh[c] = false;
            }
        }
        __syncwarp();
    }
}

size_t skb_setup_bytes(int KP) { return sizeof(double) * ((size_t)3 * KP * KP + KP); }
size_t skb_smem_bytes(int KP)
{
    const int SPW = (32 / (KP / 8)) * 4;
    return sizeof(double) * ((size_t)2 * KP * KP + KP + (size_t)SKB_WARPS * 2 * KP * SPW);
}
size_t skb_scratch_bytes(int KP, int ctas)
{
    const int SPW = (32 / (KP / 8)) * 4;
    return sizeof(double) * (size_t)ctas * SKB_WARPS * SPW * 2 * KP;
}
int skb_pad(int K) { return K <= 16 ? 16 : (K <= 32 ? 32 : 64); }

template <int KP>
static int skb_launch_t(const double *props, int K, const SkParams &prm, const PairMap &pm, const double *setup,
                        double *scratch, int ctas, double *out, int *iters, int *absn, int *status,
                        unsigned long long *counter, long long *redo, unsigned long long *n_redo, cudaStream_t st)
{
    const size_t smem = skb_smem_bytes(KP);
    PILOT_CUDA(cudaFuncSetAttribute(sinkhorn_batched_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double *K0 = setup, *K0T = K0 + KP * KP, *MK = K0T + KP * KP, *c0 = MK + KP * KP;
    sinkhorn_batched_kernel<KP><<<ctas, SKB_WARPS * 32, smem, st>>>(props, K, prm, pm, K0, K0T, MK, c0, scratch, out,
                                                                   iters, absn, status, counter, redo, n_redo);
    PILOT_LAUNCH_CHECK();
    return 0;
}

int skb_launch(const double *props, int K, const double *cost, const SkParams &prm, const PairMap &pm,
               double *setup, double *scratch, int ctas, double *out, int *iters, int *absn, int *status,
               unsigned long long *counter, long long *redo, unsigned long long *n_redo, cudaStream_t st)
{
    const int KP = skb_pad(K);
    skb_setup_kernel<<<8, 256, 0, st>>>(cost, K, KP, prm.reg, setup, setup + KP * KP, setup + 2 * KP * KP,
                                        setup + 3 * KP * KP);
    PILOT_LAUNCH_CHECK();
    if (KP == 16) return skb_launch_t<16>(props, K, prm, pm, setup, scratch, ctas, out, iters, absn, status, counter, redo, n_redo, st);
    if (KP == 32) return skb_launch_t<32>(props, K, prm, pm, setup, scratch, ctas, out, iters, absn, status, counter, redo, n_redo, st);
    return skb_launch_t<64>(props, K, prm, pm, setup, scratch, ctas, out, iters, absn, status, counter, redo, n_redo, st);
}

}  // namespace pilot
