// Kernel (4): all-pairs exact EMD.  Replaces the loop over ot.emd2(a_i, a_j, cost)
// (reference pilotpy/tools/Trajectory.py:507-511; POT emd_c / LEMON network simplex).
//
// One warp per problem, persistent warps pulling problems from a global counter.
// Algorithm: primal network simplex on the K x K transportation graph, re-designed so
// that every tree operation is a constant-depth warp-parallel step instead of the
// pointer chasing of the CPU reference:
//   * the basis tree is stored as parent[] + flow[] + pi[] + a SUBTREE BITMASK per
//     node (<=128 nodes -> 4 x 32-bit words; node x = slot*32 + lane is owned by `lane`)
//   * join / cycle:   node y is on the i-side of the cycle iff sub[y] has i but not j
//                     (one AND per node, no walk)
//   * ratio test:     REDUX min over the owned decreasing arcs, Cunningham tie-break
//                     by cycle order (= subtree popcount)
//   * re-hanging:     stem nodes find their stem child through a scatter, then
//                     sub[] of stem / ancestors are fixed with AND/OR of the cut mask
//   * potentials:     the cut subtree IS a bitmask -> one predicated add per node
//   * pricing:        block search over BR rows x all columns (lane = column), shared
//                     cost matrix in shared memory for all warps of the CTA
//   * start basis:    diagonal arcs (i,i) with min(a_i,b_i) + north-west corner on the
//                     residuals -> no artificial arcs, |pi| = O(max cost), ~half the pivots
// Any exact solver returns the same optimum; parity with the
// oracle is <1e-12 relative.
#include "common.cuh"

namespace pilot {

constexpr int EMD_WARPS = 16;
constexpr int EMD_BLOCK_ROWS = 4;

template <int NW> struct EmdSmem {
    static constexpr int KP = 32 * NW;   // padded types per side
    static constexpr int N = 2 * KP;     // nodes
    static constexpr int NS = 2 * NW;    // slots per lane == mask words
    unsigned int sub[N][NS];
    double flow[N];
    double pi[N];
    double amt[KP];
    unsigned char parent[N];
    unsigned char tmpc[N];
    unsigned char ord[N];
};

__device__ __forceinline__ int pop_lowest(unsigned (&m)[2], int nw)
{
    for (int w = 0; w < nw; ++w)
        if (m[w]) {
            int b = __ffs(m[w]) - 1;
            m[w] &= m[w] - 1;
            return w * 32 + b;
        }
    return -1;
}
__device__ __forceinline__ bool any_left(const unsigned (&m)[2], int nw)
{
    unsigned r = 0;
    for (int w = 0; w < nw; ++w) r |= m[w];
    return r != 0;
}

template <int NW>
__global__ void __launch_bounds__(EMD_WARPS * 32, NW == 2 ? 2 : 3)
emd_pairs_kernel(const double *__restrict__ props, int K, const double *__restrict__ cost, PairMap pm,
                 long long max_pivots, double *__restrict__ out, int *__restrict__ status,
                 int *__restrict__ pivots_out, unsigned long long *__restrict__ counter)
{
    using SM = EmdSmem<NW>;
    constexpr int KP = SM::KP, NS = SM::NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sM = reinterpret_cast<double *>(smem_raw);  // KP x KP, row-major, zero padded
    SM *sw = reinterpret_cast<SM *>(smem_raw + sizeof(double) * KP * KP) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;

    for (int t = threadIdx.x; t < KP * KP; t += blockDim.x) {
        const int i = t / KP, j = t - i * KP;
        sM[t] = (i < K && j < K) ? cost[i * K + j] : 0.0;
    }
    __syncthreads();

    const double EPS = 8.8817841970012523e-15;  // 40 ulp, relative to max(|pi_i|,|pi_j|,|c|)

    for (;;) {
        unsigned long long l = 0;
        if (lane == 0) l = atomicAdd(counter, 1ULL);
        l = __shfl_sync(0xffffffffu, l, 0);
        if ((long long)l >= pm.n_local) break;
        int si_, sj_;
        global_to_ij(pm, local_to_global(pm, (long long)l), si_, sj_);
        const double *pa = props + (long long)si_ * K, *pb = props + (long long)sj_ * K;

        // ---------------- load a, b ; emd2's rescale of b ----------------
        double av[NW], bv[NW];
        bool valid[NW];
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int c = 0; c < NW; ++c) {
            const int idx = lane + 32 * c;
            valid[c] = idx < K;
            av[c] = valid[c] ? pa[idx] : 0.0;
            bv[c] = valid[c] ? pb[idx] : 0.0;
            sa += av[c];
            sb += bv[c];
        }
        sa = warp_sum_d(sa);
        sb = warp_sum_d(sb);
#pragma unroll
        for (int c = 0; c < NW; ++c) bv[c] = __ddiv_rn(__dmul_rn(bv[c], sa), sb);

        // ---------------- start basis: parallel part ----------------
        unsigned surm[2] = {0, 0}, defm[2] = {0, 0};
#pragma unroll
        for (int c = 0; c < NW; ++c) {
            const int idx = lane + 32 * c;
            const int r = idx, cn = KP + idx;
            const bool sur = valid[c] && av[c] >= bv[c];
            surm[c] = __ballot_sync(0xffffffffu, sur);
            defm[c] = __ballot_sync(0xffffffffu, valid[c] && !sur);
#pragma unroll
            for (int w = 0; w < NS; ++w) { sw->sub[r][w] = 0u; sw->sub[cn][w] = 0u; }
            sw->parent[r] = 255; sw->parent[cn] = 255;
            sw->flow[r] = 0.0; sw->flow[cn] = 0.0;
            sw->pi[r] = 0.0; sw->pi[cn] = 0.0;
            if (valid[c]) {
                const unsigned bit = 1u << lane;
                if (sur) {
                    sw->parent[cn] = (unsigned char)r;  sw->flow[cn] = bv[c];
                    sw->amt[idx] = av[c] - bv[c];
                    sw->sub[cn][NW + c] = bit;
                    sw->sub[r][c] = bit; sw->sub[r][NW + c] = bit;
                } else {
                    sw->parent[r] = (unsigned char)cn;  sw->flow[r] = av[c];
                    sw->amt[idx] = bv[c] - av[c];
                    sw->sub[r][c] = bit;
                    sw->sub[cn][NW + c] = bit; sw->sub[cn][c] = bit;
                }
            }
        }
        __syncwarp();

        // ---------------- start basis: serial north-west corner (lane 0) ----------------
        int root = 0, chain = 0;
        if (lane == 0) {
            int n_ord = 0;
            const bool has_s = any_left(surm, NW), has_d = any_left(defm, NW);
            if (has_s && has_d) {
                int si = pop_lowest(surm, NW), dj = pop_lowest(defm, NW);
                root = si;
                double rs = sw->amt[si], rd = sw->amt[dj];
                bool placed_col = false;
                for (;;) {
                    const double f = fmin(rs, rd);
                    const double m = sM[si * KP + dj];
                    if (!placed_col) {
                        const int x = KP + dj;
                        sw->parent[x] = (unsigned char)si; sw->flow[x] = f;
                        sw->pi[x] = sw->pi[si] + m;
                        sw->ord[n_ord++] = (unsigned char)x;
                        placed_col = true;
                    } else {
                        const int x = si, p = KP + dj;
                        sw->parent[x] = (unsigned char)p; sw->flow[x] = f;
                        sw->pi[x] = sw->pi[p] - m;
                        sw->ord[n_ord++] = (unsigned char)x;
                    }
                    const bool last_s = !any_left(surm, NW), last_d = !any_left(defm, NW);
                    if (last_s && last_d) break;
                    if ((rs <= rd && !last_s) || last_d) {
                        rd = fmax(rd - rs, 0.0);
                        si = pop_lowest(surm, NW);
                        rs = sw->amt[si];
                    } else {
                        rs = fmax(rs - rd, 0.0);
                        dj = pop_lowest(defm, NW);
                        rd = sw->amt[dj];
                        placed_col = false;
                    }
                }
            } else if (has_s) {
                // a == b everywhere: chain the (row, leaf col) pairs with zero-flow arcs row_t -> col_{t-1}
                chain = 1;
                int prev = pop_lowest(surm, NW);
                root = prev;
                sw->pi[KP + prev] = sw->pi[prev] + sM[prev * KP + prev];
                sw->ord[n_ord++] = (unsigned char)(KP + prev);
                while (any_left(surm, NW)) {
                    const int cur = pop_lowest(surm, NW);
                    sw->parent[cur] = (unsigned char)(KP + prev); sw->flow[cur] = 0.0;
                    sw->pi[cur] = sw->pi[KP + prev] - sM[cur * KP + prev];
                    sw->ord[n_ord++] = (unsigned char)cur;
                    sw->pi[KP + cur] = sw->pi[cur] + sM[cur * KP + cur];
                    sw->ord[n_ord++] = (unsigned char)(KP + cur);
                    prev = cur;
                }
            } else {
                // every a_i < b_i (only through rounding): chain (col, leaf row) pairs, arcs row_{t-1} -> col_t
                chain = 1;
                int prev = pop_lowest(defm, NW);
                root = KP + prev;
                sw->pi[prev] = sw->pi[KP + prev] - sM[prev * KP + prev];
                sw->ord[n_ord++] = (unsigned char)prev;
                while (any_left(defm, NW)) {
                    const int cur = pop_lowest(defm, NW);
                    sw->parent[KP + cur] = (unsigned char)prev; sw->flow[KP + cur] = 0.0;
                    sw->pi[KP + cur] = sw->pi[prev] + sM[prev * KP + cur];
                    sw->ord[n_ord++] = (unsigned char)(KP + cur);
                    sw->pi[cur] = sw->pi[KP + cur] - sM[cur * KP + cur];
                    sw->ord[n_ord++] = (unsigned char)cur;
                    prev = cur;
                }
            }
            // subtree masks: children were placed after their parents
            for (int t = n_ord - 1; t >= 0; --t) {
                const int x = sw->ord[t], p = sw->parent[x];
#pragma unroll
                for (int w = 0; w < NS; ++w) sw->sub[p][w] |= sw->sub[x][w];
            }
        }
        root = __shfl_sync(0xffffffffu, root, 0);
        chain = __shfl_sync(0xffffffffu, chain, 0);
        __syncwarp();
        if (!chain) {
            // leaf potentials (parents are final now)
#pragma unroll
            for (int c = 0; c < NW; ++c) {
                const int idx = lane + 32 * c;
                if (valid[c]) {
                    const int r = idx, cn = KP + idx;
                    const double m = sM[idx * KP + idx];
                    if (sw->parent[cn] == r && av[c] >= bv[c]) sw->pi[cn] = sw->pi[r] + m;
                    else if (sw->parent[r] == cn && !(av[c] >= bv[c])) sw->pi[r] = sw->pi[cn] - m;
                }
            }
            __syncwarp();
        }

        // ---------------- simplex iterations ----------------
        int r0 = 0, npiv = 0, st = PILOT_ST_CONVERGED;
        for (;;) {
            double pj[NW];
#pragma unroll
            for (int c = 0; c < NW; ++c) pj[c] = sw->pi[KP + lane + 32 * c];

            // ---- block-search pricing ----
            int ei = -1, ej = -1;
            double erc = 0.0;
            for (int scanned = 0; scanned < K;) {
                const int rows = min(EMD_BLOCK_ROWS, K - scanned);
                double brc = 0.0;
                int bi = 0, bc = 0;
                for (int rr = 0; rr < rows; ++rr) {
                    int i = r0 + rr;
                    if (i >= K) i -= K;
                    const double pr = sw->pi[i];
#pragma unroll
                    for (int c = 0; c < NW; ++c) {
                        const double rc = (sM[i * KP + lane + 32 * c] + pr) - pj[c];
                        if (valid[c] && rc < brc) { brc = rc; bi = i; bc = c; }
                    }
                }
                r0 += rows; if (r0 >= K) r0 -= K;
                scanned += rows;
                // warp arg-min (most negative): order by the bits of -rc
                const unsigned long long kb = brc < 0.0 ? (unsigned long long)__double_as_longlong(-brc) : 0ULL;
                const unsigned khi = (unsigned)(kb >> 32), klo = (unsigned)kb;
                const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
                const unsigned mlo = __reduce_max_sync(0xffffffffu, khi == mhi ? klo : 0u);
                if ((mhi | mlo) == 0u) continue;
                const unsigned win = __ballot_sync(0xffffffffu, khi == mhi && klo == mlo);
                const int wl = __ffs(win) - 1;
                const int ci = __shfl_sync(0xffffffffu, bi, wl);
                const int cj = __shfl_sync(0xffffffffu, bc, wl) * 32 + wl;
                const double rc = -__longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
                const double sc = fmax(fmax(fabs(sw->pi[ci]), fabs(sw->pi[KP + cj])), fabs(sM[ci * KP + cj]));
                if (rc < -EPS * sc) { ei = ci; ej = cj; erc = rc; break; }
            }
            if (ei < 0) break;  // optimal
            if (++npiv > max_pivots) { st = PILOT_ST_MAXITER; break; }

            // ---- classify the owned nodes against the cycle of (ei -> KP+ej) ----
            const int jn = KP + ej;
            const int wi = ei >> 5, wj = jn >> 5;
            const unsigned bitI = 1u << (ei & 31), bitJ = 1u << (jn & 31);
            unsigned sb[NS][NS];
            int pc[NS];
            bool isI[NS], isJ[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int y = s * 32 + lane;
                unsigned hi = 0, hj = 0;
                int p = 0;
#pragma unroll
                for (int w = 0; w < NS; ++w) {
                    sb[s][w] = sw->sub[y][w];
                    p += __popc(sb[s][w]);
                    if (w == wi) hi = sb[s][w] & bitI;
                    if (w == wj) hj = sb[s][w] & bitJ;
                }
                pc[s] = p;
                isI[s] = hi && !hj;
                isJ[s] = hj && !hi;
            }

            // ---- ratio test: min flow over decreasing arcs, last-in-cycle-order on ties ----
            unsigned long long bestf = ~0ULL;
            unsigned bestord = 0;
            int besty = -1;
            double fl[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int y = s * 32 + lane;
                fl[s] = (isI[s] || isJ[s]) ? sw->flow[y] : 0.0;
                const bool dec = (s < NW) ? isI[s] : isJ[s];
                if (dec) {
                    const unsigned long long fb = (unsigned long long)__double_as_longlong(fl[s]);
                    const unsigned od = (s < NW) ? (unsigned)(256 - pc[s]) : (unsigned)(256 + pc[s]);
                    if (fb < bestf || (fb == bestf && od > bestord)) { bestf = fb; bestord = od; besty = y; }
                }
            }
            const unsigned fhi = (unsigned)(bestf >> 32), flo = (unsigned)bestf;
            const unsigned m1 = __reduce_min_sync(0xffffffffu, fhi);
            if (besty < 0 && m1 == 0xffffffffu) {
                // no lane has a candidate (m1 comes from ~0): unbounded -- cannot happen for a balanced problem
            }
            const unsigned m2 = __reduce_min_sync(0xffffffffu, fhi == m1 ? flo : 0xffffffffu);
            const bool tie = besty >= 0 && fhi == m1 && flo == m2;
            const unsigned m3 = __reduce_max_sync(0xffffffffu, tie ? bestord : 0u);
            if (m3 == 0u) { st = PILOT_ST_UNBOUNDED; break; }
            const unsigned lw = __ballot_sync(0xffffffffu, tie && bestord == m3);
            const int u_out = __shfl_sync(0xffffffffu, besty, __ffs(lw) - 1);
            const double delta = __longlong_as_double((long long)(((unsigned long long)m1 << 32) | m2));
            const bool sideJ = m3 > 256u;
            const int u_in = sideJ ? jn : ei, v_in = sideJ ? ei : jn;

            unsigned T2[NS];
            int pc_out = 0;
#pragma unroll
            for (int w = 0; w < NS; ++w) { T2[w] = sw->sub[u_out][w]; pc_out += __popc(T2[w]); }
            const int wo = u_out >> 5, wv = v_in >> 5;
            const unsigned bitO = 1u << (u_out & 31), bitV = 1u << (v_in & 31);
            const double sigma = sideJ ? erc : -erc;

            // ---- phase A: push delta round the cycle; stem children announce themselves ----
            bool stem[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int y = s * 32 + lane;
                const bool cyc = isI[s] || isJ[s];
                const bool dec = (s < NW) ? isI[s] : isJ[s];
                if (cyc) {
                    fl[s] = dec ? fl[s] - delta : fl[s] + delta;
                    sw->flow[y] = fl[s];
                }
                stem[s] = (sideJ ? isJ[s] : isI[s]) && pc[s] <= pc_out;
                if (stem[s] && y != u_out) sw->tmpc[sw->parent[y]] = (unsigned char)y;
            }
            __syncwarp();
            // ---- phase B: stem nodes fetch their stem child's (updated) flow and old subtree ----
            int ch[NS];
            double cf[NS];
            unsigned csub[NS][NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int y = s * 32 + lane;
                ch[s] = -1; cf[s] = 0.0;
                if (stem[s] && y != u_in) {
                    ch[s] = sw->tmpc[y];
                    cf[s] = sw->flow[ch[s]];
#pragma unroll
                    for (int w = 0; w < NS; ++w) csub[s][w] = sw->sub[ch[s]][w];
                }
            }
            __syncwarp();
            // ---- phase C: rewrite parent / flow / sub / pi ----
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int y = s * 32 + lane;
                const bool inT2 = (T2[s] >> lane) & 1u;
                if (stem[s]) {
                    if (y == u_in) {
                        sw->parent[y] = (unsigned char)v_in;
                        sw->flow[y] = delta;
#pragma unroll
                        for (int w = 0; w < NS; ++w) sw->sub[y][w] = T2[w];
                    } else {
                        sw->parent[y] = (unsigned char)ch[s];
                        sw->flow[y] = cf[s];
#pragma unroll
                        for (int w = 0; w < NS; ++w) sw->sub[y][w] = T2[w] & ~csub[s][w];
                    }
                } else if (!inT2) {
                    unsigned ho = 0, hv = 0;
#pragma unroll
                    for (int w = 0; w < NS; ++w) {
                        if (w == wo) ho = sb[s][w] & bitO;
                        if (w == wv) hv = sb[s][w] & bitV;
                    }
                    if ((ho != 0) != (hv != 0)) {
#pragma unroll
                        for (int w = 0; w < NS; ++w)
                            sw->sub[y][w] = ho ? (sb[s][w] & ~T2[w]) : (sb[s][w] | T2[w]);
                    }
                }
                if (inT2) sw->pi[y] += sigma;
            }
            __syncwarp();
        }

        // ---------------- objective: sum of flow * cost over the basic arcs ----------------
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int y = s * 32 + lane;
            const int p = sw->parent[y];
            if (p != 255 && y != root) {
                const double m = (s < NW) ? sM[y * KP + (p - KP)] : sM[p * KP + (y - KP)];
                acc += sw->flow[y] * m;
            }
        }
        acc = warp_sum_d(acc);
        if (lane == 0) {
            out[l] = acc;
            if (status) status[l] = st;
            if (pivots_out) pivots_out[l] = npiv;
        }
        __syncwarp();
    }
}

template <int NW>
static int emd_launch(const double *props, int K, const double *cost, const PairMap &pm, long long max_pivots,
                      double *out, int *status, int *pivots, unsigned long long *counter, cudaStream_t st)
{
    using SM = EmdSmem<NW>;
    const size_t smem = sizeof(double) * SM::KP * SM::KP + sizeof(SM) * EMD_WARPS;
    PILOT_CUDA(cudaFuncSetAttribute(emd_pairs_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    PILOT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, emd_pairs_kernel<NW>, EMD_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long long ctas = (long long)sm_count() * per_sm;
    const long long need = (pm.n_local + EMD_WARPS - 1) / EMD_WARPS;
    if (ctas > need) ctas = need;
    if (ctas < 1) ctas = 1;
    emd_pairs_kernel<NW><<<(unsigned)ctas, EMD_WARPS * 32, smem, st>>>(props, K, cost, pm, max_pivots, out, status,
                                                                     pivots, counter);
    PILOT_LAUNCH_CHECK();
    return 0;
}

}  // namespace pilot

extern "C" int pilot_emd_pairs(const double *props, int S, int K, const double *cost, int64_t max_pivots,
                               const pilot_pair_range *range, double *out, int32_t *status, int32_t *pivots,
                               void *workspace, size_t workspace_bytes, void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(props && cost && out && workspace, "pilot_emd_pairs: NULL pointer");
    PILOT_CHECK_ARG(S >= 1, "pilot_emd_pairs: S=%d", S);
    PILOT_CHECK_ARG(K >= 1 && K <= 64, "pilot_emd_pairs: K=%d outside the supported range [1, 64]", K);
    PILOT_CHECK_ARG(workspace_bytes >= 256, "pilot_emd_pairs: workspace too small");
    PairMap pm;
    int rc = make_pair_map(range, S, &pm);
    if (rc) return rc;
    if (pm.n_local == 0) return 0;
    if (max_pivots <= 0) max_pivots = 100000;  // POT numItermax default
    cudaStream_t st = (cudaStream_t)stream;
    PILOT_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
    if (K <= 32) return emd_launch<1>(props, K, cost, pm, max_pivots, out, status, pivots, (unsigned long long *)workspace, st);
    return emd_launch<2>(props, K, cost, pm, max_pivots, out, status, pivots, (unsigned long long *)workspace, st);
}
