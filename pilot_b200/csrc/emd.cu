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
//   * compaction:     a cycle has ~7 of the 128 nodes: the cycle nodes are ballot-compacted so
//                     that ONE lane handles ONE cycle node (ratio test, flow push, re-hanging);
//                     only cycles longer than 32 nodes take the 4-nodes-per-lane general path
//   * start basis:    diagonal arcs (i,i) with min(a_i,b_i), then a greedy north-west corner on
//                     the residuals: the staircase always continues with the CHEAPEST remaining
//                     column (row) of the current row (column) -> no artificial arcs,
//                     |pi| = O(max cost), ~4x fewer pivots than the artificial-root start
// Any exact solver returns the same optimum; parity with the
// oracle is <1e-12 relative.
#include <stdlib.h>
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
    unsigned char tmpc[N];   // stem-child scatter during pivots; placement order (`ord`) during the start basis
    unsigned char size[N];   // popcount of sub[]
    unsigned char clist[64]; // compacted cycle nodes (bit 7: j-side)
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
                 int *__restrict__ pivots_out, unsigned long long *__restrict__ counter, int fast_limit, int block_rows)
{
    using SM = EmdSmem<NW>;
    constexpr int KP = SM::KP, NS = SM::NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LDM = KP + 1;  // odd row stride: rows AND columns of the cost matrix are conflict free
    double *sM = reinterpret_cast<double *>(smem_raw);  // KP x LDM, row-major, zero padded
    SM *sw = reinterpret_cast<SM *>(smem_raw + sizeof(double) * KP * LDM) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int t = threadIdx.x; t < KP * KP; t += blockDim.x) {
        const int i = t / KP, j = t - i * KP;
        sM[i * LDM + j] = (i < K && j < K) ? cost[i * K + j] : 0.0;
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

        // ---------------- start basis: greedy north-west corner (whole warp, uniform state) ----------------
        int root = 0, chain = 0, n_ord = 0;
        {
            const bool has_s = any_left(surm, NW), has_d = any_left(defm, NW);
            if (has_s && has_d) {
                // order-preserving 64-bit image of a double (costs may be negative in general)
                auto okey = [](double v) -> unsigned long long {
                    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
                    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
                };
                // arg-min over the still available rows/columns; lane l looks at l and l + 32
                auto pick = [&](const unsigned (&avail)[2], bool want_col, int fixed) -> int {
                    unsigned long long best = ~0ULL;
                    int bidx = -1;
#pragma unroll
                    for (int c = 0; c < NW; ++c) {
                        const int idx = lane + 32 * c;
                        if ((avail[c] >> lane) & 1u) {
                            const double m = want_col ? sM[fixed * LDM + idx] : sM[idx * LDM + fixed];
                            const unsigned long long kk = okey(m);
                            if (kk < best) { best = kk; bidx = idx; }
                        }
                    }
                    const unsigned bh = (unsigned)(best >> 32), bl = (unsigned)best;
                    const unsigned mh = __reduce_min_sync(0xffffffffu, bh);
                    const unsigned ml = __reduce_min_sync(0xffffffffu, bh == mh ? bl : 0xffffffffu);
                    const unsigned win = __ballot_sync(0xffffffffu, bidx >= 0 && bh == mh && bl == ml);
                    return __shfl_sync(0xffffffffu, bidx, __ffs(win) - 1);
                };
                int si = pop_lowest(surm, NW);
                root = si;
                int dj = pick(defm, true, si);
                defm[dj >> 5] &= ~(1u << (dj & 31));
                double rs = sw->amt[si], rd = sw->amt[dj];
                bool placed_col = false;
                for (;;) {
                    const double f = fmin(rs, rd);
                    const double m = sM[si * LDM + dj];
                    if (lane == 0) {
                        if (!placed_col) {
                            const int x = KP + dj;
                            sw->parent[x] = (unsigned char)si; sw->flow[x] = f;
                            sw->pi[x] = sw->pi[si] + m;
                            sw->tmpc[n_ord] = (unsigned char)x;
                        } else {
                            const int x = si, p = KP + dj;
                            sw->parent[x] = (unsigned char)p; sw->flow[x] = f;
                            sw->pi[x] = sw->pi[p] - m;
                            sw->tmpc[n_ord] = (unsigned char)x;
                        }
                    }
                    ++n_ord;
                    placed_col = true;
                    __syncwarp();
                    const bool last_s = !any_left(surm, NW), last_d = !any_left(defm, NW);
                    if (last_s && last_d) break;
                    if ((rs <= rd && !last_s) || last_d) {
                        rd = fmax(rd - rs, 0.0);
                        si = pick(surm, false, dj);        // cheapest remaining row for the current column
                        surm[si >> 5] &= ~(1u << (si & 31));
                        rs = sw->amt[si];
                    } else {
                        rs = fmax(rs - rd, 0.0);
                        dj = pick(defm, true, si);         // cheapest remaining column for the current row
                        defm[dj >> 5] &= ~(1u << (dj & 31));
                        rd = sw->amt[dj];
                        placed_col = false;
                    }
                }
            } else if (lane == 0) {
                if (has_s) {
                    // a == b everywhere: chain the (row, leaf col) pairs with zero-flow arcs row_t -> col_{t-1}
                    int prev = pop_lowest(surm, NW);
                    root = prev;
                    sw->pi[KP + prev] = sw->pi[prev] + sM[prev * LDM + prev];
                    sw->tmpc[n_ord++] = (unsigned char)(KP + prev);
                    while (any_left(surm, NW)) {
                        const int cur = pop_lowest(surm, NW);
                        sw->parent[cur] = (unsigned char)(KP + prev); sw->flow[cur] = 0.0;
                        sw->pi[cur] = sw->pi[KP + prev] - sM[cur * LDM + prev];
                        sw->tmpc[n_ord++] = (unsigned char)cur;
                        sw->pi[KP + cur] = sw->pi[cur] + sM[cur * LDM + cur];
                        sw->tmpc[n_ord++] = (unsigned char)(KP + cur);
                        prev = cur;
                    }
                } else {
                    // every a_i < b_i (only through rounding): chain (col, leaf row) pairs, arcs row_{t-1} -> col_t
                    int prev = pop_lowest(defm, NW);
                    root = KP + prev;
                    sw->pi[prev] = sw->pi[KP + prev] - sM[prev * LDM + prev];
                    sw->tmpc[n_ord++] = (unsigned char)prev;
                    while (any_left(defm, NW)) {
                        const int cur = pop_lowest(defm, NW);
                        sw->parent[KP + cur] = (unsigned char)prev; sw->flow[KP + cur] = 0.0;
                        sw->pi[KP + cur] = sw->pi[prev] + sM[prev * LDM + cur];
                        sw->tmpc[n_ord++] = (unsigned char)(KP + cur);
                        sw->pi[cur] = sw->pi[KP + cur] - sM[cur * LDM + cur];
                        sw->tmpc[n_ord++] = (unsigned char)cur;
                        prev = cur;
                    }
                }
            }
            if (!(has_s && has_d)) {
                chain = 1;
                root = __shfl_sync(0xffffffffu, root, 0);
                n_ord = __shfl_sync(0xffffffffu, n_ord, 0);
            }
            __syncwarp();
            // subtree masks: children were placed after their parents
            if (lane == 0)
                for (int t = n_ord - 1; t >= 0; --t) {
                    const int x = sw->tmpc[t], p = sw->parent[x];
#pragma unroll
                    for (int w = 0; w < NS; ++w) sw->sub[p][w] |= sw->sub[x][w];
                }
        }
        __syncwarp();
        if (!chain) {
            // leaf potentials (parents are final now)
#pragma unroll
            for (int c = 0; c < NW; ++c) {
                const int idx = lane + 32 * c;
                if (valid[c]) {
                    const int r = idx, cn = KP + idx;
                    const double m = sM[idx * LDM + idx];
                    if (sw->parent[cn] == r && av[c] >= bv[c]) sw->pi[cn] = sw->pi[r] + m;
                    else if (sw->parent[r] == cn && !(av[c] >= bv[c])) sw->pi[r] = sw->pi[cn] - m;
                }
            }
            __syncwarp();
        }

        // subtree sizes (popcount of the masks), kept next to the masks from here on
#pragma unroll
        for (int sl_ = 0; sl_ < NS; ++sl_) {
            const int y = sl_ * 32 + lane;
            int p = 0;
#pragma unroll
            for (int w = 0; w < NS; ++w) p += __popc(sw->sub[y][w]);
            sw->size[y] = (unsigned char)p;
        }
        __syncwarp();

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
                const int rows = min(block_rows, K - scanned);
                double brc = 0.0;
                int bi = 0, bc = 0;
                for (int rr = 0; rr < rows; ++rr) {
                    int i = r0 + rr;
                    if (i >= K) i -= K;
                    const double pr = sw->pi[i];
#pragma unroll
                    for (int c = 0; c < NW; ++c) {
                        const double rc = (sM[i * LDM + lane + 32 * c] + pr) - pj[c];
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
                const double sc = fmax(fmax(fabs(sw->pi[ci]), fabs(sw->pi[KP + cj])), fabs(sM[ci * LDM + cj]));
                if (rc < -EPS * sc) { ei = ci; ej = cj; erc = rc; break; }
            }
            if (ei < 0) break;  // optimal
            if (++npiv > max_pivots) { st = PILOT_ST_MAXITER; break; }

            // ---- classify the owned nodes against the cycle of (ei -> KP+ej) ----
            const int jn = KP + ej;
            const int wi = ei >> 5, wj = jn >> 5;
            const unsigned bitI = 1u << (ei & 31), bitJ = 1u << (jn & 31);
            // compact the cycle nodes: i-side = has i but not j, j-side = has j but not i
            int ncyc = 0;
#pragma unroll
            for (int sl_ = 0; sl_ < NS; ++sl_) {
                const int y = sl_ * 32 + lane;
                const unsigned hi_ = sw->sub[y][wi] & bitI, hj_ = sw->sub[y][wj] & bitJ;
                const bool cI = hi_ && !hj_, cJ = hj_ && !hi_;
                const unsigned balI = __ballot_sync(0xffffffffu, cI), balJ = __ballot_sync(0xffffffffu, cJ);
                const int nI = __popc(balI);
                if (cI) { const int pos = ncyc + __popc(balI & lt_mask); if (pos < 64) sw->clist[pos] = (unsigned char)y; }
                if (cJ) { const int pos = ncyc + nI + __popc(balJ & lt_mask); if (pos < 64) sw->clist[pos] = (unsigned char)(y | 0x80); }
                ncyc += nI + __popc(balJ);
            }
            __syncwarp();
            if (ncyc <= fast_limit) {
                // ---------- one lane per cycle node ----------
                const bool has = lane < ncyc;
                const int ye = has ? sw->clist[lane] : 0;
                const int y = ye & 0x7f;
                const bool onJ = (ye & 0x80) != 0;
                const double f = has ? sw->flow[y] : 0.0;
                const int sz = has ? sw->size[y] : 0;
                const int par = has ? sw->parent[y] : 0;
                const bool dec = has && (onJ ? (y >= KP) : (y < KP));  // i-side rows / j-side columns lose flow
                // ratio test: min flow, ties -> last in cycle order (Cunningham)
                const unsigned long long fb = dec ? (unsigned long long)__double_as_longlong(f) : ~0ULL;
                const unsigned od = dec ? (onJ ? (unsigned)(256 + sz) : (unsigned)(256 - sz)) : 0u;
                const unsigned fh = (unsigned)(fb >> 32), fl_ = (unsigned)fb;
                const unsigned m1 = __reduce_min_sync(0xffffffffu, fh);
                const unsigned m2 = __reduce_min_sync(0xffffffffu, fh == m1 ? fl_ : 0xffffffffu);
                const bool tie = dec && fh == m1 && fl_ == m2;
                const unsigned m3 = __reduce_max_sync(0xffffffffu, tie ? od : 0u);
                if (m3 == 0u) { st = PILOT_ST_UNBOUNDED; break; }
                const int wl = __ffs(__ballot_sync(0xffffffffu, tie && od == m3)) - 1;
                const int u_out = __shfl_sync(0xffffffffu, y, wl);
                const int pc_out = __shfl_sync(0xffffffffu, sz, wl);
                const double delta = __longlong_as_double((long long)(((unsigned long long)m1 << 32) | m2));
                const bool sideJ = m3 > 256u;
                const int u_in = sideJ ? jn : ei, v_in = sideJ ? ei : jn;
                const double sigma = sideJ ? erc : -erc;
                unsigned T2[NS];
#pragma unroll
                for (int w = 0; w < NS; ++w) T2[w] = sw->sub[u_out][w];
                // phase A: push delta round the cycle; stem children announce themselves
                const double fU = dec ? f - delta : f + delta;
                const bool same = has && (onJ == sideJ);
                const bool stem = same && sz <= pc_out;
                if (has) sw->flow[y] = fU;
                if (stem && y != u_out) sw->tmpc[par] = (unsigned char)y;
                __syncwarp();
                // phase B: stem nodes fetch their stem child's (updated) flow, old subtree and size
                int ch = 0, csz = 0;
                double cf = 0.0;
                unsigned csub[NS];
#pragma unroll
                for (int w = 0; w < NS; ++w) csub[w] = 0u;
                if (stem && y != u_in) {
                    ch = sw->tmpc[y];
                    cf = sw->flow[ch];
                    csz = sw->size[ch];
#pragma unroll
                    for (int w = 0; w < NS; ++w) csub[w] = sw->sub[ch][w];
                }
                __syncwarp();
                // phase C: rewrite parent / flow / sub / size of the cycle nodes
                if (stem) {
                    if (y == u_in) {
                        sw->parent[y] = (unsigned char)v_in;
                        sw->flow[y] = delta;
                        sw->size[y] = (unsigned char)pc_out;
#pragma unroll
                        for (int w = 0; w < NS; ++w) sw->sub[y][w] = T2[w];
                    } else {
                        sw->parent[y] = (unsigned char)ch;
                        sw->flow[y] = cf;
                        sw->size[y] = (unsigned char)(pc_out - csz);
#pragma unroll
                        for (int w = 0; w < NS; ++w) sw->sub[y][w] = T2[w] & ~csub[w];
                    }
                } else if (same) {  // above u_out on its side: loses the cut subtree
                    sw->size[y] = (unsigned char)(sz - pc_out);
#pragma unroll
                    for (int w = 0; w < NS; ++w) sw->sub[y][w] &= ~T2[w];
                } else if (has) {   // the other side: v_in and its ancestors below the join gain it
                    sw->size[y] = (unsigned char)(sz + pc_out);
#pragma unroll
                    for (int w = 0; w < NS; ++w) sw->sub[y][w] |= T2[w];
                }
                // potentials of the re-hung subtree
#pragma unroll
                for (int sl_ = 0; sl_ < NS; ++sl_)
                    if ((T2[sl_] >> lane) & 1u) sw->pi[sl_ * 32 + lane] += sigma;
                __syncwarp();
                continue;
            }
            // ---------- general path (cycles longer than 32 nodes): 4 nodes per lane ----------
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
            // keep the stored subtree sizes in step with the masks
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int y = s * 32 + lane;
                int p = 0;
#pragma unroll
                for (int w = 0; w < NS; ++w) p += __popc(sw->sub[y][w]);
                sw->size[y] = (unsigned char)p;
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
                const double m = (s < NW) ? sM[y * LDM + (p - KP)] : sM[p * LDM + (y - KP)];
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
    const size_t smem = sizeof(double) * SM::KP * (SM::KP + 1) + sizeof(SM) * EMD_WARPS;
    PILOT_CUDA(cudaFuncSetAttribute(emd_pairs_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    PILOT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, emd_pairs_kernel<NW>, EMD_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long long ctas = (long long)sm_count() * per_sm;
    const long long need = (pm.n_local + EMD_WARPS - 1) / EMD_WARPS;
    if (ctas > need) ctas = need;
    if (ctas < 1) ctas = 1;
    // PILOT_EMD_FAST_LIMIT (0..32, tests only): longest cycle handled by the one-lane-per-node path
    int fast_limit = 32;
    if (const char *e = getenv("PILOT_EMD_FAST_LIMIT")) {
        fast_limit = atoi(e);
        if (fast_limit < 0) fast_limit = 0;
        if (fast_limit > 32) fast_limit = 32;
    }
    // pricing block: measured on 400-500 K problems, rows 2/4/6/8/16/64 -> K = 64: 37.6/32.6/35.8/37.4/34.5/57.1 ms
    // (pivots 90 ... 49: a full Dantzig pass halves the pivots but prices 16x more arcs per pivot); K = 30, 40: 8
    // rows are 3-5 % faster than 4
    const int block_rows = K <= 48 ? 8 : EMD_BLOCK_ROWS;
    emd_pairs_kernel<NW><<<(unsigned)ctas, EMD_WARPS * 32, smem, st>>>(props, K, cost, pm, max_pivots, out, status,
                                                                     pivots, counter, fast_limit, block_rows);
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
