/*
 * oracle/emd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the exact-EMD arithmetic reached from the reference at
 *   /root/reference/pilotpy/tools/Trajectory.py:511   ot.emd2(a, b, cost)
 * The arithmetic itself lives in a third-party dependency that is NOT vendored
 * under /root/reference: POT (pot>=0.9.1,<0.10.0, setup.py:19), files
 *   ot/lp/__init__.py::emd2 -> ot/lp/emd_wrap.pyx::emd_c -> ot/lp/EMD_wrap.cpp
 *   -> ot/lp/network_simplex_simple.h (LEMON-derived primal network simplex).
 * This file restates the published algorithm:
 *   - complete bipartite graph n x m, arc cost M[i][j], supplies a_i, demands b_j
 *   - zero-mass rows/columns are dropped before the solve
 *   - one artificial root, big-M artificial arcs, all real arcs start non-basic
 *   - "mixed" arc storage order, block-search pivot rule with block size
 *     max(floor(sqrt(n*m)), 10) and the relative-epsilon entering test
 *   - strongly-feasible leaving rule (strict '<' on the first path, '<=' on the second)
 *   - objective accumulated as sum(flow * M) over real arcs
 * Parity is UNPINNED against POT itself (POT cannot be installed here); the
 * optimum is cross-checked against SciPy HiGHS in tests/test_oracle_ot.py and, wherever
 * POT is importable, against ot.emd2 itself in tests/test_oracle_vs_pot.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Work counters (pivots, arcs priced, potential updates, cycle steps) are
 * exported because SURVEY.md 8(d) defines the EMD kernel's algorithmic work as
 * 3*A + 2*U + 2*C of THIS solver on the same input.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    long long pivots;
    long long arcs_priced;
    long long pot_updates;
    long long cycle_steps;
} pilot_oracle_emd_stats;

enum { EMD_INFEASIBLE = 0, EMD_OPTIMAL = 1, EMD_UNBOUNDED = 2, EMD_MAX_ITER = 3 };

#define ST_TREE 0
#define ST_LOWER 1

typedef struct {
    int N;          /* real nodes (ns + nt) */
    int root;
    long narcs;     /* real arcs */
    int *src, *tgt; /* per arc (real + artificial) */
    double *cost, *flow;
    signed char *state;
    /* per node */
    int *parent, *pred, *first_child, *next_sib, *prev_sib, *depth;
    signed char *up; /* 1: pred arc runs node -> parent */
    double *pi;
} ns_t;

static void unlink_child(ns_t *g, int u)
{
    int p = g->parent[u];
    if (g->prev_sib[u] >= 0) g->next_sib[g->prev_sib[u]] = g->next_sib[u];
    else g->first_child[p] = g->next_sib[u];
    if (g->next_sib[u] >= 0) g->prev_sib[g->next_sib[u]] = g->prev_sib[u];
    g->prev_sib[u] = g->next_sib[u] = -1;
}

static void link_child(ns_t *g, int u, int p)
{
    g->parent[u] = p;
    g->prev_sib[u] = -1;
    g->next_sib[u] = g->first_child[p];
    if (g->first_child[p] >= 0) g->prev_sib[g->first_child[p]] = u;
    g->first_child[p] = u;
}

/* Solve min <G,M>, G1=a, G^T1=b on the already zero-filtered problem. */
static int ns_solve(int ns, int nt, const double *a, const double *b,
                    const double *M, int ldm, const int *rid, const int *cid,
                    long long max_iter, double *G, int ldg, double *cost_out,
                    double *alpha, double *beta, pilot_oracle_emd_stats *st)
{
    ns_t g;
    const int N = ns + nt;
    const long narcs = (long)ns * nt;
    const long tot = narcs + N;
    int rc = EMD_OPTIMAL;
    g.N = N; g.root = N; g.narcs = narcs;
    g.src = malloc(sizeof(int) * tot);
    g.tgt = malloc(sizeof(int) * tot);
    g.cost = malloc(sizeof(double) * tot);
    g.flow = calloc(tot, sizeof(double));
    g.state = malloc(tot);
    g.parent = malloc(sizeof(int) * (N + 1));
    g.pred = malloc(sizeof(int) * (N + 1));
    g.first_child = malloc(sizeof(int) * (N + 1));
    g.next_sib = malloc(sizeof(int) * (N + 1));
    g.prev_sib = malloc(sizeof(int) * (N + 1));
    g.depth = malloc(sizeof(int) * (N + 1));
    g.up = malloc(N + 1);
    g.pi = malloc(sizeof(double) * (N + 1));

    /* mixed arc order (A.1 step 4) */
    long blk = (long)floor(sqrt((double)narcs));
    if (blk < 10) blk = 10;
    {
        long p = 0, wrap = 0;
        double maxc = 0.0;
        for (int i = 0; i < ns; ++i)
            for (int j = 0; j < nt; ++j) {
                g.src[p] = i;
                g.tgt[p] = ns + j;
                g.cost[p] = M[(long)rid[i] * ldm + cid[j]];
                g.state[p] = ST_LOWER;
                if (fabs(g.cost[p]) > maxc) maxc = fabs(g.cost[p]);
                p += blk;
                if (p >= narcs) p = ++wrap;
            }
        double art = (maxc + 1.0) * (double)(N + 1);
        /* artificial star around the root */
        for (int u = 0; u <= N; ++u) {
            g.first_child[u] = g.next_sib[u] = g.prev_sib[u] = -1;
        }
        g.parent[N] = -1; g.pred[N] = -1; g.depth[N] = 0; g.pi[N] = 0.0; g.up[N] = 0;
        double ssum = 0.0;
        for (int u = N - 1; u >= 0; --u) {
            long e = narcs + u;
            double sup = (u < ns) ? a[u] : -b[u - ns];
            ssum += sup;
            g.state[e] = ST_TREE;
            g.pred[u] = (int)e;
            g.depth[u] = 1;
            link_child(&g, u, N);
            if (sup >= 0) {
                g.up[u] = 1; g.src[e] = u; g.tgt[e] = N;
                g.flow[e] = sup; g.cost[e] = 0.0; g.pi[u] = 0.0;
            } else {
                g.up[u] = 0; g.src[e] = N; g.tgt[e] = u;
                g.flow[e] = -sup; g.cost[e] = art; g.pi[u] = art;
            }
        }
        if (fabs(ssum) > 1e-8) { rc = EMD_INFEASIBLE; goto done; }
    }

    {
        const double EPS = 2.2204460492503131e-15; /* 10 * DBL_EPSILON, as POT */
        long next_arc = 0;
        long long iter = 0;
        for (;;) {
            /* ---- block-search pricing ---- */
            long in_arc = -1;
            double minrc = 0.0;
            {
                long cnt = blk, e = next_arc, scanned = 0;
                int found = 0;
                while (scanned < narcs) {
                    double c = g.state[e] * (g.cost[e] + g.pi[g.src[e]] - g.pi[g.tgt[e]]);
                    st->arcs_priced++;
                    if (c < minrc) { minrc = c; in_arc = e; }
                    ++scanned;
                    if (++e == narcs) e = 0;
                    if (--cnt == 0) {
                        if (in_arc >= 0) {
                            double pa = fabs(g.pi[g.src[in_arc]]), pb = fabs(g.pi[g.tgt[in_arc]]);
                            double sc = pa > pb ? pa : pb;
                            if (fabs(g.cost[in_arc]) > sc) sc = fabs(g.cost[in_arc]);
                            if (minrc < -EPS * sc) { found = 1; break; }
                        }
                        cnt = blk;
                    }
                }
                if (!found) {
                    if (in_arc < 0) break;
                    double pa = fabs(g.pi[g.src[in_arc]]), pb = fabs(g.pi[g.tgt[in_arc]]);
                    double sc = pa > pb ? pa : pb;
                    if (fabs(g.cost[in_arc]) > sc) sc = fabs(g.cost[in_arc]);
                    if (!(minrc < -EPS * sc)) break; /* optimal */
                }
                next_arc = e;
            }
            if (++iter > max_iter) { rc = EMD_MAX_ITER; break; }
            st->pivots++;

            /* ---- join node ---- */
            int first = g.src[in_arc], second = g.tgt[in_arc];
            int u = first, v = second;
            while (u != v) {
                if (g.depth[u] >= g.depth[v]) u = g.parent[u];
                else v = g.parent[v];
            }
            int join = u;

            /* ---- leaving arc (strongly feasible rule) ---- */
            double delta = INFINITY;
            int u_out = -1, side = 0;
            for (u = first; u != join; u = g.parent[u]) {
                st->cycle_steps++;
                if (g.up[u]) {
                    double d = g.flow[g.pred[u]];
                    if (d < delta) { delta = d; u_out = u; side = 1; }
                }
            }
            for (u = second; u != join; u = g.parent[u]) {
                st->cycle_steps++;
                if (!g.up[u]) {
                    double d = g.flow[g.pred[u]];
                    if (d <= delta) { delta = d; u_out = u; side = 2; }
                }
            }
            if (u_out < 0) { rc = EMD_UNBOUNDED; break; }
            int u_in = side == 1 ? first : second;
            int v_in = side == 1 ? second : first;

            /* ---- push delta round the cycle ---- */
            if (delta > 0) {
                g.flow[in_arc] += delta;
                for (u = first; u != join; u = g.parent[u])
                    g.flow[g.pred[u]] += g.up[u] ? -delta : delta;
                for (u = second; u != join; u = g.parent[u])
                    g.flow[g.pred[u]] += g.up[u] ? delta : -delta;
            }
            g.state[in_arc] = ST_TREE;
            g.state[g.pred[u_out]] = ST_LOWER;

            /* ---- re-hang the cut-off subtree: reverse the stem u_in..u_out ---- */
            unlink_child(&g, u_out);
            {
                int child = u_in, new_parent = v_in;
                int carry_pred = (int)in_arc;
                signed char carry_up = (u_in == g.src[in_arc]);
                while (1) {
                    int old_parent = g.parent[child];
                    int old_pred = g.pred[child];
                    signed char old_up = g.up[child];
                    int last = (child == u_out);
                    if (!last) unlink_child(&g, child);
                    link_child(&g, child, new_parent);
                    g.pred[child] = carry_pred;
                    g.up[child] = carry_up;
                    if (last) break;
                    carry_pred = old_pred;
                    carry_up = !old_up;
                    new_parent = child;
                    child = old_parent;
                }
            }

            /* ---- shift potentials and depths of the re-hung subtree ---- */
            {
                double sigma = (u_in == g.src[in_arc]) ? -minrc : minrc;
                int x = u_in;
                for (;;) {
                    g.pi[x] += sigma;
                    g.depth[x] = g.depth[g.parent[x]] + 1;
                    st->pot_updates++;
                    if (g.first_child[x] >= 0) { x = g.first_child[x]; continue; }
                    while (x != u_in && g.next_sib[x] < 0) x = g.parent[x];
                    if (x == u_in) break;
                    x = g.next_sib[x];
                }
            }
        }
    }

    /* artificial arcs still carrying flow => infeasible (cannot happen for balanced input) */
    if (rc == EMD_OPTIMAL)
        for (long e = narcs; e < tot; ++e)
            if (g.flow[e] > 1e-8) { rc = EMD_INFEASIBLE; break; }

done:
    {
        double c = 0.0;
        if (rc != EMD_INFEASIBLE) {
            /* accumulate in natural (i,j) order over the mixed storage */
            long p = 0, wrap = 0;
            for (int i = 0; i < ns; ++i)
                for (int j = 0; j < nt; ++j) {
                    double f = g.flow[p];
                    if (f != 0.0) {
                        c += f * g.cost[p];
                        if (G) G[(long)rid[i] * ldg + cid[j]] = f;
                    }
                    p += blk;
                    if (p >= narcs) p = ++wrap;
                }
            if (alpha) for (int i = 0; i < ns; ++i) alpha[rid[i]] = -g.pi[i];
            if (beta) for (int j = 0; j < nt; ++j) beta[cid[j]] = g.pi[ns + j];
        }
        *cost_out = c;
    }
    free(g.src); free(g.tgt); free(g.cost); free(g.flow); free(g.state);
    free(g.parent); free(g.pred); free(g.first_child); free(g.next_sib);
    free(g.prev_sib); free(g.depth); free(g.up); free(g.pi);
    return rc;
}

/*
 * EMD_wrap-level entry: a (n), b (m) already made equal-mass by the caller
 * (ot.emd2 does b = b * a.sum() / b.sum() in NumPy before calling emd_c).
 * G (n*m, may be NULL) is zero-filled then receives the optimal plan.
 * Returns the POT result code.
 */
int pilot_oracle_emd(int n, int m, const double *a, const double *b,
                     const double *M, long long max_iter, double *cost,
                     double *G, double *alpha, double *beta,
                     pilot_oracle_emd_stats *stats)
{
    pilot_oracle_emd_stats local = {0, 0, 0, 0};
    pilot_oracle_emd_stats *st = stats ? stats : &local;
    int *rid = malloc(sizeof(int) * (n > 0 ? n : 1));
    int *cid = malloc(sizeof(int) * (m > 0 ? m : 1));
    double *aa = malloc(sizeof(double) * (n > 0 ? n : 1));
    double *bb = malloc(sizeof(double) * (m > 0 ? m : 1));
    int ns = 0, nt = 0, rc;
    memset(st, 0, sizeof(*st));
    if (G) memset(G, 0, sizeof(double) * (size_t)n * m);
    if (alpha) memset(alpha, 0, sizeof(double) * n);
    if (beta) memset(beta, 0, sizeof(double) * m);
    *cost = 0.0;
    for (int i = 0; i < n; ++i) {
        if (a[i] > 0) { rid[ns] = i; aa[ns++] = a[i]; }
        else if (a[i] < 0) { rc = EMD_INFEASIBLE; goto out; }
    }
    for (int j = 0; j < m; ++j) {
        if (b[j] > 0) { cid[nt] = j; bb[nt++] = b[j]; }
        else if (b[j] < 0) { rc = EMD_INFEASIBLE; goto out; }
    }
    if (ns == 0 || nt == 0) { rc = EMD_OPTIMAL; goto out; }
    rc = ns_solve(ns, nt, aa, bb, M, m, rid, cid, max_iter, G, m, cost, alpha, beta, st);
out:
    free(rid); free(cid); free(aa); free(bb);
    return rc;
}

/*
 * Batched driver for the CPU baseline: all ordered pairs (i, j) of the rows of
 * P (S x K), i in [row0, row1), j in [0, S); same per-pair arithmetic as the
 * Python-level loop at Trajectory.py:508-511 (including emd2's rescale of b).
 * out is (row1-row0) x S.  Returns the number of non-OPTIMAL solves.
 */
int pilot_oracle_emd_rows(int S, int K, const double *P, const double *M,
                          int row0, int row1, double *out,
                          pilot_oracle_emd_stats *stats)
{
    int bad = 0;
    double *bs = malloc(sizeof(double) * K);
    pilot_oracle_emd_stats acc = {0, 0, 0, 0}, one;
    for (int i = row0; i < row1; ++i) {
        const double *a = P + (size_t)i * K;
        double sa = 0.0;
        for (int k = 0; k < K; ++k) sa += a[k];
        for (int j = 0; j < S; ++j) {
            const double *b = P + (size_t)j * K;
            double sb = 0.0, c;
            for (int k = 0; k < K; ++k) sb += b[k];
            for (int k = 0; k < K; ++k) bs[k] = b[k] * sa / sb;
            if (pilot_oracle_emd(K, K, a, bs, M, 100000, &c, NULL, NULL, NULL, &one) != EMD_OPTIMAL)
                ++bad;
            acc.pivots += one.pivots; acc.arcs_priced += one.arcs_priced;
            acc.pot_updates += one.pot_updates; acc.cycle_steps += one.cycle_steps;
            out[(size_t)(i - row0) * S + j] = c;
        }
    }
    if (stats) *stats = acc;
    free(bs);
    return bad;
}
