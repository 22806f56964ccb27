/*
 * oracle/sinkhorn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, FP64) of the regularised-OT arithmetic reached from
 *   /root/reference/pilotpy/tools/Trajectory.py:515
 *       ot.sinkhorn2(a, b, cost, reg, method="sinkhorn_stabilized")
 * The arithmetic lives in POT (pot>=0.9.1,<0.10.0, setup.py:19; not vendored in
 * /root/reference): ot/bregman/_sinkhorn.py::sinkhorn2 -> sinkhorn_stabilized
 * with its defaults numItermax=1000, tau=1e3, stopThr=1e-9, print_period=20.
 * The algorithm restated here is SURVEY.md Appendix A.2, statement by statement,
 * in the *reference form* (the Gibbs kernel is rebuilt from alpha/beta, the
 * error is taken from the log-form plan).  A NumPy twin lives in
 * oracle/pilot_oracle.py (sinkhorn_stabilized_np); the two are cross-checked
 * in tests/test_oracle_ot.py.  Parity against POT itself is UNPINNED here; wherever POT is
 * importable tests/test_oracle_vs_pot.py checks both against ot.sinkhorn2.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int iters;        /* number of loop bodies executed */
    int absorptions;  /* number of alpha/beta absorptions */
    int status;       /* 0 converged, 1 hit numItermax, 2 NaN rollback */
    double err;       /* last evaluated marginal error */
} pilot_oracle_sk_info;

static void build_K(int n, int m, const double *M, const double *al,
                    const double *be, double reg, double *Kmat)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j)
            Kmat[(size_t)i * m + j] = exp(-(M[(size_t)i * m + j] - al[i] - be[j]) / reg);
}

/* returns sum(M * Gamma) */
double pilot_oracle_sinkhorn2(int n, int m, const double *a, const double *b,
                              const double *M, double reg, int num_iter_max,
                              double tau, double stop_thr, int check_every,
                              pilot_oracle_sk_info *info)
{
    double *Kmat = malloc(sizeof(double) * (size_t)n * m);
    double *al = calloc(n, sizeof(double)), *be = calloc(m, sizeof(double));
    double *u = malloc(sizeof(double) * n), *v = malloc(sizeof(double) * m);
    double *up = malloc(sizeof(double) * n), *vp = malloc(sizeof(double) * m);
    double *col = malloc(sizeof(double) * m);
    double err = 1.0;
    int ii, absorbed = 0, status = 1;
    for (int i = 0; i < n; ++i) u[i] = 1.0 / n;
    for (int j = 0; j < m; ++j) v[j] = 1.0 / m;
    build_K(n, m, M, al, be, reg, Kmat);
    for (ii = 0; ii < num_iter_max; ++ii) {
        memcpy(up, u, sizeof(double) * n);
        memcpy(vp, v, sizeof(double) * m);
        /* v = b / (K^T u) */
        for (int j = 0; j < m; ++j) col[j] = 0.0;
        for (int i = 0; i < n; ++i) {
            const double ui = u[i];
            const double *row = Kmat + (size_t)i * m;
            for (int j = 0; j < m; ++j) col[j] += row[j] * ui;
        }
        for (int j = 0; j < m; ++j) v[j] = b[j] / col[j];
        /* u = a / (K v) */
        double mu = 0.0, mv = 0.0;
        for (int i = 0; i < n; ++i) {
            const double *row = Kmat + (size_t)i * m;
            double s = 0.0;
            for (int j = 0; j < m; ++j) s += row[j] * v[j];
            u[i] = a[i] / s;
            if (fabs(u[i]) > mu) mu = fabs(u[i]);
        }
        for (int j = 0; j < m; ++j) if (fabs(v[j]) > mv) mv = fabs(v[j]);
        if (mu > tau || mv > tau) {
            for (int i = 0; i < n; ++i) { al[i] += reg * log(u[i]); u[i] = 1.0 / n; }
            for (int j = 0; j < m; ++j) { be[j] += reg * log(v[j]); v[j] = 1.0 / m; }
            build_K(n, m, M, al, be, reg, Kmat);
            ++absorbed;
        }
        if (ii % check_every == 0) {
            for (int j = 0; j < m; ++j) col[j] = 0.0;
            for (int i = 0; i < n; ++i) {
                const double lu = log(u[i]);
                for (int j = 0; j < m; ++j)
                    col[j] += exp(-(M[(size_t)i * m + j] - al[i] - be[j]) / reg + lu + log(v[j]));
            }
            double s2 = 0.0;
            for (int j = 0; j < m; ++j) { double d = col[j] - b[j]; s2 += d * d; }
            err = sqrt(s2);
        }
        if (err <= stop_thr) { status = 0; ++ii; break; }
        {
            int bad = 0;
            for (int i = 0; i < n; ++i) if (isnan(u[i])) bad = 1;
            for (int j = 0; j < m; ++j) if (isnan(v[j])) bad = 1;
            if (bad) {
                memcpy(u, up, sizeof(double) * n);
                memcpy(v, vp, sizeof(double) * m);
                status = 2; ++ii;
                break;
            }
        }
    }
    double cost = 0.0;
    for (int i = 0; i < n; ++i) {
        const double lu = log(u[i]);
        for (int j = 0; j < m; ++j) {
            double mij = M[(size_t)i * m + j];
            cost += mij * exp(-(mij - al[i] - be[j]) / reg + lu + log(v[j]));
        }
    }
    if (info) { info->iters = ii; info->absorptions = absorbed; info->status = status; info->err = err; }
    free(Kmat); free(al); free(be); free(u); free(v); free(up); free(vp); free(col);
    return cost;
}

/* rows [row0,row1) x all S columns of the ordered pair matrix (Trajectory.py:513-515) */
void pilot_oracle_sinkhorn_rows(int S, int K, const double *P, const double *M,
                                double reg, int row0, int row1, double *out,
                                int *iters, int *absorptions)
{
    pilot_oracle_sk_info inf;
    for (int i = row0; i < row1; ++i)
        for (int j = 0; j < S; ++j) {
            size_t o = (size_t)(i - row0) * S + j;
            out[o] = pilot_oracle_sinkhorn2(K, K, P + (size_t)i * K, P + (size_t)j * K, M,
                                            reg, 1000, 1e3, 1e-9, 20, &inf);
            if (iters) iters[o] = inf.iters;
            if (absorptions) absorptions[o] = inf.absorptions;
        }
}
