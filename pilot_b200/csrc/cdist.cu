// Kernel (2b): K x K distances between cell-type centroids, FP64.  Replaces
// squareform(pdist(centroids, metric)) (reference pilotpy/tools/Trajectory.py:468-469;
// SciPy _distance_wrap / _distance_pybind arithmetic) and the cost / cost.max()
// normalisation fed to the OT solvers (Trajectory.py:101).
// K <= a few hundred, D <= a few hundred: negligible work, one thread per (i<j) pair,
// sequential non-contracted accumulation so results track SciPy to ~1 ulp.
#include "common.cuh"

namespace pilot {

__global__ void cdist_prep_kernel(const double *__restrict__ C, int K, int D, int metric,
                                  double *__restrict__ norms, double *__restrict__ means)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const double *x = C + (long long)i * D;
    double mean = 0.0;
    if (metric == PILOT_METRIC_CORRELATION) {
        // numpy mean(axis=1): pairwise sum; sequential here (<= 1 ulp apart)
        double s = 0.0;
        for (int d = 0; d < D; ++d) s = __dadd_rn(s, x[d]);
        mean = __ddiv_rn(s, (double)D);
    }
    double ss = 0.0;
    for (int d = 0; d < D; ++d) {
        const double v = __dsub_rn(x[d], mean);
        ss = __dadd_rn(ss, __dmul_rn(v, v));
    }
    norms[i] = sqrt(ss);
    means[i] = mean;
}

__global__ void cdist_pair_kernel(const double *__restrict__ C, int K, int D, int metric,
                                  const double *__restrict__ norms, const double *__restrict__ means,
                                  double *__restrict__ cost)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)K * K) return;
    const int i = (int)(t / K), j = (int)(t - (long long)i * K);
    if (i == j) { cost[t] = 0.0; return; }
    if (i > j) return;
    const double *u = C + (long long)i * D, *v = C + (long long)j * D;
    double acc = 0.0, r;
    switch (metric) {
    case PILOT_METRIC_BRAYCURTIS: {
        double den = 0.0;
        for (int d = 0; d < D; ++d) {
            acc = __dadd_rn(acc, fabs(__dsub_rn(u[d], v[d])));
            den = __dadd_rn(den, fabs(__dadd_rn(u[d], v[d])));
        }
        r = __ddiv_rn(acc, den);
        break;
    }
    case PILOT_METRIC_CANBERRA:
        for (int d = 0; d < D; ++d) {
            const double num = fabs(__dsub_rn(u[d], v[d])), den = __dadd_rn(fabs(u[d]), fabs(v[d]));
            if (den != 0.0) acc = __dadd_rn(acc, __ddiv_rn(num, den));  // SciPy: a 0 / 0 term contributes 0
        }
        r = acc;
        break;
    case PILOT_METRIC_SEUCLIDEAN:
        // V = var(centroids, axis=0, ddof=1), SciPy's default for pdist(..., 'seuclidean')
        for (int d = 0; d < D; ++d) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s = __dadd_rn(s, C[(long long)k * D + d]);
            const double mean = __ddiv_rn(s, (double)K);
            double ss = 0.0;
            for (int k = 0; k < K; ++k) {
                const double e = __dsub_rn(C[(long long)k * D + d], mean);
                ss = __dadd_rn(ss, __dmul_rn(e, e));
            }
            const double var = __ddiv_rn(ss, (double)(K - 1));
            const double df = __dsub_rn(u[d], v[d]);
            acc = __dadd_rn(acc, __ddiv_rn(__dmul_rn(df, df), var));
        }
        r = sqrt(acc);
        break;
    case PILOT_METRIC_HAMMING:
        for (int d = 0; d < D; ++d) acc += u[d] != v[d] ? 1.0 : 0.0;
        r = __ddiv_rn(acc, (double)D);
        break;
    case PILOT_METRIC_COSINE:
    case PILOT_METRIC_CORRELATION: {
        const double mu = means[i], mv = means[j];
        for (int d = 0; d < D; ++d)
            acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(u[d], mu), __dsub_rn(v[d], mv)));
        double c = __ddiv_rn(acc, __dmul_rn(norms[i], norms[j]));
        if (fabs(c) > 1.0) c = copysign(1.0, c);  // SciPy clips rounding overshoot
        r = __dsub_rn(1.0, c);
        break;
    }
    case PILOT_METRIC_EUCLIDEAN:
    case PILOT_METRIC_MINKOWSKI:  // SciPy's default p = 2 takes its Euclidean routine (bit-identical results)
    case PILOT_METRIC_SQEUCLIDEAN:
        for (int d = 0; d < D; ++d) {
            const double df = __dsub_rn(u[d], v[d]);
            acc = __dadd_rn(acc, __dmul_rn(df, df));
        }
        r = metric == PILOT_METRIC_SQEUCLIDEAN ? acc : sqrt(acc);
        break;
    case PILOT_METRIC_CITYBLOCK:
        for (int d = 0; d < D; ++d) acc = __dadd_rn(acc, fabs(__dsub_rn(u[d], v[d])));
        r = acc;
        break;
    default:  // PILOT_METRIC_CHEBYSHEV
        for (int d = 0; d < D; ++d) acc = fmax(acc, fabs(__dsub_rn(u[d], v[d])));
        r = acc;
        break;
    }
    cost[(long long)i * K + j] = r;
    cost[(long long)j * K + i] = r;
}

// one CTA: max over the matrix, then cost_norm = cost / max (IEEE division, as NumPy)
__global__ void cdist_norm_kernel(const double *__restrict__ cost, int K, double *__restrict__ cost_norm,
                                  double *__restrict__ cost_max)
{
    __shared__ double s_red[32];
    const int n = K * K;
    double m = -INFINITY;
    bool has_nan = false;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const double v = cost[t];
        if (v != v) has_nan = true;
        m = fmax(m, v);
    }
    m = warp_max_d(m);
    has_nan = __any_sync(0xffffffffu, has_nan);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = has_nan ? NAN : m;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : -INFINITY;
        bool nn = v != v;
        nn = __any_sync(0xffffffffu, nn);
        v = warp_max_d(nn ? -INFINITY : v);
        if (threadIdx.x == 0) s_red[0] = nn ? NAN : v;  // ndarray.max() propagates NaN
    }
    __syncthreads();
    const double mx = s_red[0];
    if (threadIdx.x == 0 && cost_max) *cost_max = mx;
    if (cost_norm)
        for (int t = threadIdx.x; t < n; t += blockDim.x) cost_norm[t] = __ddiv_rn(cost[t], mx);
}

}  // namespace pilot

extern "C" int pilot_cdist(const double *centroids_f64, int K, int D, int metric, double *cost,
                           double *cost_norm, double *cost_max, void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(K >= 1 && D >= 1 && centroids_f64 && cost, "pilot_cdist: bad argument");
    PILOT_CHECK_ARG(metric >= PILOT_METRIC_COSINE && metric <= PILOT_METRIC_HAMMING,
                    "pilot_cdist: unsupported metric id %d", metric);
    PILOT_CHECK_ARG(K <= 4096, "pilot_cdist: K=%d too large", K);
    PILOT_CHECK_ARG(cost_norm != nullptr, "pilot_cdist: cost_norm is required (it doubles as scratch)");
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 1) {
        // a single centroid: cost = [[0]]; cost / cost.max() = 0/0 = NaN like NumPy
        PILOT_CUDA(cudaMemsetAsync(cost, 0, sizeof(double), st));
        cdist_norm_kernel<<<1, 32, 0, st>>>(cost, 1, cost_norm, cost_max);
        PILOT_LAUNCH_CHECK();
        return 0;
    }
    double *scratch = cost_norm;  // norms[K], means[K]: K*K >= 2K for K >= 2
    cdist_prep_kernel<<<(K + 127) / 128, 128, 0, st>>>(centroids_f64, K, D, metric, scratch, scratch + K);
    PILOT_LAUNCH_CHECK();
    const long long n = (long long)K * K;
    cdist_pair_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(centroids_f64, K, D, metric, scratch,
                                                                  scratch + K, cost);
    PILOT_LAUNCH_CHECK();
    cdist_norm_kernel<<<1, 1024, 0, st>>>(cost, K, cost_norm, cost_max);
    PILOT_LAUNCH_CHECK();
    return 0;
}
