// Pipe-peak microbenchmarks: the FP64 / FP32 FMA rates and the FP64 mma.sync (DMMA) rate of
// this GPU, used as roofline denominators for the Sinkhorn and EMD kernels (SURVEY.md 8d:
// MEASURED_PEAKS.json holds only HBM and bf16 tensor peaks).
#include "common.cuh"

namespace pilot {

constexpr int PEAK_ITERS = 4096;
constexpr int PEAK_ILP = 8;

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *out, T seed)
{
    T a[PEAK_ILP];
#pragma unroll
    for (int i = 0; i < PEAK_ILP; ++i) a[i] = seed + (T)(threadIdx.x + i);
    const T m = (T)0.999999, c = (T)1e-7;
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_ILP; ++i) a[i] = a[i] * m + c;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < PEAK_ILP; ++i) s += a[i];
    if (s == (T)12345.678) out[0] = s;  // never true; keeps the loop alive
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, double seed)
{
    double c[PEAK_ILP][2];
#pragma unroll
    for (int i = 0; i < PEAK_ILP; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const double a = seed + threadIdx.x * 1e-3, b = seed * 0.5 + threadIdx.x * 1e-3;
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PEAK_ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

}  // namespace pilot

extern "C" int pilot_pipe_peak(int kind, double *h_tflops, void *stream)
{
    using namespace pilot;
    PILOT_CHECK_ARG(kind >= 0 && kind <= 2 && h_tflops, "pilot_pipe_peak: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    double *d = nullptr;
    PILOT_CUDA(cudaMalloc((void **)&d, 64));
    cudaEvent_t e0, e1;
    PILOT_CUDA(cudaEventCreate(&e0));
    PILOT_CUDA(cudaEventCreate(&e1));
    const int ctas = sm_count() * 8, th = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        PILOT_CUDA(cudaEventRecord(e0, st));
        if (kind == 0) fma_peak_kernel<double><<<ctas, th, 0, st>>>(d, 1.0);
        else if (kind == 1) fma_peak_kernel<float><<<ctas, th, 0, st>>>((float *)d, 1.0f);
        else dmma_peak_kernel<<<ctas, th, 0, st>>>(d, 1.0);
        PILOT_CUDA(cudaEventRecord(e1, st));
        PILOT_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        PILOT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops;
        if (kind <= 1) flops = 2.0 * PEAK_ITERS * PEAK_ILP * (double)ctas * th;
        else flops = 2.0 * 8 * 8 * 4 * (double)PEAK_ILP * PEAK_ITERS * (double)ctas * (th / 32);
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *h_tflops = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    PILOT_LAUNCH_CHECK();
    return 0;
}
