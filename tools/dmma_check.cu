// Standalone check: is mma.sync.m8n8k4.f64 correct and how fast is it on this GPU (vs DFMA)?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dmma_check tools/dmma_check.cu && /tmp/dmma_check
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>

__global__ void dmma_gemm_check(const double* A, const double* B, double* C)
{
    // one warp: C[8x8] = A[8x4] * B[4x8]; A row-major (8x4), B "col" fragment: B[k][n]
    int lane = threadIdx.x;
    int g = lane >> 2, t = lane & 3;
    double a = A[g * 4 + t];       // A[row=g][k=t]
    double b = B[t * 8 + g];       // B[k=t][n=g]
    double c0 = 0, c1 = 0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    C[g * 8 + 2 * t] = c0;         // C[row=g][col=2t], [2t+1]
    C[g * 8 + 2 * t + 1] = c1;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_rate(double* out, int iters)
{
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = 1.0 + threadIdx.x * 1e-3, b = 0.5 + threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_rate(double* out, int iters)
{
    double c[NACC];
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    double hA[32], hB[32], hC[64], ref[64];
    for (int i = 0; i < 32; ++i) { hA[i] = sin(i + 1.0); hB[i] = cos(2.0 * i + 0.5); }
    for (int r = 0; r < 8; ++r) for (int n = 0; n < 8; ++n) { double s = 0; for (int k = 0; k < 4; ++k) s += hA[r * 4 + k] * hB[k * 8 + n]; ref[r * 8 + n] = s; }
    double *dA, *dB, *dC, *dout;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dC, sizeof(hC));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    dmma_gemm_check<<<1, 32>>>(dA, dB, dC);
    cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost);
    double md = 0; for (int i = 0; i < 64; ++i) md = fmax(md, fabs(hC[i] - ref[i]));
    printf("dmma correctness: max abs diff %.3e (%s)\n", md, cudaGetErrorString(cudaGetLastError()));
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int ctas = sms * 8, th = 256, iters = 8192;
    cudaMalloc(&dout, sizeof(double) * ctas * th);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](auto kern, const char* name, double flops_per_thread_iter) {
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0); kern<<<ctas, th>>>(dout, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
        }
        double tf = flops_per_thread_iter * iters * (double)ctas * th / (best * 1e-3) / 1e12;
        printf("%-22s %8.3f ms  %8.2f TFLOP/s  (%s)\n", name, best, tf, cudaGetErrorString(cudaGetLastError()));
    };
    run(dmma_rate<1>, "dmma 1 acc", 512.0 / 32);
    run(dmma_rate<2>, "dmma 2 acc", 2 * 512.0 / 32);
    run(dmma_rate<4>, "dmma 4 acc", 4 * 512.0 / 32);
    run(dmma_rate<8>, "dmma 8 acc", 8 * 512.0 / 32);
    run(dmma_rate<16>, "dmma 16 acc", 16 * 512.0 / 32);
    run(dfma_rate<8>, "dfma 8 acc", 8 * 2.0);
    run(dfma_rate<16>, "dfma 16 acc", 16 * 2.0);
    run(dfma_rate<32>, "dfma 32 acc", 32 * 2.0);
    return 0;
}
