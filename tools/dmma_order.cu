// Which scalar operation order reproduces mma.sync.m8n8k4.f64 bit for bit?  (Decides whether the warp-form
// Sinkhorn tail can be made bit-identical to the DMMA panels.)   nvcc -arch=sm_100a -o dmma_order dmma_order.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// A: 8 x KK row-major, B: KK x 8 (col-major fragments: B[k][n]), C = A B accumulated over KK/4 DMMAs
__global__ void k(const double *A, const double *B, int KK, double *Cmma, double *Cseq, double *Crev, double *Cpair,
                  double *Cnofma)
{
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    for (int ks = 0; ks < KK / 4; ++ks) {
        const double a = A[g * KK + 4 * ks + t];      // A fragment: row g, col t
        const double b = B[(4 * ks + t) * 8 + g];     // B fragment: row t (k), col g (n)
        dmma884(c0, c1, a, b);
    }
    // C fragment: row g, cols 2t, 2t+1
    Cmma[g * 8 + 2 * t] = c0;
    Cmma[g * 8 + 2 * t + 1] = c1;
    // scalar emulations: thread handles entries (row, col) = (lane/4, 2*(lane%4) + {0,1})
    for (int h = 0; h < 2; ++h) {
        const int row = g, col = 2 * t + h;
        double s = 0.0, r = 0.0, p = 0.0, q = 0.0;
        for (int ks = 0; ks < KK / 4; ++ks) {
            const double *a = A + row * KK + 4 * ks;
            const double b0 = B[(4 * ks + 0) * 8 + col], b1 = B[(4 * ks + 1) * 8 + col];
            const double b2 = B[(4 * ks + 2) * 8 + col], b3 = B[(4 * ks + 3) * 8 + col];
            s = fma(a[3], b3, fma(a[2], b2, fma(a[1], b1, fma(a[0], b0, s))));          // k ascending FMA chain
            r = fma(a[0], b0, fma(a[1], b1, fma(a[2], b2, fma(a[3], b3, r))));          // k descending
            p = p + (fma(a[1], b1, a[0] * b0) + fma(a[3], b3, a[2] * b2));              // pairwise, then accumulate
            q = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q, __dmul_rn(a[0], b0)), __dmul_rn(a[1], b1)),
                                    __dmul_rn(a[2], b2)), __dmul_rn(a[3], b3));        // unfused
        }
        Cseq[row * 8 + col] = s; Crev[row * 8 + col] = r; Cpair[row * 8 + col] = p; Cnofma[row * 8 + col] = q;
    }
}

int main()
{
    const int KK = 64, trials = 2000;
    double *A, *B, *C[5];
    cudaMallocManaged(&A, 8 * KK * 8); cudaMallocManaged(&B, KK * 8 * 8);
    for (int i = 0; i < 5; ++i) cudaMallocManaged(&C[i], 64 * 8);
    long mism[4] = {0, 0, 0, 0};
    srand(1);
    for (int tr = 0; tr < trials; ++tr) {
        for (int i = 0; i < 8 * KK; ++i) A[i] = exp(-8.0 * rand() / RAND_MAX) * (tr % 2 ? 1.0 : (rand() % 2 ? 1 : -1));
        for (int i = 0; i < KK * 8; ++i) B[i] = exp(6.0 * rand() / RAND_MAX - 3.0);
        k<<<1, 32>>>(A, B, KK, C[0], C[1], C[2], C[3], C[4]);
        cudaDeviceSynchronize();
        for (int e = 0; e < 64; ++e)
            for (int v = 0; v < 4; ++v)
                if (C[0][e] != C[v + 1][e]) ++mism[v];
    }
    printf("entries %d: mismatches vs DMMA: fma-chain k-ascending %ld, k-descending %ld, pairwise %ld, unfused %ld\n",
           trials * 64, mism[0], mism[1], mism[2], mism[3]);
    return 0;
}
