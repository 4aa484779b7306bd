// Development aid: can an SM sub-partition issue integer / FP32 work in the cycles a saturated FP64 pipe leaves free?
// One kernel, 8 warps per CTA, one CTA per SM: `role_mask` bit w says warp w runs the FP64 loop (dependent-free DFMA chains),
// the other warps run an integer (IMAD / LOP3) or FP32 (FFMA) loop.  Times: FP64 warps alone, the others alone, both together.
// If together == max(alone) the pipes co-issue; if together == sum, a DFMA blocks the dispatch port for its two cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/coissue tools/coissue.cu && tools/bin/coissue
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) k(unsigned fp64_mask, unsigned other_mask, int other_kind, int iters, double* outd, unsigned* outu) {
    const int warp = threadIdx.x >> 5;
    if (fp64_mask >> warp & 1) {
        double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
        const double m = 1.0000001, c = 1e-9;
        for (int i = 0; i < iters; i++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
        outd[blockIdx.x * 256 + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    } else if (other_mask >> warp & 1) {
        if (other_kind == 0) {
            unsigned a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
            for (int i = 0; i < iters; i++) {
                a0 = a0 * 2654435761u + 12345u; a1 = (a1 ^ a0) + (a1 >> 3); a2 = a2 * 40503u + a1; a3 = (a3 & a2) | (a3 << 1);
                a4 = a4 * 2654435761u + 777u; a5 = (a5 ^ a4) + (a5 >> 5); a6 = a6 * 40503u + a5; a7 = (a7 & a6) | (a7 << 2);
            }
            outu[blockIdx.x * 256 + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
        } else {
            float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
            const float m = 1.0001f, c = 1e-5f;
            for (int i = 0; i < iters; i++) {
                a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
                a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
            }
            outu[blockIdx.x * 256 + threadIdx.x] = __float_as_uint(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7);
        }
    }
}

static float run(unsigned fm, unsigned om, int kind, int iters, double* d, unsigned* u) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 256>>>(fm, om, kind, iters, d, u);
    cudaEventRecord(e0);
    k<<<148, 256>>>(fm, om, kind, iters, d, u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    double* d; unsigned* u;
    cudaMalloc(&d, 148 * 256 * 8); cudaMalloc(&u, 148 * 256 * 4);
    const int iters = 200000;
    // warps 0-3 FP64 (one per sub-partition), warps 4-7 the other role
    for (int kind = 0; kind < 2; kind++) {
        const float a = run(0x0F, 0x00, kind, iters, d, u), b = run(0x00, 0xF0, kind, iters, d, u), both = run(0x0F, 0xF0, kind, iters, d, u);
        const float a2 = run(0xFF, 0x00, kind, iters, d, u);
        printf("%s: FP64 on 4 warps %.2f ms (on 8 warps %.2f), %s on 4 warps %.2f ms, both %.2f ms  (max %.2f, sum %.2f)\n", kind ? "FP32 FFMA" : "INT IMAD/LOP3",
               a, a2, kind ? "FFMA" : "INT", b, both, a > b ? a : b, a + b);
    }
    printf("coissue done (%s)\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
