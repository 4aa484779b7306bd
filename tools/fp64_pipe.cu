// tools/fp64_pipe.cu -- development aid: measures the achievable DFMA issue rate of one B200 as a function of
// independent chains per thread (ILP) and resident warps per SM sub-partition.
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
template <int ILP, bool CONSTOP>
__global__ void k(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = CONSTOP ? __fma_rn(acc[i], 0.999999, 1e-10) : __fma_rn(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, bool C>
void run(int warps_per_smsp, double* d) {
    const int threads = 128 * warps_per_smsp > 1024 ? 1024 : 128 * warps_per_smsp;   // one CTA per SM when possible
    const int ctas_per_sm = (128 * warps_per_smsp + threads - 1) / threads;
    const int blocks = 148 * ctas_per_sm;
    const int iters = 20000 / ILP;
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        CHECK(cudaEventRecord(e0));
        k<ILP, C><<<blocks, threads>>>(d, iters, 0.999999, 1e-10);
        CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * iters * 16.0 * ILP;
    printf("ILP %2d warps/SMSP %2d const %d: %8.3f ms  %6.2f T DFMA/s  = %5.1f DFMA/clk/SM @1.965GHz\n", ILP, warps_per_smsp, (int)C, best,
           ops / best / 1e9, ops / (best * 1e-3) / 148 / 1.965e9);
}
int main() {
    double* d; CHECK(cudaMalloc(&d, 148 * 16 * 1024 * sizeof(double)));
    run<1, false>(1, d); run<2, false>(1, d); run<4, false>(1, d); run<8, false>(1, d);
    run<1, false>(2, d); run<2, false>(2, d); run<4, false>(2, d);
    run<1, false>(3, d); run<2, false>(3, d); run<4, false>(3, d);
    run<1, false>(4, d); run<2, false>(4, d); run<4, false>(4, d);
    run<1, false>(8, d); run<2, false>(8, d); run<1, false>(16, d);
    run<4, true>(3, d); run<4, true>(8, d);
    return 0;
}
