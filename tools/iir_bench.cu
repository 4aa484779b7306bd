// tools/iir_bench.cu -- development aid: times alternative instantiations of the K4 kernel (k_iir_t) on synthetic PCM
// laid out exactly as the engine lays it out (planar f32, 16-window blocks with one warm-up window per chain).
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/iir_bench tools/iir_bench.cu
//   ./tools/iir_bench [tracks=200]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../mp3rgain_b200/csrc/rg_coeffs.h"
#include "../mp3rgain_b200/csrc/k_analysis.cuh"

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_fill(float* p, size_t n, uint32_t seed) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint32_t s = (uint32_t)i * 2654435761u + seed;
        s ^= s >> 15; s *= 2246822519u; s ^= s >> 13; s *= 3266489917u; s ^= s >> 16;
        p[i] = ((int)(s & 0xFFFF) - 32768) / 32768.0f * 0.25f;
    }
}

template <int NV, int THREADS, int MINB>
static double run(const char* name, const RgChain* d_chains, uint32_t n_chains, float* d_peaks, double* d_wsum, size_t n_w,
                  std::vector<double>& ref, double audio_s, const RgCoefFast& cf) {
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    const uint32_t blocks = (n_chains + THREADS - 1) / THREADS;
    float best = 1e9f;
    for (int it = 0; it < 4; it++) {
        CHECK(cudaEventRecord(a));
        k_iir_fast_t<NV, THREADS, MINB><<<blocks, THREADS>>>(d_chains, n_chains, d_peaks, 0, cf);
        CHECK(cudaEventRecord(b));
        CHECK(cudaEventSynchronize(b));
        float ms; CHECK(cudaEventElapsedTime(&ms, a, b));
        if (it) best = std::min(best, ms);
    }
    std::vector<double> h(n_w);
    CHECK(cudaMemcpy(h.data(), d_wsum, n_w * sizeof(double), cudaMemcpyDeviceToHost));
    double maxrel = 0;
    if (ref.empty()) ref = h;
    else for (size_t i = 0; i < n_w; i++) maxrel = std::max(maxrel, fabs(h[i] - ref[i]) / (fabs(ref[i]) + 1e-300));
    cudaFuncAttributes fa;
    CHECK(cudaFuncGetAttributes(&fa, k_iir_fast_t<NV, THREADS, MINB>));
    int occ = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_iir_fast_t<NV, THREADS, MINB>, THREADS, 0));
    printf("%-28s %8.3f ms  %7.1f audio-h/s  regs %3d local %4zu  warps/SM %2d  max rel diff vs first %.2e\n", name, best,
           audio_s / 3600.0 / (best * 1e-3), fa.numRegs, fa.localSizeBytes, occ * THREADS / 32, maxrel);
    return best;
}

int main(int argc, char** argv) {
    const int n_tracks = argc > 1 ? atoi(argv[1]) : 200;
    const uint32_t win = 2205, bw = argc > 2 ? (uint32_t)atoi(argv[2]) : 16;
    const size_t frames = 6891ull * 1152;
    const uint32_t n_win = (uint32_t)((frames + win - 1) / win);
    const size_t plane = frames;
    float* d_pcm; double* d_wsum; float* d_wmax; float* d_peaks; RgChain* d_chains;
    CHECK(cudaMalloc(&d_pcm, (size_t)n_tracks * 2 * plane * sizeof(float) + 64));
    const size_t n_w = (size_t)n_tracks * 2 * n_win;
    CHECK(cudaMalloc(&d_wsum, n_w * sizeof(double)));
    CHECK(cudaMalloc(&d_wmax, n_w * sizeof(float)));
    k_fill<<<148 * 8, 256>>>(d_pcm, (size_t)n_tracks * 2 * plane, 12345u);
    CHECK(cudaDeviceSynchronize());
    std::vector<RgChain> chains;
    for (int t = 0; t < n_tracks; t++)
        for (int c = 0; c < 2; c++) {
            const float* src = d_pcm + ((size_t)t * 2 + c) * plane;
            double* wb = d_wsum + ((size_t)t * 2 + c) * n_win;
            float* mb = d_wmax + ((size_t)t * 2 + c) * n_win;
            for (uint32_t w0 = 0; w0 < n_win; w0 += bw) {
                const uint32_t w1 = std::min(n_win, w0 + bw);
                const uint64_t s0 = (uint64_t)w0 * win, s1 = std::min<uint64_t>((uint64_t)w1 * win, frames);
                const uint64_t warm = std::min<uint64_t>(s0, win);
                RgChain ch{};
                ch.src = src + (s0 - warm); ch.wsum = wb + w0; ch.wmax = mb + w0;
                ch.n_warm = (uint32_t)warm; ch.n_main = (uint32_t)(s1 - s0); ch.win = win; ch.rate_idx = 4; ch.flags = s0 == 0;
                chains.push_back(ch);
            }
        }
    CHECK(cudaMalloc(&d_chains, chains.size() * sizeof(RgChain)));
    CHECK(cudaMemcpy(d_chains, chains.data(), chains.size() * sizeof(RgChain), cudaMemcpyHostToDevice));
    CHECK(cudaMalloc(&d_peaks, chains.size() * sizeof(float)));
    RgCoefFast cf;
    {
        const double* row = RG_COEFS[4];   // 44100
        for (int i = 1; i <= 10; i++) cf.a[i - 1] = row[i];
        for (int i = 0; i <= 10; i++) cf.b[i] = row[11 + i] * 32768.0;
        cf.A[0] = row[23]; cf.A[1] = row[24];
        for (int i = 0; i < 3; i++) cf.B[i] = row[25 + i];
        cf.k2 = 2.2e-7;
        for (int i = 0; i < 10; i++) cf.s0[i] = 0;
        cf.u0[0] = cf.u0[1] = cf.uz[0] = cf.uz[1] = 0;
    }
    const double audio_s = (double)n_tracks * frames / 44100.0;
    printf("%d tracks, %zu chains, %.1f audio-hours\n", n_tracks, chains.size(), audio_s / 3600);
    std::vector<double> ref;
    const uint32_t nc = (uint32_t)chains.size();
#define RUN(NV, T, B) run<NV, T, B>("nv" #NV " T" #T " minb" #B, d_chains, nc, d_peaks, d_wsum, n_w, ref, audio_s, cf)
    if (argc > 2) { printf("block_windows %u\n", bw); RUN(3, 256, 2); return 0; }
    RUN(5, 128, 3);
    RUN(2, 128, 3);
    RUN(2, 128, 4);
    RUN(3, 128, 4);
    RUN(4, 128, 4);
    RUN(2, 256, 2);
    RUN(3, 256, 2);
    RUN(2, 64, 8);
    RUN(3, 64, 8);
    RUN(2, 128, 5);
    RUN(2, 128, 6);
    RUN(1, 128, 4);
    RUN(1, 128, 5);
    RUN(1, 128, 6);
    RUN(1, 128, 8);
    RUN(1, 64, 10);
    RUN(1, 64, 12);
    RUN(3, 128, 4);
    RUN(5, 128, 3);
    RUN(1, 256, 3);
    RUN(1, 32, 20);
    RUN(1, 32, 24);
    return 0;
}
