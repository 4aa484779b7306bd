// Development aid: do the packed f32x2 instructions (FFMA2 / FADD2 / FMUL2, with the swapped / negated / broadcast operand forms
// the kernels rely on) round exactly like the scalar __fmaf_rn / __fadd_rn / __fmul_rn, subnormals and signed zeros included?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/f32x2_check tools/f32x2_check.cu && tools/bin/f32x2_check
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "../mp3rgain_b200/csrc/k_ptx.cuh"

__device__ unsigned int d_bad[10];

__device__ __forceinline__ bool same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }

__global__ void check(const float* __restrict__ v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 6 > n) return;
    const float a = v[i], b = v[i + 1], c = v[i + 2], d = v[i + 3], e = v[i + 4], f = v[i + 5];
    float2 r;
    r = k2_ffma2(make_float2(a, b), make_float2(c, d), make_float2(e, f));
    if (!same(r.x, __fmaf_rn(a, c, e)) || !same(r.y, __fmaf_rn(b, d, f))) atomicAdd(&d_bad[0], 1u);
    r = k2_fadd2(make_float2(a, b), make_float2(c, d));
    if (!same(r.x, __fadd_rn(a, c)) || !same(r.y, __fadd_rn(b, d))) atomicAdd(&d_bad[1], 1u);
    r = k2_fsub2(make_float2(a, b), make_float2(c, d));
    if (!same(r.x, __fsub_rn(a, c)) || !same(r.y, __fsub_rn(b, d))) atomicAdd(&d_bad[2], 1u);
    r = k2_fmul2(make_float2(a, b), make_float2(c, d));
    if (!same(r.x, __fmul_rn(a, c)) || !same(r.y, __fmul_rn(b, d))) atomicAdd(&d_bad[3], 1u);
    // broadcast operand, swapped + half-negated operand (the complex-arithmetic forms)
    r = k2_ffma2(make_float2(b, -a), make_float2(c, c), k2_fmul2(make_float2(a, b), make_float2(d, d)));
    if (!same(r.x, __fmaf_rn(b, c, __fmul_rn(a, d))) || !same(r.y, __fmaf_rn(-a, c, __fmul_rn(b, d)))) atomicAdd(&d_bad[4], 1u);
    const float2 q = make_float2(e, f), m = make_float2(a, b);
    const float2 miq = make_float2(q.y, -q.x);
    r = k2_fadd2(m, miq);
    if (!same(r.x, __fadd_rn(a, f)) || !same(r.y, __fsub_rn(b, e))) atomicAdd(&d_bad[5], 1u);
    r = k2_fsub2(m, miq);
    if (!same(r.x, __fsub_rn(a, f)) || !same(r.y, __fadd_rn(b, e))) atomicAdd(&d_bad[6], 1u);
    r = k2_fmul2(k2_fsub2(make_float2(a, b), make_float2(c, d)), make_float2(0.5002f, 0.5002f));
    if (!same(r.x, __fmul_rn(__fsub_rn(a, c), 0.5002f)) || !same(r.y, __fmul_rn(__fsub_rn(b, d), 0.5002f))) atomicAdd(&d_bad[7], 1u);
    // a product followed by a sum / difference must round twice (ptxas contracts the plain packed forms: see k_ptx.cuh)
    r = k2_fadd2(k2_fmul2(make_float2(a, b), make_float2(c, d)), make_float2(e, f));
    if (!same(r.x, __fadd_rn(__fmul_rn(a, c), e)) || !same(r.y, __fadd_rn(__fmul_rn(b, d), f))) atomicAdd(&d_bad[8], 1u);
    r = k2_fsub2(make_float2(e, f), k2_fmul2(make_float2(a, b), make_float2(0.8660254f, 0.8660254f)));
    if (!same(r.x, __fsub_rn(e, __fmul_rn(a, 0.8660254f))) || !same(r.y, __fsub_rn(f, __fmul_rn(b, 0.8660254f)))) atomicAdd(&d_bad[9], 1u);
}

int main() {
    const int n = 1 << 22;
    float* h = new float[n];
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < n; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        uint32_t bits = (uint32_t)(s >> 16);
        const uint32_t kind = (uint32_t)(s >> 60);
        if (kind < 5) bits &= 0x807FFFFFu;                                   // subnormals and zeros
        else if (kind < 8) bits = (bits & 0x807FFFFFu) | ((1u + (bits >> 23) % 6u) << 23);   // just above the subnormal range
        else if (kind < 14) bits = (bits & 0x807FFFFFu) | ((100u + (bits >> 23) % 56u) << 23);   // ordinary magnitudes
        if (kind == 4 && (s & 0x300) == 0) bits &= 0x80000000u;              // signed zeros
        memcpy(&h[i], &bits, 4);
    }
    float* d;
    cudaMalloc(&d, n * sizeof(float));
    cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice);
    check<<<n / 256, 256>>>(d, n);
    unsigned int bad[10];
    cudaMemcpyFromSymbol(bad, d_bad, sizeof(bad));
    const char* names[10] = {"fma", "add", "sub", "mul", "complex mul (swap, negated half, broadcast)", "m + (-i q)", "m - (-i q)", "(a - b) * imm", "product then sum (two roundings)", "sum minus product (two roundings)"};
    int total = 0;
    for (int k = 0; k < 10; k++) { printf("%-48s mismatches %u of %d\n", names[k], bad[k], n - 5); total += bad[k]; }
    printf("f32x2_check %s (%s)\n", total ? "FAILED" : "ok", cudaGetErrorString(cudaGetLastError()));
    return total != 0;
}
