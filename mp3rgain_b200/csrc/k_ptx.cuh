// k_ptx.cuh -- the few inline-PTX primitives the decode kernels share (PTX ISA 8.x, sm_90+; compiled for sm_100a only):
// mbarrier transaction barriers and 1-D bulk copies (TMA unit, `cp.async.bulk`: UBLKCP in the SASS), and named barriers.
// tests/dev/simt_emu/build_emu.py replaces these bodies by name with the interpreter's rendezvous primitives.
#pragma once
#include <cstdint>

__device__ __forceinline__ uint32_t k2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k2_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k2_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void k2_mbar_expect(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k2_bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(k2_smem_u32(dst)), "l"(src), "r"(bytes), "r"(k2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void k2_mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "K2_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra K2_DONE;\n"
        "bra K2_WAIT;\n"
        "K2_DONE:\n"
        "}" ::"r"(k2_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void k2_pair_sync(int pair) {   // immediate ids, so that ptxas reserves 5 barriers and not 16
    switch (pair) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}
// prmt.b32 with the selector's replicate-sign bits honoured (bit 3 of a nibble: the result byte is the selected byte's sign,
// 0x00 or 0xFF).  The CUDA intrinsic __byte_perm masks those bits away.
__device__ __forceinline__ uint32_t k2_prmt(uint32_t a, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0u), "r"(sel));
    return r;
}
// cp.async: 16 bytes global -> shared without a register round trip (LDGSTS), tracked in per-thread groups
__device__ __forceinline__ void k2_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(k2_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void k2_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void k2_cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void k2_cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// Packed f32x2 arithmetic (sm_100+: FFMA2 / FADD2 / FMUL2 in the SASS): two IEEE round-to-nearest operations per instruction, each
// half rounded exactly like the scalar __fmaf_rn / __fadd_rn / __fmul_rn (no flush-to-zero, no contraction across halves).  The
// kernels that use them are bound by instruction issue, not by the FMA pipe: the pipe does the same work, in half the issue
// slots.  A half built as make_float2(w, w) costs nothing: ptxas folds it into the instruction's scalar-broadcast operand.
__device__ __forceinline__ float2 k2_ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 k2_fadd2(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 k2_fsub2(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// The packed product is written as fma(a, b, -0.0): x + (-0.0) == x for every x, signed zeros included.  A plain mul.rn.f32x2 is
// not safe: ptxas 12.9 contracts it with a following add.rn.f32x2 / sub.rn.f32x2 into one FFMA2 (it honours .rn for the scalar
// forms only; --fmad=false does not stop it, and a literal -0.0 addend is folded back into a multiplication first), which rounds
// once instead of twice.  The -0.0 therefore comes from constant memory, where ptxas cannot see its value; it costs one register
// per thread and no instruction (FFMA2 takes it as a scalar-broadcast operand).  tools/f32x2_check.cu checks the instructions
// themselves against the scalar forms on the device.
__constant__ float K2_NEGZERO = -0.0f;
__device__ __forceinline__ float2 k2_fmul2(float2 a, float2 b) {
    float2 d;
    const float nz = K2_NEGZERO;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %6}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(nz));
    return d;
}
