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
