// cuda_runtime.h (tests/dev/simt_emu) -- a CPU SIMT interpreter for the engine's CUDA sources.  TEST INFRASTRUCTURE ONLY.
//
// What it is: the product's kernel sources (mp3rgain_b200/csrc/*.cuh, rg_engine.cu) compiled UNCHANGED by g++ against this
// header (found first on the include path, so it stands in for the toolkit's <cuda_runtime.h>), with `<<<...>>>` launches
// rewritten by build_emu.py.  Every CUDA thread of a block runs as a ucontext fiber; __syncthreads, named barriers, the
// warp collectives (__shfl*_sync, __ballot_sync, __any_sync, __reduce_*_sync, __match_any_sync, __syncwarp) and the
// mbarrier / bulk-copy helpers of K2 are rendezvous points between fibers.  Blocks run one after another, streams run in
// issue order, device memory is host memory (poisoned on allocation).
//
// What it is for: running the -m gpu parity tests' logic in the GPU-less container, so that a change to a kernel's index
// arithmetic, walk logic or table handling is checked against the oracle before it costs GPU minutes (tests/test_emu_*.py).
// What it is not: a product path (nothing under mp3rgain_b200/ knows it exists; the product loads librg_engine.so only), a
// timing model, or a race detector (one legal schedule is executed; -ffp-contract=off stands in for -fmad=false, libm for
// libdevice).
#pragma once
#include <ucontext.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <time.h>
#include <type_traits>
#include <source_location>
#include <vector>

#define RG_SIMT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__
#define __constant__ static
#define warpSize 32

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(4))) short2 { short x, y; };
struct __attribute__((aligned(4))) ushort2 { unsigned short x, y; };
struct __attribute__((aligned(8))) short4 { short x, y, z, w; };
struct __attribute__((aligned(8))) ushort4 { unsigned short x, y, z, w; };
struct __attribute__((aligned(4))) uchar4 { unsigned char x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline short2 make_short2(short x, short y) { return short2{x, y}; }
static inline ushort2 make_ushort2(unsigned short x, unsigned short y) { return ushort2{x, y}; }
static inline short4 make_short4(short x, short y, short z, short w) { return short4{x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }

// ---------------------------------------------------------------------------------------------------------------------
namespace emu {

// Context switch: a hand-written callee-saved-register switch on x86-64 (swapcontext costs a sigprocmask system call per
// switch, a third of the run time of a collective-heavy kernel); ucontext elsewhere.
#if defined(__x86_64__)
#define EMU_ASM_SWITCH 1
extern "C" void emu_switch_ctx(void** save_sp, void* new_sp);
asm(".text\n.local emu_switch_ctx\n.type emu_switch_ctx,@function\nemu_switch_ctx:\n"
    "pushq %rbp\npushq %rbx\npushq %r12\npushq %r13\npushq %r14\npushq %r15\n"
    "movq %rsp, (%rdi)\nmovq %rsi, %rsp\n"
    "popq %r15\npopq %r14\npopq %r13\npopq %r12\npopq %rbx\npopq %rbp\nret\n"
    ".size emu_switch_ctx, .-emu_switch_ctx\n");
#endif

struct Fiber {
    ucontext_t ctx;
    void* sp;
    uint3 tid;
    unsigned lin;          // linear thread id
    bool done, waiting;
    // warp collectives
    bool c_wait, c_done;
    uint32_t c_mask;
    int c_op;
    uint64_t c_val, c_res;
    int c_arg;
    const char* c_file; unsigned c_line;
    unsigned stc_calls;    // __syncthreads_count / _and / _or calls of this thread (they alternate between two accumulators)
};

struct State {
    std::vector<Fiber> fib;
    std::vector<char*> stacks;
    ucontext_t sched;
    void* sched_sp = nullptr;
    Fiber* cur = nullptr;
    unsigned nthreads = 0, live = 0;
    dim3 blockDim, gridDim;
    uint3 blockIdx;
    // block barrier 0 and named barriers 1..15
    unsigned bar_arrived[16] = {}, bar_gen[16] = {};
    int stc_acc[2] = {0, 0};
    uint64_t progress = 0;
    const void* body = nullptr;
    void (*thunk)(const void*) = nullptr;
    unsigned char* dyn_smem = nullptr;
    uint64_t launches = 0, fibers_run = 0;
    const char* kernel = "";
};
inline State& S() { static State s; return s; }
static const size_t STACK = 192 * 1024;
static const size_t DYN_SMEM = 256 * 1024;

inline void die(const char* msg) {
    State& s = S();
    fprintf(stderr, "simt_emu: %s (kernel %s, block %u, thread %u)\n", msg, s.kernel, s.blockIdx.x, s.cur ? s.cur->lin : 0);
    abort();
}
#ifdef EMU_ASM_SWITCH
inline void to_sched(Fiber* f) { emu_switch_ctx(&f->sp, S().sched_sp); }
inline void to_fiber(Fiber* f) { emu_switch_ctx(&S().sched_sp, f->sp); }
#else
inline void to_sched(Fiber* f) { swapcontext(&f->ctx, &S().sched); }
inline void to_fiber(Fiber* f) { swapcontext(&S().sched, &f->ctx); }
#endif
inline void yield() { to_sched(S().cur); }
inline void entry() {
    State& s = S();
    s.thunk(s.body);
    s.cur->done = true;
    s.live--;
    s.progress++;
    to_sched(s.cur);
}

template <class F> void thunk_of(const void* p) { (*static_cast<const F*>(p))(); }

template <class F> void launch(dim3 grid, dim3 block, size_t smem, const F& body) {
    State& s = S();
    if (s.cur) die("nested launch");
    const unsigned nt = block.x * block.y * block.z;
    if (!nt || nt > 1024) die("bad block size");
    if (smem > DYN_SMEM) die("dynamic shared memory too large");
    if (!s.dyn_smem) s.dyn_smem = (unsigned char*)aligned_alloc(128, DYN_SMEM);
    if (s.fib.size() < nt) s.fib.resize(nt);
    while (s.stacks.size() < nt) s.stacks.push_back((char*)aligned_alloc(64, STACK));
    s.blockDim = block; s.gridDim = grid; s.nthreads = nt;
    s.body = &body; s.thunk = &thunk_of<F>;
    s.launches++;
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        s.blockIdx = uint3{bx, by, bz};
        for (auto& b : s.bar_arrived) b = 0;
        memset(s.dyn_smem, getenv("SIMT_EMU_POISON") ? (int)strtol(getenv("SIMT_EMU_POISON"), nullptr, 0) : 0xCD, smem);
        for (unsigned t = 0; t < nt; t++) {
            Fiber& f = s.fib[t];
#ifdef EMU_ASM_SWITCH
            {
                uintptr_t top = ((uintptr_t)s.stacks[t] + STACK) & ~(uintptr_t)15;
                void** P = (void**)(top - 16);     // `ret` pops the entry address from P; entry then sees rsp = 8 mod 16
                P[0] = (void*)&entry; P[1] = nullptr;
                void** sp = P - 6;                 // r15 r14 r13 r12 rbx rbp
                for (int k = 0; k < 6; k++) sp[k] = nullptr;
                f.sp = sp;
            }
#else
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = s.stacks[t];
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())entry, 0);
#endif
            f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            f.lin = t; f.done = false; f.waiting = false; f.c_wait = f.c_done = false; f.stc_calls = 0;
        }
        s.live = nt;
        s.fibers_run += nt;
        while (s.live) {
            const uint64_t p0 = s.progress;
            bool ran_free = false;
            for (unsigned t = 0; t < nt; t++) {
                Fiber& f = s.fib[t];
                if (f.done) continue;
                if (!f.waiting) ran_free = true;
                s.cur = &f;
                to_fiber(&f);
            }
            s.cur = nullptr;
            if (s.live && !ran_free && s.progress == p0) {
                for (unsigned t = 0; t < nt; t++) {
                    const Fiber& f = s.fib[t];
                    if (!f.done && f.c_wait) fprintf(stderr, "simt_emu: thread %u waits at %s:%u (op %d, mask %08x)\n", t, f.c_file, f.c_line, f.c_op, f.c_mask);
                    else if (!f.done) fprintf(stderr, "simt_emu: thread %u waits at a barrier / mbarrier\n", t);
                }
                s.cur = &s.fib[0];
                die("deadlock: every live thread waits and nothing completes");
            }
        }
    }
    s.cur = nullptr;
}

// barrier `id` over `count` threads (0 = all live threads of the block)
inline void barrier(int id, unsigned count) {
    State& s = S();
    Fiber* f = s.cur;
    const unsigned gen = s.bar_gen[id];
    s.bar_arrived[id]++;
    for (;;) {
        if (s.bar_gen[id] != gen) break;
        const unsigned need = count ? count : s.live;
        if (s.bar_arrived[id] >= need) { s.bar_arrived[id] = 0; s.bar_gen[id]++; s.progress++; break; }
        f->waiting = true; yield(); f->waiting = false;
    }
}

enum { OP_SHFL, OP_SHFL_XOR, OP_SHFL_UP, OP_SHFL_DOWN, OP_BALLOT, OP_ADD, OP_MIN, OP_MAX, OP_SMIN, OP_SMAX, OP_MATCH, OP_SYNC };

typedef std::source_location SL;
inline uint64_t collective(uint32_t mask, int op, uint64_t val, int arg, int width, const SL& loc) {
    State& s = S();
    Fiber* f = s.cur;
    const unsigned w0 = f->lin & ~31u, lane = f->lin & 31u;
    const unsigned wn = (s.nthreads - w0) < 32 ? (s.nthreads - w0) : 32;
    if (!(mask >> lane & 1)) die("collective: the calling lane is not in its own mask");
    if (wn < 32) mask &= (1u << wn) - 1;
    f->c_wait = true; f->c_done = false; f->c_mask = mask; f->c_op = op; f->c_val = val; f->c_arg = arg; f->c_file = loc.file_name(); f->c_line = loc.line();
    for (;;) {
        if (f->c_done) break;
        bool all = true;
        for (unsigned l = 0; l < wn && all; l++)
            if (mask >> l & 1) {
                Fiber& g = s.fib[w0 + l];
                if (g.done) die("collective: a lane named in the mask has exited");
                if (!g.c_wait || g.c_done) all = false;
                // a lane waiting at ANOTHER collective (other mask, operation or call site) has not arrived at this one: legal when
                // its mask does not name this lane (disjoint groups of a diverged warp); if it does, both wait for ever and
                // the scheduler's deadlock check reports the two call sites
                else if (g.c_mask != mask || g.c_op != op || g.c_line != f->c_line || g.c_file != f->c_file) all = false;
            }
        if (all) {
            uint64_t red = 0; bool first = true;
            for (unsigned l = 0; l < wn; l++) if (mask >> l & 1) {
                const uint64_t v = s.fib[w0 + l].c_val;
                switch (op) {
                    case OP_BALLOT: if (v) red |= 1ull << l; break;
                    case OP_ADD: red = (uint32_t)(red + v); break;
                    case OP_MIN: red = first ? v : ((uint32_t)v < (uint32_t)red ? v : red); break;
                    case OP_MAX: red = first ? v : ((uint32_t)v > (uint32_t)red ? v : red); break;
                    case OP_SMIN: red = first ? v : ((int32_t)v < (int32_t)red ? v : red); break;
                    case OP_SMAX: red = first ? v : ((int32_t)v > (int32_t)red ? v : red); break;
                    default: break;
                }
                first = false;
            }
            for (unsigned l = 0; l < wn; l++) if (mask >> l & 1) {
                Fiber& g = s.fib[w0 + l];
                const int a = g.c_arg;
                const int seg = (int)l & ~(width - 1);
                int src = (int)l;
                uint64_t r = 0;
                switch (op) {
                    case OP_SHFL: src = seg + (a & (width - 1)); break;
                    case OP_SHFL_XOR: src = (int)l ^ a; if (src >= seg + width) src = (int)l; break;
                    case OP_SHFL_UP: src = (int)l - a; if (src < seg) src = (int)l; break;
                    case OP_SHFL_DOWN: src = (int)l + a; if (src >= seg + width) src = (int)l; break;
                    case OP_MATCH:
                        for (unsigned k = 0; k < wn; k++) if ((mask >> k & 1) && s.fib[w0 + k].c_val == g.c_val) r |= 1ull << k;
                        break;
                    case OP_SYNC: break;
                    default: r = red; break;
                }
                if (op <= OP_SHFL_DOWN) {
                    // reading an inactive / out-of-mask lane is undefined on hardware; the emulator returns the lane's own
                    // value for lanes outside the warp and complains for lanes outside the mask
                    if (src < 0 || src >= (int)wn) src = (int)l;
                    if (!(mask >> src & 1)) r = 0xDEADBEEFDEADBEEFull;   // poison: the value is undefined on hardware
                    else r = s.fib[w0 + src].c_val;
                }
                g.c_res = r;
            }
            for (unsigned l = 0; l < wn; l++) if (mask >> l & 1) { s.fib[w0 + l].c_done = true; }
            s.progress++;
            break;
        }
        f->waiting = true; yield(); f->waiting = false;
    }
    f->c_wait = false; f->c_done = false;
    return f->c_res;
}

template <class T> inline uint64_t to_bits(T v) { static_assert(sizeof(T) <= 8, "shuffle of a wide type"); uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

// mbarrier emulation (one 64-bit word in shared memory): bit 0 phase, bits 1-15 pending arrivals, bits 16-31 expected
// arrivals, bits 32-63 outstanding transaction bytes (signed)
struct MBar { uint32_t phase : 1, pending : 15, count : 16; int32_t tx; };
static_assert(sizeof(MBar) == 8, "mbarrier word");
inline void mbar_check(MBar* b) {
    if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->count; S().progress++; }
}
inline void mbar_init(unsigned long long* bar, uint32_t count) { MBar* b = (MBar*)bar; b->phase = 0; b->pending = count; b->count = count; b->tx = 0; }
inline void mbar_expect(unsigned long long* bar, uint32_t bytes) { MBar* b = (MBar*)bar; b->tx += (int32_t)bytes; if (!b->pending) die("mbarrier: more arrivals than its count"); b->pending--; mbar_check(b); }
inline void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    if (bytes % 16 || (uintptr_t)dst % 16 || (uintptr_t)src % 16) die("cp.async.bulk: size and both addresses must be multiples of 16");
    memcpy(dst, src, bytes);
    MBar* b = (MBar*)bar; b->tx -= (int32_t)bytes; mbar_check(b);
}
inline void mbar_wait(unsigned long long* bar, uint32_t parity) {
    MBar* b = (MBar*)bar;
    Fiber* f = S().cur;
    while (b->phase == (parity & 1)) { f->waiting = true; yield(); f->waiting = false; }
}
}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().blockIdx)
#define blockDim (emu::S().blockDim)
#define gridDim (emu::S().gridDim)
#define EMU_LAUNCH(grid, block, smem, stream, ...) emu::S().kernel = #__VA_ARGS__, emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { __VA_ARGS__; })
#define EMU_DYN_SMEM(T) reinterpret_cast<T*>(emu::S().dyn_smem)

static inline void __syncthreads() { emu::barrier(0, 0); }
// block-wide vote: every thread adds its predicate to one of two alternating accumulators, a barrier, everybody reads, a second
// barrier, everybody clears (a thread can only reach the accumulator's next use after all threads have passed the call in between)
static inline int __syncthreads_count(int pred) {
    emu::State& s = emu::S();
    const unsigned slot = s.cur->stc_calls++ & 1;
    s.stc_acc[slot] += pred ? 1 : 0;
    emu::barrier(0, 0);
    const int r = s.stc_acc[slot];
    emu::barrier(0, 0);
    s.stc_acc[slot] = 0;
    return r;
}
static inline int __syncthreads_and(int pred) { const int live = (int)emu::S().live; return __syncthreads_count(pred) == live; }
static inline int __syncthreads_or(int pred) { return __syncthreads_count(pred) != 0; }
static inline void __syncwarp(unsigned mask = 0xffffffffu, const emu::SL& loc = emu::SL::current()) { emu::collective(mask, emu::OP_SYNC, 0, 0, 32, loc); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int w = 32, const emu::SL& loc = emu::SL::current()) { return emu::from_bits<T>(emu::collective(m, emu::OP_SHFL, emu::to_bits(v), src, w, loc)); }
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int w = 32, const emu::SL& loc = emu::SL::current()) { return emu::from_bits<T>(emu::collective(m, emu::OP_SHFL_XOR, emu::to_bits(v), x, w, loc)); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int w = 32, const emu::SL& loc = emu::SL::current()) { return emu::from_bits<T>(emu::collective(m, emu::OP_SHFL_UP, emu::to_bits(v), (int)d, w, loc)); }
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int w = 32, const emu::SL& loc = emu::SL::current()) { return emu::from_bits<T>(emu::collective(m, emu::OP_SHFL_DOWN, emu::to_bits(v), (int)d, w, loc)); }
static inline unsigned __ballot_sync(unsigned m, int p, const emu::SL& loc = emu::SL::current()) { return (unsigned)emu::collective(m, emu::OP_BALLOT, p != 0, 0, 32, loc); }
static inline int __any_sync(unsigned m, int p, const emu::SL& loc = emu::SL::current()) { return emu::collective(m, emu::OP_BALLOT, p != 0, 0, 32, loc) != 0; }
static inline int __all_sync(unsigned m, int p, const emu::SL& loc = emu::SL::current()) { return emu::collective(m, emu::OP_BALLOT, p == 0, 0, 32, loc) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v, const emu::SL& loc = emu::SL::current()) { return (unsigned)emu::collective(m, emu::OP_ADD, v, 0, 32, loc); }
static inline int __reduce_add_sync(unsigned m, int v, const emu::SL& loc = emu::SL::current()) { return (int)(unsigned)emu::collective(m, emu::OP_ADD, (unsigned)v, 0, 32, loc); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v, const emu::SL& loc = emu::SL::current()) { return (unsigned)emu::collective(m, emu::OP_MIN, v, 0, 32, loc); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v, const emu::SL& loc = emu::SL::current()) { return (unsigned)emu::collective(m, emu::OP_MAX, v, 0, 32, loc); }
static inline int __reduce_min_sync(unsigned m, int v, const emu::SL& loc = emu::SL::current()) { return (int)(unsigned)emu::collective(m, emu::OP_SMIN, (unsigned)v, 0, 32, loc); }
static inline int __reduce_max_sync(unsigned m, int v, const emu::SL& loc = emu::SL::current()) { return (int)(unsigned)emu::collective(m, emu::OP_SMAX, (unsigned)v, 0, 32, loc); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v, const emu::SL& loc = emu::SL::current()) { unsigned r = 0; for (int b = 0; b < 32; b++) if (emu::collective(m, emu::OP_BALLOT, v >> b & 1, 0, 32, loc)) r |= 1u << b; return r; }
template <class T> static inline unsigned __match_any_sync(unsigned m, T v, const emu::SL& loc = emu::SL::current()) { return (unsigned)emu::collective(m, emu::OP_MATCH, emu::to_bits(v), 0, 32, loc); }

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (sh & 31)) >> 32); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh) { sh = sh > 32 ? 32 : sh; return (unsigned)((((uint64_t)hi << 32) | lo) >> sh); }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned sh) { sh = sh > 32 ? 32 : sh; return (unsigned)(((((uint64_t)hi << 32) | lo) << sh) >> 32); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= (v >> i & 1) << (31 - i); return r; }
namespace emu {
// prmt.b32, default mode: nibble bits 0..2 select a byte of {b, a}, bit 3 replaces it by its sign (0x00 / 0xFF)
static inline unsigned prmt(unsigned a, unsigned b, unsigned s) {
    const uint64_t ab = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned byte = (unsigned)(ab >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
}
// the CUDA intrinsic uses only three bits of each selector nibble (nvcc masks the rest away: seen in the SASS, round 2)
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) { return emu::prmt(a, b, s & 0x7777u); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __mulhi(int a, int b) { return (int)(((int64_t)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }

// arithmetic with stated rounding: the translation unit is built with -ffp-contract=off, so plain operators do not fuse
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
// cvt.rzi.s32.f64 on silicon: a NaN with an .f64 source converts to MININT, not 0 (PTX ISA, cvt; seen on B200 in round 1)
static inline int __double2int_rz(double d) { if (d != d) return (-2147483647 - 1); if (d >= 2147483647.0) return 2147483647; if (d <= -2147483648.0) return (-2147483647 - 1); return (int)d; }
static inline int __float2int_rz(float d) { if (d != d) return 0; return __double2int_rz((double)d); }
static inline int __float2int_rn(float d) { if (d != d) return 0; return __double2int_rz(nearbyint((double)d)); }
static inline float __int2float_rn(int i) { return (float)i; }
static inline float __uint2float_rn(unsigned i) { return (float)i; }
static inline double __int2double_rn(int i) { return (double)i; }
static inline float __saturatef(float x) { return x != x ? 0.f : (x < 0.f ? 0.f : (x > 1.f ? 1.f : x)); }

// CUDA's integer min / max overload set (mixed signedness promotes to unsigned)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

// atomics: blocks and threads are interleaved on one OS thread, so plain read-modify-write is atomic
template <class T, class U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicSub(T* p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U> static inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> static inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T* p, U c, V v) { T o = *p; if (o == (T)c) *p = (T)v; return o; }

// ---------------------------------------------------------------------------------------------------------------------
// runtime API: one device, synchronous streams, host memory
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
typedef struct emu_stream* cudaStream_t;
struct emu_event { double t; };
typedef emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaStreamDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "simt_emu error"); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    const char* e = getenv("SIMT_EMU_SMS");
    *v = a == cudaDevAttrMultiProcessorCount ? (e ? atoi(e) : 148) : 0;
    return cudaSuccess;
}
static inline cudaError_t emu_alloc(void** p, size_t n) {
    void* q = nullptr;
    if (posix_memalign(&q, 256, n ? n : 1)) { *p = nullptr; return cudaErrorMemoryAllocation; }
    // fresh device memory holds no particular value: poison it (SIMT_EMU_POISON=<byte>, default 0xA5), so that a result that
    // depends on memory nobody wrote changes with the poison; in page-sized strides only beyond 64 MiB (the big scratch
    // buffers of the engine are mostly never touched)
    static const int poison = getenv("SIMT_EMU_POISON") ? (int)strtol(getenv("SIMT_EMU_POISON"), nullptr, 0) : 0xA5;
    if (n <= (64u << 20)) memset(q, poison, n);
    else for (size_t i = 0; i < n; i += 4096) ((unsigned char*)q)[i] = (unsigned char)poison;
    *p = q;
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return emu_alloc((void**)p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return emu_alloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t ev, unsigned = 0) {
    // streams run in issue order, so a wait is satisfied iff the event was recorded earlier in program order
    if (ev && ev->t < 0) { fprintf(stderr, "simt_emu: cudaStreamWaitEvent on an event that was never recorded\n"); }
    return cudaSuccess;
}
static inline double emu_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event{-1.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emu_now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
// a query may always answer "not yet": SIMT_EMU_QUERY_NOT_READY=1 makes it do so, which drives the host code down the paths it
// takes when the device is still busy (e.g. batches planned while later index groups are outstanding)
static inline cudaError_t cudaEventQuery(cudaEvent_t) { static const bool nr = getenv("SIMT_EMU_QUERY_NOT_READY") != nullptr; return nr ? cudaErrorNotReady : cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    if (!a || !b || a->t < 0 || b->t < 0) { *ms = 0; return cudaErrorInvalidValue; }
    *ms = (float)(b->t - a->t);
    return cudaSuccess;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
