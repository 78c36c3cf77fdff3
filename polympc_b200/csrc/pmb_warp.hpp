// pmb_warp.hpp — the thread/warp vocabulary every kernel body in this engine is written against.
//
// Under nvcc `Warp` is an empty struct whose members are the CUDA warp intrinsics (zero overhead) and `launch<Body>()`
// starts a real kernel.  When the same sources are compiled by g++ with -DPMB_EMU (tests/warp_emu only — TEST
// INFRASTRUCTURE, never part of the shipped library) `Warp` is backed by a lock-step fiber scheduler so that the very
// same kernel bodies can be executed on a CPU in the `-m "not gpu"` suite and compared bit for bit with the oracle.
// The product library is always the nvcc build; it has no CPU execution path.
//
// Rules for kernel bodies (they make both back ends agree):
//   * every w.sync()/w.shfl*/w.ballot/w.block_sync() is executed by ALL threads of the block, converged;
//   * a kernel body is a struct with `static constexpr int THREADS`, `MIN_BLOCKS` (resident CTAs per SM the register
//     allocation must allow) and
//       template-free  static PMB_DEV void run(const Warp& w, int block, unsigned char* smem, Args... args);
#pragma once
#include <cstdint>
#include <cstddef>
#include "pmb_detmath.h"

#if defined(__CUDACC__) && !defined(PMB_EMU)
// =====================================================================================================================
#include <cuda_runtime.h>
#define PMB_DEV __device__ __forceinline__
#define PMB_DEV_NOINLINE __device__ __noinline__
#define PMB_UNROLL _Pragma("unroll")
#define PMB_NOUNROLL _Pragma("unroll 1")
#define PMB_UNROLL4 _Pragma("unroll 4")

namespace pmb {

struct Warp {
    PMB_DEV int lane() const { return (int)(threadIdx.x & 31u); }
    PMB_DEV int tid() const { return (int)threadIdx.x; }
    PMB_DEV int nthreads() const { return (int)blockDim.x; }
    PMB_DEV int warp_id() const { return (int)(threadIdx.x >> 5); }
    PMB_DEV void sync() const { __syncwarp(); }
    PMB_DEV void block_sync() const { __syncthreads(); }
    PMB_DEV double shfl(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    PMB_DEV int shfl(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    PMB_DEV double shfl_xor(double v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    PMB_DEV int shfl_xor(int v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    PMB_DEV double shfl_down(double v, int d) const { return __shfl_down_sync(0xffffffffu, v, d); }
    PMB_DEV unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    PMB_DEV bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
    PMB_DEV bool all(bool p) const { return __all_sync(0xffffffffu, p) != 0; }
    PMB_DEV unsigned long long clock() const { return (unsigned long long)clock64(); }
    /** warp-wide integer reductions (REDUX.SYNC) */
    PMB_DEV unsigned reduce_max(unsigned v) const { return __reduce_max_sync(0xffffffffu, v); }
    PMB_DEV unsigned reduce_min(unsigned v) const { return __reduce_min_sync(0xffffffffu, v); }
    /** fp64 tensor-core tile product D = A (8 x 4) B (4 x 8) + C (8 x 8): lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and
     *  C[l >> 2][2 (l & 3) + {0, 1}]  (mma.sync.m8n8k4.f64, SASS DMMA) */
    PMB_DEV void dmma(double& c0, double& c1, double a, double b) const
    {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
};

PMB_DEV int atomic_add(int* p, int v) { return atomicAdd(p, v); }
PMB_DEV void atomic_add_u64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }

template <class Body, class... Args>
__global__ void __launch_bounds__(Body::THREADS, Body::MIN_BLOCKS) pmb_kernel(Args... args)
{
    extern __shared__ __align__(16) unsigned char pmb_smem[];
    Warp w;
    Body::run(w, (int)blockIdx.x, pmb_smem, args...);
}

/** launch `grid` blocks of Body::THREADS threads with `smem` bytes of dynamic shared memory on `stream` */
template <class Body, class... Args>
inline cudaError_t launch(int grid, size_t smem, cudaStream_t stream, Args... args)
{
    if (grid <= 0) return cudaSuccess;
    if (smem > 48 * 1024) {   // per device and cheap: set unconditionally (a cached flag would be wrong on a second device)
        cudaError_t e = cudaFuncSetAttribute(pmb_kernel<Body, Args...>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    pmb_kernel<Body, Args...><<<grid, Body::THREADS, smem, stream>>>(args...);
    return cudaGetLastError();
}

/** number of CTAs of this kernel that are resident at once on the current device (SMs x CTAs per SM): the grid of a
 *  persistent launch.  0 on error. */
template <class Body, class... Args>
inline int resident_ctas(size_t smem, Args...)
{
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(pmb_kernel<Body, Args...>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pmb_kernel<Body, Args...>, Body::THREADS, smem) != cudaSuccess) return 0;
    return sms * per_sm;
}

} // namespace pmb

#else
// =====================================================================================================================
// Lock-step emulation (tests/warp_emu).  One fiber per thread of a block; every sync-type call ends a phase.
#ifndef PMB_EMU
#error "pmb_warp.hpp: compile with nvcc, or with -DPMB_EMU for the test-only warp emulator"
#endif
#include <ucontext.h>
#include <vector>
#include <functional>
#include <cstdlib>
#include <cstring>
#define PMB_DEV inline
#define PMB_DEV_NOINLINE
#define PMB_UNROLL
#define PMB_NOUNROLL
#define PMB_UNROLL4

namespace pmb {

struct EmuBlock {
    int nthreads = 0, cur = 0;
    std::vector<ucontext_t> ctx;
    ucontext_t main_ctx;
    std::vector<char> done;
    std::vector<uint64_t> slot, slot2;
    std::vector<unsigned char> pred;
    struct Bar { int count = 0; int gen = 0; };
    Bar block_bar;
    std::vector<Bar> warp_bar;
    char* stacks = nullptr;
    std::function<void(int)> body;
    static EmuBlock*& current() { static thread_local EmuBlock* p = nullptr; return p; }
    static void trampoline()
    {
        EmuBlock* b = current();
        const int t = b->cur;
        b->body(t);
        b->done[t] = 1;
        // returning activates uc_link (main_ctx)
    }
    void yield() { const int t = cur; swapcontext(&ctx[t], &main_ctx); }
    /** true barrier among `participants` fibers: the last arrival releases the generation, the others yield until then */
    void barrier(Bar& b, int participants, int t)
    {
        const int g = b.gen;
        if (++b.count == participants) { b.count = 0; ++b.gen; return; }
        while (b.gen == g) { cur = t; yield(); }
    }
    void warp_barrier(int t) { const int wi = t >> 5; const int cnt = (nthreads - (wi << 5)) >= 32 ? 32 : (nthreads - (wi << 5)); barrier(warp_bar[wi], cnt, t); }
    void block_barrier(int t) { barrier(block_bar, nthreads, t); }
    void run(int n, size_t stack_bytes, std::function<void(int)> f)
    {
        nthreads = n; body = std::move(f);
        ctx.assign(n, ucontext_t()); done.assign(n, 0); slot.assign(n, 0); slot2.assign(n, 0); pred.assign(n, 0);
        block_bar = Bar(); warp_bar.assign((n + 31) / 32, Bar());
        stacks = (char*)std::malloc(stack_bytes * (size_t)n);
        EmuBlock* prev = current();
        current() = this;
        for (int t = 0; t < n; ++t) {
            getcontext(&ctx[t]);
            ctx[t].uc_stack.ss_sp = stacks + stack_bytes * (size_t)t;
            ctx[t].uc_stack.ss_size = stack_bytes;
            ctx[t].uc_link = &main_ctx;
            makecontext(&ctx[t], (void (*)())&EmuBlock::trampoline, 0);
        }
        bool alive = true;
        while (alive) {
            alive = false;
            for (int t = 0; t < n; ++t) {
                if (done[t]) continue;
                cur = t;
                swapcontext(&main_ctx, &ctx[t]);
                if (!done[t]) alive = true;
            }
        }
        current() = prev;
        std::free(stacks); stacks = nullptr;
    }
};

struct Warp {
    EmuBlock* b; int t;
    int lane() const { return t & 31; }
    int tid() const { return t; }
    int nthreads() const { return b->nthreads; }
    int warp_id() const { return t >> 5; }
    void sync() const { b->warp_barrier(t); }
    void block_sync() const { b->block_barrier(t); }
    uint64_t xchg(uint64_t bits, int src_lane) const
    {
        b->slot[t] = bits;
        sync();
        int s = (t & ~31) | (src_lane & 31);
        if (s >= b->nthreads) s = t;
        const uint64_t r = b->slot[s];
        sync();
        return r;
    }
    double shfl(double v, int src) const { return dm::from_bits(xchg(dm::to_bits(v), src)); }
    int shfl(int v, int src) const { return (int)(int64_t)xchg((uint64_t)(int64_t)v, src); }
    double shfl_xor(double v, int m) const { return shfl(v, lane() ^ m); }
    int shfl_xor(int v, int m) const { return shfl(v, lane() ^ m); }
    double shfl_down(double v, int d) const { return (lane() + d < 32) ? shfl(v, lane() + d) : (shfl(v, lane()), v); }
    unsigned ballot(bool p) const
    {
        b->pred[t] = p ? 1 : 0;
        sync();
        unsigned m = 0;
        const int base = t & ~31;
        for (int l = 0; l < 32 && base + l < b->nthreads; ++l) if (b->pred[base + l]) m |= (1u << l);
        sync();
        return m;
    }
    unsigned long long clock() const { return 0; }
    unsigned reduce_max(unsigned v) const
    {
        b->slot[t] = v;
        sync();
        unsigned m = 0;
        const int base = t & ~31;
        for (int l = 0; l < 32 && base + l < b->nthreads; ++l) { const unsigned o = (unsigned)b->slot[base + l]; if (o > m) m = o; }
        sync();
        return m;
    }
    unsigned reduce_min(unsigned v) const { return ~reduce_max(~v); }
    /** emulated mma.m8n8k4.f64: c += sum_kk a[r][kk] b[kk][n] as an ascending fused chain (the hardware's internal order is not
     *  specified; the fast-arithmetic path is tolerance-checked, not bit-checked) */
    void dmma(double& c0, double& c1, double a, double bv) const
    {
        b->slot[t] = dm::to_bits(a); b->slot2[t] = dm::to_bits(bv);
        sync();
        const int base = t & ~31, l = t & 31, r = l >> 2, q = l & 3;
        for (int i = 0; i < 2; ++i) {
            const int n = 2 * q + i;
            double acc = i ? c1 : c0;
            for (int kk = 0; kk < 4; ++kk)
                acc = dm::fma(dm::from_bits(b->slot[base + r * 4 + kk]), dm::from_bits(b->slot2[base + n * 4 + kk]), acc);
            (i ? c1 : c0) = acc;
        }
        sync();
    }
    bool any(bool p) const { return ballot(p) != 0; }
    bool all(bool p) const { const unsigned m = ballot(p); const int cnt = (b->nthreads - (t & ~31)) >= 32 ? 32 : (b->nthreads - (t & ~31)); return m == (cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u)); }
};

inline int atomic_add(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline void atomic_add_u64(unsigned long long* p, unsigned long long v) { *p += v; }

typedef void* cudaStream_t_emu;

template <class Body, class... Args>
inline int launch(int grid, size_t smem, void* /*stream*/, Args... args)
{
    std::vector<unsigned char> sm(smem + 64);
    unsigned char* smp = sm.data() + ((16 - ((uintptr_t)sm.data() & 15)) & 15);
    for (int blk = grid - 1; blk >= 0; --blk) {   // last block first: persistent kernels then run with a rotated warp numbering
        EmuBlock eb;
        eb.run(Body::THREADS, Body::EMU_STACK_BYTES, [&](int t) {
            Warp w{&eb, t};
            Body::run(w, blk, smp, args...);
        });
    }
    return 0;
}

/** emulator: a handful of "resident" CTAs (they run one after the other; the first drains the queue) */
template <class Body, class... Args>
inline int resident_ctas(size_t, Args...) { return 3; }

} // namespace pmb
#endif
