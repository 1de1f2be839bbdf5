// cuda_runtime.h — NOT the CUDA header: a host stand-in that lets g++ compile voxel-rs_b200/csrc/{traverse,kernels}.cuh.
//
// TEST INFRASTRUCTURE. tests/emu builds the product's kernel sources for the CPU (this directory is put first on the include path,
// so `#include <cuda_runtime.h>` in traverse.cuh finds this file) and runs the very same kernel code against the oracle without a
// GPU: every CUDA thread of a CTA is a fiber (ucontext) on one OS thread, one CTA at a time; warp collectives and __syncthreads
// are rendezvous points of those fibers. It checks the LOGIC of the kernels (work fetch, refill, wavefront hand-over, traversal,
// shading) — not their timing, memory model or PTX. The product never includes it: nvcc finds the real header.
#pragma once
#ifndef VX_HOST_EMULATION
#define VX_HOST_EMULATION 1
#endif
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ thread_local          /* all fibers of the (single) running CTA live on one OS thread; build with
                                            -fno-extern-tls-init (extern __shared__ arrays have no dynamic initialiser) */
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct emu_dim3 { unsigned x = 0, y = 0, z = 0; };

namespace emu {

struct Warp {
    uint64_t in[32];
    uint64_t out[2][32];
    unsigned arrived = 0, gen = 0, live = 32;
};
struct Cta {
    unsigned arrived = 0, gen = 0, live = 0;
};
struct Lane {
    emu_dim3 tid, bid, bdim, gdim;
    unsigned lane = 0;
    Warp* warp = nullptr;
    Cta* cta = nullptr;
    ucontext_t ctx;
    bool done = false;
    const unsigned* wait_word = nullptr;   // blocked until *wait_word != wait_value
    unsigned wait_value = 0;
};

extern Lane* cur;              // the fiber that is running
extern ucontext_t scheduler;   // where a blocked fiber returns to
extern char* smem_base;        // start of the CTA's dynamic shared memory (vx::smem_raw)
extern unsigned long long collectives, switches;

inline void block_on(const unsigned* word, unsigned value) {
    Lane* me = cur;
    me->wait_word = word; me->wait_value = value;
    ++switches;
    swapcontext(&me->ctx, &scheduler);
}

// Every live lane of the warp contributes one value; returns the 32 contributions (slot of an exited lane: its last one).
inline const uint64_t* exchange(uint64_t v) {
    Lane* me = cur;
    Warp* w = me->warp;
    w->in[me->lane] = v;
    const unsigned g = w->gen;
    if (++w->arrived == w->live) {
        std::memcpy(w->out[g & 1u], w->in, sizeof(w->in));
        w->arrived = 0;
        w->gen = g + 1;
        ++collectives;
    } else {
        while (w->gen == g) block_on(&w->gen, g);
    }
    return w->out[g & 1u];
}
inline void cta_barrier() {
    Lane* me = cur;
    Cta* c = me->cta;
    const unsigned g = c->gen;
    if (++c->arrived == c->live) { c->arrived = 0; c->gen = g + 1; }
    else while (c->gen == g) block_on(&c->gen, g);
}

}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::cur->bdim)
#define gridDim (emu::cur->gdim)

// ---- memory ----
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)((const char*)p - emu::smem_base); }
static inline uint32_t& vx_emu_shared_u32(uint32_t byte_address) { return *reinterpret_cast<uint32_t*>(emu::smem_base + byte_address); }
static inline void __syncthreads() { emu::cta_barrier(); }
static inline void __threadfence_system() {}
static inline void __threadfence() {}
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline void __nanosleep(unsigned) {}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicAnd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o & v; return o; }

// ---- warp collectives (full masks only, which is all the kernels use) ----
static inline unsigned __ballot_sync(unsigned, int pred) {
    const uint64_t* v = emu::exchange(pred ? 1u : 0u);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)(v[i] & 1u) << i;
    return m;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
template <typename T> static inline T __shfl_sync(unsigned, T value, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t bits = 0;
    std::memcpy(&bits, &value, sizeof(T));
    const uint64_t* v = emu::exchange(bits);
    T out;
    std::memcpy(&out, &v[src & 31], sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T value, int lane_mask) {
    return __shfl_sync(mask, value, (int)(emu::cur->lane ^ (unsigned)lane_mask));
}

// ---- arithmetic intrinsics ----
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u)); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
