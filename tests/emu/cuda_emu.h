// cuda_emu.h — a small fiber-based SIMT emulator.  TEST INFRASTRUCTURE ONLY.
//
// Lets tests/ compile zrk_modulation_b200/csrc/*.cu{h} with g++ (-DZRK_EMU) and run the very
// same kernel source on the CPU so that the host-side logic (state build, log reconstruction,
// the per-tick control flow) is covered by `pytest -m "not gpu"` in a container without a GPU.
// Every CUDA thread is a ucontext fiber; barriers and warp collectives are cooperative yield
// points.  It is never built into, loaded by, or a fallback for the product library.
#ifndef ZRK_CUDA_EMU_H
#define ZRK_CUDA_EMU_H

#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

struct uint4 {
    unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline long long __double_as_longlong(double d) {
    long long r;
    memcpy(&r, &d, 8);
    return r;
}
struct double2 {
    double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct float4 {
    float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace zrk_emu {

struct Barrier {
    int count = 0, gen = 0;
};
struct Warp {
    Barrier bar;
    uint64_t vals[32];
};
struct Block {
    std::vector<ucontext_t> ctx;
    std::vector<char *> stacks;
    std::vector<char> done;
    std::vector<Warp> warps;
    Barrier named[17];
    unsigned char *smem = nullptr;
    int n_threads = 0, n_done = 0;
    ucontext_t main_ctx;
    int cur = 0;
    const std::function<void()> *body = nullptr;
};

extern Block *g_block;
extern dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

inline void yield() {
    Block *b = g_block;
    swapcontext(&b->ctx[b->cur], &b->main_ctx);
}
inline void barrier_wait(Barrier &bar, int expected) {
    int gen = bar.gen;
    if (++bar.count >= expected) {
        bar.count = 0;
        bar.gen++;
    } else {
        while (bar.gen == gen) yield();
    }
}
inline void named_barrier(int id, int n) { barrier_wait(g_block->named[id], n); }
inline unsigned char *dyn_smem() { return g_block->smem; }
inline Warp &my_warp() { return g_block->warps[g_block->cur >> 5]; }
inline int my_lane() { return g_block->cur & 31; }

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);

}  // namespace zrk_emu

#define threadIdx zrk_emu::g_threadIdx
#define blockIdx zrk_emu::g_blockIdx
#define blockDim zrk_emu::g_blockDim
#define gridDim zrk_emu::g_gridDim

inline void __syncthreads() {
    zrk_emu::Block *b = zrk_emu::g_block;
    zrk_emu::barrier_wait(b->named[0], b->n_threads - b->n_done);
}
inline int __syncthreads_or(int pred) {
    zrk_emu::Block *b = zrk_emu::g_block;
    static int acc[2];
    const int slot = b->named[0].gen & 1;  // the accumulator of the barrier generation being formed
    if (b->named[0].count == 0) acc[slot] = 0;
    acc[slot] |= pred ? 1 : 0;
    __syncthreads();
    return acc[slot];
}
inline void __syncwarp(unsigned = 0xffffffffu) { zrk_emu::barrier_wait(zrk_emu::my_warp().bar, 32); }

template <typename T>
inline T zrk_emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    zrk_emu::Warp &w = zrk_emu::my_warp();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.vals[zrk_emu::my_lane()] = raw;
    zrk_emu::barrier_wait(w.bar, 32);
    uint64_t got = w.vals[src_lane & 31];
    zrk_emu::barrier_wait(w.bar, 32);
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
    return zrk_emu_exchange(v, src);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    int l = zrk_emu::my_lane();
    T got = zrk_emu_exchange(v, l >= (int)d ? l - (int)d : l);
    return l >= (int)d ? got : v;
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
    return zrk_emu_exchange(v, zrk_emu::my_lane() ^ m);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    zrk_emu::Warp &w = zrk_emu::my_warp();
    w.vals[zrk_emu::my_lane()] = pred ? 1 : 0;
    zrk_emu::barrier_wait(w.bar, 32);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)(w.vals[i] & 1) << i;
    zrk_emu::barrier_wait(w.bar, 32);
    return r;
}
inline int atomicMin(int *p, int v) {
    int o = *p;
    if (v < o) *p = v;
    return o;
}
inline int atomicAdd(int *p, int v) {
    int o = *p;
    *p = o + v;
    return o;
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    unsigned long long o = *p;
    *p = o + v;
    return o;
}
inline int atomicExch(int *p, int v) {
    int o = *p;
    *p = v;
    return o;
}
inline int atomicOr(int *p, int v) {
    int o = *p;
    *p = o | v;
    return o;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __fdividef(float a, float b) { return a / b; }
inline int __reduce_min_sync(unsigned, int v) {
    zrk_emu::Warp &w = zrk_emu::my_warp();
    w.vals[zrk_emu::my_lane()] = (uint64_t)(int64_t)v;
    zrk_emu::barrier_wait(w.bar, 32);
    int m = v;
    for (int i = 0; i < 32; i++) {
        int x = (int)(int64_t)w.vals[i];
        if (x < m) m = x;
    }
    zrk_emu::barrier_wait(w.bar, 32);
    return m;
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    zrk_emu::Warp &w = zrk_emu::my_warp();
    w.vals[zrk_emu::my_lane()] = v;
    zrk_emu::barrier_wait(w.bar, 32);
    unsigned m = v;
    for (int i = 0; i < 32; i++) m = (unsigned)w.vals[i] < m ? (unsigned)w.vals[i] : m;
    zrk_emu::barrier_wait(w.bar, 32);
    return m;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    zrk_emu::Warp &w = zrk_emu::my_warp();
    w.vals[zrk_emu::my_lane()] = v;
    zrk_emu::barrier_wait(w.bar, 32);
    unsigned m = v;
    for (int i = 0; i < 32; i++) m = (unsigned)w.vals[i] > m ? (unsigned)w.vals[i] : m;
    zrk_emu::barrier_wait(w.bar, 32);
    return m;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    zrk_emu::Warp &w = zrk_emu::my_warp();
    w.vals[zrk_emu::my_lane()] = v;
    zrk_emu::barrier_wait(w.bar, 32);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)w.vals[i];
    zrk_emu::barrier_wait(w.bar, 32);
    return m;
}
inline int __reduce_add_sync(unsigned, int v) {
    zrk_emu::Warp &w = zrk_emu::my_warp();
    w.vals[zrk_emu::my_lane()] = (uint64_t)(int64_t)v;
    zrk_emu::barrier_wait(w.bar, 32);
    int m = 0;
    for (int i = 0; i < 32; i++) m += (int)(int64_t)w.vals[i];
    zrk_emu::barrier_wait(w.bar, 32);
    return m;
}
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline unsigned __match_any_sync(unsigned, int v) {
    zrk_emu::Warp &w = zrk_emu::my_warp();
    w.vals[zrk_emu::my_lane()] = (uint64_t)(int64_t)v;
    zrk_emu::barrier_wait(w.bar, 32);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if ((int)(int64_t)w.vals[i] == v) r |= 1u << i;
    zrk_emu::barrier_wait(w.bar, 32);
    return r;
}

#endif
