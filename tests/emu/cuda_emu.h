// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A tiny CUDA-on-CPU execution shim.
//
// The container the CPU test-suite runs in has no GPU.  To exercise the *same kernel source* that ships in
// libmocap_b200.so (mocapv2_b200/csrc/*.cu) there, tests/emu/build_emu.py compiles those files with g++ and
// -DMOCAP_EMU, which makes csrc/common.cuh include this header instead of <cuda_runtime.h>.  Every CUDA
// thread of a block becomes an OS thread (blocks run one after the other), __syncthreads() is a real
// barrier, warp collectives exchange through a per-warp mailbox, atomics are __atomic builtins and "device
// memory" is host memory.  The result, tests/emu/_build/libmocap_emu.so, is loaded ONLY by tests/ (never by
// the mocapv2_b200 package, which refuses to run without a CUDA device) and is orders of magnitude slower
// than the oracle: it is a debugger for kernel logic, not a CPU path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
struct uint4 { uint32_t x, y, z, w; };
struct int4 { int32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline int4 make_int4(int32_t x, int32_t y, int32_t z, int32_t w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 1; return 0; }

typedef double* cudaEvent_t;
double emu_now_ms();
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new double(0.0); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { *e = emu_now_ms(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(*b - *a); return 0; }

namespace emu {
struct WarpBox {
    std::unique_ptr<std::barrier<>> bar;
    unsigned long long slot[32];
    int n;
};
struct Block {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<WarpBox> warps;
    int n_threads = 0;
};
extern Block g_block;
extern std::vector<unsigned char> g_dyn_smem;
extern thread_local uint3_emu t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
inline unsigned char* dyn_smem() { return g_dyn_smem.data(); }
inline int linear_tid() { return (int)(t_threadIdx.x + g_blockDim.x * (t_threadIdx.y + g_blockDim.y * t_threadIdx.z)); }
inline WarpBox& my_warp() { return g_block.warps[linear_tid() >> 5]; }
inline int my_lane() { return linear_tid() & 31; }
// all live lanes of the warp deposit v, then every lane may read any slot
inline void exchange(unsigned long long v, unsigned long long out[32], int* n)
{
    WarpBox& w = my_warp();
    w.slot[my_lane()] = v;
    w.bar->arrive_and_wait();
    for (int i = 0; i < w.n; ++i) out[i] = w.slot[i];
    *n = w.n;
    w.bar->arrive_and_wait();
}
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim

static inline void __syncthreads() { emu::g_block.bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::my_warp().bar->arrive_and_wait(); }

static inline unsigned __ballot_sync(unsigned, int pred)
{
    unsigned long long s[32]; int n;
    emu::exchange(pred ? 1ull : 0ull, s, &n);
    unsigned m = 0;
    for (int i = 0; i < n; ++i) if (s[i]) m |= 1u << i;
    return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __reduce_min_sync(unsigned, int v)
{
    unsigned long long s[32]; int n;
    emu::exchange((unsigned long long)(long long)v, s, &n);
    int r = v;
    for (int i = 0; i < n; ++i) r = std::min(r, (int)(long long)s[i]);
    return r;
}
static inline int __reduce_max_sync(unsigned, int v)
{
    unsigned long long s[32]; int n;
    emu::exchange((unsigned long long)(long long)v, s, &n);
    int r = v;
    for (int i = 0; i < n; ++i) r = std::max(r, (int)(long long)s[i]);
    return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
    unsigned long long s[32]; int n;
    emu::exchange(v, s, &n);
    unsigned r = 0;
    for (int i = 0; i < n; ++i) r |= (unsigned)s[i];
    return r;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src)
{
    static_assert(sizeof(T) <= 8, "shfl payload");
    unsigned long long s[32], raw = 0; int n;
    memcpy(&raw, &v, sizeof(T));
    emu::exchange(raw, s, &n);
    T r; memcpy(&r, &s[src & 31], sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int delta)
{
    unsigned long long s[32], raw = 0; int n;
    memcpy(&raw, &v, sizeof(T));
    emu::exchange(raw, s, &n);
    int src = emu::my_lane() + delta;
    if (src >= n || src >= 32) return v;
    T r; memcpy(&r, &s[src], sizeof(T));
    return r;
}

template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int delta)
{
    unsigned long long s[32], raw = 0; int n;
    memcpy(&raw, &v, sizeof(T));
    emu::exchange(raw, s, &n);
    int src = emu::my_lane() - delta;
    if (src < 0) return v;
    T r; memcpy(&r, &s[src], sizeof(T));
    return r;
}

// ---- atomics ---------------------------------------------------------------------------------------------------
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMin(int* p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline int atomicMax(int* p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

// ---- intrinsics (build with -ffp-contract=off so that a*b+c is never fused) ---------------------------------------
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel)
{
    unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned n = (sel >> (4 * i)) & 0xf, b = (unsigned)(src >> (8 * (n & 7))) & 0xffu;
        if (n & 8) b = (b & 0x80u) ? 0xffu : 0u;          // sign-replicate mode
        r |= b << (8 * i);
    }
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh) { return sh >= 32 ? hi : (sh ? (lo >> sh) | (hi << (32 - sh)) : lo); }
static inline unsigned __dp2a_hi(unsigned a, unsigned b, unsigned c) { return c + (a & 0xffffu) * ((b >> 16) & 0xffu) + (a >> 16) * (b >> 24); }
static inline unsigned __dp2a_lo(unsigned a, unsigned b, unsigned c) { return c + (a & 0xffffu) * (b & 0xffu) + (a >> 16) * ((b >> 8) & 0xffu); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(v);
}
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline long long __double2ll_rn(double a) { return llrint(a); }
static inline int __double2int_rz(double a) { return (int)a; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline double rsqrt(double a) { return 1.0 / sqrt(a); }
static inline unsigned __vsetgtu4_emu(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int k = 0; k < 4; ++k) if (((a >> (8 * k)) & 0xff) > ((b >> (8 * k)) & 0xff)) r |= 0xffu << (8 * k);
    return r;
}
using std::abs;
using std::max;
using std::min;
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
