// TEST INFRASTRUCTURE ONLY — a tiny host-side CUDA execution emulator.
//
// The build container has no GPU, so kernel logic (indexing, scans, inter-CTA chaining)
// cannot be exercised there.  When the kernel sources are compiled with g++ and
// -DDMST_EMULATE this header supplies just enough of the CUDA programming model to run
// them on the CPU: each thread block runs as real OS threads with a barrier for
// __syncthreads(), warp shuffles go through a per-warp exchange buffer, blocks of a grid
// run one after another in launch order.  It exists to debug kernels before spending GPU
// time; the shipped library is never built this way and the product never loads the
// emulated library (diffmst_b200/_lib.py only loads libdiffmst_b200.so and refuses an
// emulated build).
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return {a, b}; }

using std::min;
using std::max;
typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0

namespace emul {
struct WarpCtx {
    uint32_t buf[32];
    std::unique_ptr<std::barrier<>> bar;
};
struct BlockCtx {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<WarpCtx> warps;
    std::vector<unsigned char> dyn_smem;
    int nthreads;
};
extern thread_local uint3_ t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern thread_local BlockCtx* t_block;
extern thread_local int t_static_smem_key;
}  // namespace emul

#define threadIdx (emul::t_threadIdx)
#define blockIdx (emul::t_blockIdx)
#define blockDim (emul::t_blockDim)
#define gridDim (emul::t_gridDim)
#define warpSize 32

static inline void __syncthreads() { emul::t_block->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    emul::t_block->warps[emul::t_threadIdx.x / 32].bar->arrive_and_wait();
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

namespace emul {
template <typename T>
static inline T shfl_generic(T v, int src_lane, bool valid) {
    static_assert(sizeof(T) == 4, "4-byte shuffles only");
    WarpCtx& w = t_block->warps[t_threadIdx.x / 32];
    int lane = t_threadIdx.x % 32;
    uint32_t raw;
    std::memcpy(&raw, &v, 4);
    w.buf[lane] = raw;
    w.bar->arrive_and_wait();
    uint32_t got = valid ? w.buf[src_lane & 31] : raw;
    w.bar->arrive_and_wait();
    T out;
    std::memcpy(&out, &got, 4);
    return out;
}
}  // namespace emul
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emul::shfl_generic(v, src, true); }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = emul::t_threadIdx.x % 32;
    return emul::shfl_generic(v, lane - (int)d, lane - (int)d >= 0);
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = emul::t_threadIdx.x % 32;
    return emul::shfl_generic(v, lane + (int)d, lane + (int)d < 32);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    int lane = emul::t_threadIdx.x % 32;
    return emul::shfl_generic(v, lane ^ m, true);
}

static inline unsigned __ballot_sync(unsigned, int pred) {
    emul::WarpCtx& w = emul::t_block->warps[emul::t_threadIdx.x / 32];
    int lane = emul::t_threadIdx.x % 32;
    w.buf[lane] = pred ? 1u : 0u;
    w.bar->arrive_and_wait();
    unsigned m = 0;
    int n = std::min(32, emul::t_block->nthreads - 32 * (int)(emul::t_threadIdx.x / 32));
    for (int i = 0; i < n; ++i) m |= (w.buf[i] & 1u) << i;
    w.bar->arrive_and_wait();
    return m;
}

static inline int atomicAdd(int* p, int v) {
    return reinterpret_cast<std::atomic<int>*>(p)->fetch_add(v);
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
    return reinterpret_cast<std::atomic<unsigned>*>(p)->fetch_add(v);
}
static inline int atomicMin(int* p, int v) {
    auto* a = reinterpret_cast<std::atomic<int>*>(p);
    int old = a->load();
    while (old > v && !a->compare_exchange_weak(old, v)) {}
    return old;
}
static inline int atomicMax(int* p, int v) {
    auto* a = reinterpret_cast<std::atomic<int>*>(p);
    int old = a->load();
    while (old < v && !a->compare_exchange_weak(old, v)) {}
    return old;
}
static inline float atomicAdd(float* p, float v) {
    auto* a = reinterpret_cast<std::atomic<uint32_t>*>(p);
    uint32_t old = a->load();
    for (;;) {
        float f;
        std::memcpy(&f, &old, 4);
        f += v;
        uint32_t nw;
        std::memcpy(&nw, &f, 4);
        if (a->compare_exchange_weak(old, nw)) break;
    }
    float r;
    std::memcpy(&r, &old, 4);
    return r;
}

static inline float emul_log2f(float x) { return std::log2(x); }
#define __log2f emul_log2f
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __ldg(const float* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline float __ldcg(const float* p) { return *p; }
static inline float4 __ldcg(const float4* p) { return *p; }
static inline void __stcg(float* p, float v) { *p = v; }
static inline void __stcg(float4* p, float4 v) { *p = v; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }

#define DMST_SHARED_ARRAY(type, name, count) \
    type* name = reinterpret_cast<type*>(emul::static_smem(__COUNTER__, sizeof(type) * (count)))

namespace emul {
// Static __shared__ arrays: one allocation per (block, call-site key).
void* static_smem(int key, size_t bytes);
unsigned char* dynamic_smem();

template <typename F>
void launch(dim3 grid, dim3 block, size_t smem, F&& body);
}  // namespace emul

// ----------------------------------------------------------------------------------
// implementation (header-only; include once per translation unit set via DMST_EMUL_IMPL)
// ----------------------------------------------------------------------------------
#ifdef DMST_EMUL_IMPL
#include <map>
#include <mutex>
namespace emul {
thread_local uint3_ t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local BlockCtx* t_block = nullptr;
thread_local int t_static_smem_key = 0;
static std::mutex g_smem_mu;
static std::map<int, std::vector<unsigned char>> g_static_smem;

void* static_smem(int key, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_smem_mu);
    auto& v = g_static_smem[key];
    if (v.size() < bytes + 64) v.assign(bytes + 64, 0);
    uintptr_t p = reinterpret_cast<uintptr_t>(v.data());
    p = (p + 15) & ~uintptr_t(15);
    return reinterpret_cast<void*>(p);
}
unsigned char* dynamic_smem() {
    uintptr_t p = reinterpret_cast<uintptr_t>(t_block->dyn_smem.data());
    p = (p + 127) & ~uintptr_t(127);
    return reinterpret_cast<unsigned char*>(p);
}
}  // namespace emul
#endif

namespace emul {
template <typename F>
void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
    int nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                BlockCtx ctx;
                ctx.nthreads = nthreads;
                ctx.bar = std::make_unique<std::barrier<>>(nthreads);
                int nwarps = (nthreads + 31) / 32;
                ctx.warps.resize(nwarps);
                for (int w = 0; w < nwarps; ++w) {
                    int cnt = std::min(32, nthreads - 32 * w);
                    ctx.warps[w].bar = std::make_unique<std::barrier<>>(cnt);
                }
                ctx.dyn_smem.assign(smem + 256, 0);
                std::vector<std::thread> ths;
                ths.reserve(nthreads);
                for (int t = 0; t < nthreads; ++t) {
                    ths.emplace_back([&, t]() {
                        t_threadIdx = {unsigned(t % block.x), unsigned((t / block.x) % block.y),
                                       unsigned(t / (block.x * block.y))};
                        t_blockIdx = {bx, by, bz};
                        t_blockDim = block;
                        t_gridDim = grid;
                        t_block = &ctx;
                        body();
                    });
                }
                for (auto& th : ths) th.join();
            }
}
}  // namespace emul
