// Shared definitions for the diffmst_b200 kernels (sm_100a).
//
// The same sources also compile with g++ -DDMST_EMULATE against tests/emul/cuda_emul.h,
// which is test infrastructure for debugging kernel logic on a machine without a GPU.
#pragma once

#ifdef DMST_EMULATE
#include "cuda_emul.h"
#define DMST_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emul::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define DMST_DYN_SMEM(name) unsigned char* name = emul::dynamic_smem()
#define DMST_DEVICE_BUILD 0
#else
#include <cuda_runtime.h>
#define DMST_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DMST_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define DMST_SHARED_ARRAY(type, name, count) __shared__ __align__(16) type name[count]
#define DMST_DEVICE_BUILD 1
#endif

#include <stdint.h>

namespace dmst {

constexpr int kNumSections = 6;   // low_shelf, band0..3, high_shelf (mst/modules.py:124-143)
constexpr int kMaxL = 32;         // samples per thread chunk upper bound
constexpr float kDbPerLog2 = 6.020599913279624f;     // 20*log10(2)
constexpr float kLog2Per20Db = 0.16609640474436813f;  // log2(10)/20
constexpr float kLn10Over20 = 0.11512925464970229f;   // ln(10)/20
constexpr float k20OverLn10 = 8.685889638065035f;     // 20/ln(10)
constexpr float kCompEps = 1e-8f;                      // dasp compressor eps (Appendix A)

// ---------------------------------------------------------------------------------
// Per-row (track or master bus) table, produced in float64 by the prepare kernel and
// stored as float32.  One per (row, chunk length L).
//
// A biquad section in transposed direct form II has the 2-vector state s and the
// transition s' = A s + B x with A = [[-a1, 1], [-a2, 0]].  With each thread owning L
// consecutive samples, P = A^L advances the state across one thread chunk:
//   P2[j] = P^(2^j), j = 0..8: j < 5 are the warp Kogge-Stone steps, P2[5] = P^32 advances
//   across a whole warp and P2[5..8] are the steps of the second-level scan over warps,
//   Ppow[l] = P^l    (warp carry-in -> lane carry-in).
// Matrices are row-major {m00, m01, m10, m11}.
// ---------------------------------------------------------------------------------
struct SectionTab {
    float b0, b1, b2, a1, a2, inv_b0, pad0, pad1;
    float P2[9][4];
    float pad2[4];
    float Ppow[32][4];
};

struct RowTab {
    SectionTab sec[kNumSections];
    // input gain (linear), output gain (linear; master only), pan gains (tracks only)
    float g_in, g_out, gL, gR;
    // compressor: y = x_delayed * 10^((g_s + makeup)/20)
    float alpha, beta;      // one-pole smoother g_s[n] = beta*g_c[n] + alpha*g_s[n-1]
    float thr_lo;           // threshold - knee/2
    float knee, inv_knee, inv_2knee;
    float slope;            // 1/ratio - 1
    float makeup;
    float inv_ratio2;       // 1/ratio^2 (backward)
    float pad[3];
    float a2pow[9];         // alpha^(L*2^j), j = 0..8
    float pad2[3];
    float a_lane[32];       // alpha^(L*l)
    float a_i[kMaxL];       // alpha^(i+1)
};

// Flags understood by the chain kernels
constexpr unsigned kChainGain = 1u, kChainEq = 2u, kChainComp = 4u, kChainOutGain = 8u;
constexpr unsigned kChainDebugNoWait = 64u;  // measurement only (DMST_DEBUG_NOWAIT=1): skip inter-tile waits, results invalid

// Gradient partial layout (per row, per tile), see console_bwd.cu
constexpr int kGradEq = 0;        // 30 values: section*5 + {b0,b1,b2,a1,a2}
constexpr int kGradAlpha = 30, kGradThr = 31, kGradRatio = 32, kGradKnee = 33, kGradMakeup = 34;
constexpr int kGradGin = 35, kGradGout = 36, kGradGL = 37, kGradGR = 38;
constexpr int kGradCount = 40;

// ---------------------------------------------------------------------------------
// Inter-CTA signalling (tile k waits for tile k-1 of the same row).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void st_release(int* p, int v) {
#ifdef DMST_EMULATE
    reinterpret_cast<std::atomic<int>*>(p)->store(v, std::memory_order_release);
#else
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ int ld_acquire(const int* p) {
#ifdef DMST_EMULATE
    return reinterpret_cast<const std::atomic<int>*>(p)->load(std::memory_order_acquire);
#else
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void wait_flag_ge(const int* p, int v, bool skip = false) {
    if (skip) return;
    while (ld_acquire(p) < v) {
#ifndef DMST_EMULATE
        __nanosleep(64);
#else
        __nanosleep(0);
#endif
    }
}

// ---------------------------------------------------------------------------------
// "Flag in data" mailboxes for the chained section states: a value and its validity tag travel
// in one aligned 8-byte word, so a consumer needs a single L2 round trip and the producer needs
// no fence.  Mailboxes are zeroed (cudaMemsetAsync) before each launch; tag == kMailValid marks
// a published value.
// ---------------------------------------------------------------------------------
struct __align__(8) Mail { float v; int tag; };
constexpr int kMailValid = 1;
__device__ __forceinline__ void mail_put(Mail* p, float v) {
#ifdef DMST_EMULATE
    uint64_t w; Mail m{v, kMailValid}; memcpy(&w, &m, 8);
    reinterpret_cast<std::atomic<uint64_t>*>(p)->store(w, std::memory_order_release);
#else
    asm volatile("st.volatile.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_int(v)), "r"(kMailValid) : "memory");
#endif
}
__device__ __forceinline__ bool mail_try(const Mail* p, float& v) {
#ifdef DMST_EMULATE
    uint64_t w = reinterpret_cast<const std::atomic<uint64_t>*>(p)->load(std::memory_order_acquire);
    Mail m; memcpy(&m, &w, 8);
    v = m.v; return m.tag == kMailValid;
#else
    int a, b;
    asm volatile("ld.volatile.global.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
    v = __int_as_float(a);
    return b == kMailValid;
#endif
}
__device__ __forceinline__ float mail_wait(const Mail* p, bool skip = false) {
    float v;
    while (!mail_try(p, v)) {
        if (skip) break;
#ifndef DMST_EMULATE
#ifndef DMST_MAIL_SLEEP
#define DMST_MAIL_SLEEP 32
#endif
        if (DMST_MAIL_SLEEP > 0) __nanosleep(DMST_MAIL_SLEEP);
#else
        __nanosleep(0);
#endif
    }
    return v;
}

// padded shared-memory index for 128-bit accesses: 4 floats of padding per 32, so that lane l reading
// float4 #j of its own L-sample chunk (L = 16 or 32), or thread q moving float4 #q of a tile, is
// conflict-free, and every float4 stays 16-byte aligned (targets of cp.async)
__host__ __device__ __forceinline__ int pidx4(int i) { return i + ((i >> 5) << 2); }

// ---------------------------------------------------------------------------------
// Asynchronous global -> shared copies (LDGSTS) used to prefetch the next tile of a persistent CTA
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
#ifdef DMST_EMULATE
    memcpy(smem_dst, gsrc, 16);
#else
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
#ifdef DMST_EMULATE
    memcpy(smem_dst, gsrc, 4);
#else
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef DMST_EMULATE
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef DMST_EMULATE
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
// counter += v with release semantics (cumulative over a preceding __syncthreads, like st_release)
__device__ __forceinline__ void red_release_add(int* p, int v) {
#ifdef DMST_EMULATE
    reinterpret_cast<std::atomic<int>*>(p)->fetch_add(v, std::memory_order_release);
#else
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// 2^x by the MUFU unit (ex2.approx: relative error below 2^-22, flushes subnormal results)
__device__ __forceinline__ float fast_exp2(float x) {
#ifdef DMST_EMULATE
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dmst
