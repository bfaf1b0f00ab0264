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

// Everything of a row that is not an EQ section: gains, pan, compressor constants and smoother scan tables.
// (First base of RowTab, so that a kernel that needs no section tables can fetch just sizeof(CompTab) bytes.)
struct CompTab {
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
static_assert(sizeof(CompTab) % 16 == 0, "CompTab is copied with 16-byte cp.async");

struct RowTab : CompTab {
    SectionTab sec[kNumSections];
};

// ---------------------------------------------------------------------------------
// Tables of the track backward kernel (console_bwd2.cuh): per EQ section two reverse-time all-pole
// recursions driven by the gradient u at the EQ output, g = A^-T u (type 0) and h = (B/b0)^-T u (type 1),
// run in DELTA form: state (s, v) with v[n] = s[n] - s[n+1],
//     v[n] = u[n] - c0 s[n+1] + c2 v[n+1],   s[n] = s[n+1] + v[n],      c0 = 1 + c1 + c2
// for the polynomial 1 + c1 z^-1 + c2 z^-2.  For roots near z = 1 (low shelf / low band) c0 ~ w0^2 is tiny: it
// is computed in float64 and stored as its own float32, so the pole position keeps full relative precision
// (direct-form a1 ~ -2, a2 ~ 1 in float32 would lose it), the rounding noise of the recursion is amplified by
// ~1/w0 instead of ~1/w0^2, and the differences of the (smooth, huge) filtered signal that the coefficient
// gradients need are state variables instead of float32 cancellations.
// M = [[1-c0, c2], [-c0, c2]] advances (s, v) by one sample with zero input; P = M^32 (chunk of 32 samples).
// (built from the float32 c0 and c2 the kernel uses: tables and recursion agree exactly.)
// ---------------------------------------------------------------------------------
constexpr int kBwd2Chunk = 32;     // samples per thread in the EQ-gradient phase
// The two recursions of a section run in lock step as the halves of packed float2 operations
// (fma.rn.f32x2 / add.rn.f32x2: one issue slot for both): .x = g (poles), .y = h (zeros), both in the form
//     t = nc0 s + u,   v' = k2 v + t,   s' = s + v',      nc0 = -c0, k2 = c2
// (the second difference v' - v the b2 gradient needs is never formed: its correlation with e is summed by parts into
// a correlation of v with e[n] - e[n-1], console_bwd2.cuh).
struct PairTab {
    float2 nc0;          // -c0
    float2 nd2;          // -(1 - c2): the second difference as nd2 v + t (kept in the table; the kernel no longer forms it)
    float2 scale;        // factor of each recursion's sums in the tile partials: (1, 1/b0)
    float2 k2;           // c2
    float2 P2[6][4];     // P^(2^j), j = 0..4 warp scan, P2[5] = P^32 across a warp; P = M^32, M row-major
    float2 Ppow[32][4];  // P^l
};
static_assert(sizeof(PairTab) == 312 * 4 && sizeof(PairTab) % 16 == 0, "PairTab layout");
constexpr int kNumRec = 2 * kNumSections;
struct EqBwdTab { PairTab sec[kNumSections]; };

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

// Programmatic dependent launch: lets the next kernel of the stream (launched with launch_dependent, console_host.cuh)
// start while this grid is still running.  The dependent must not rely on this grid's completion: it synchronises
// with it through release / acquire flags.
__device__ __forceinline__ void griddep_launch_dependents() {
#ifndef DMST_EMULATE
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
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

// packed pairs of FP32 operations (sm_100: FFMA2 / FADD2, one issue slot for two IEEE operations)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
#ifdef DMST_EMULATE
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#else
    return __ffma2_rn(a, b, c);
#endif
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
#ifdef DMST_EMULATE
    return make_float2(a.x + b.x, a.y + b.y);
#else
    return __fadd2_rn(a, b);
#endif
}
__device__ __forceinline__ float2 shfl_down2(float2 v, int d) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
// (s1, s2) += M (t1, t2) for a pair of 2x2 matrices m = {m00, m01, m10, m11} (each a pair)
__device__ __forceinline__ void mat2_apply_acc2(const float2* m, float2 t1, float2 t2, float2& s1, float2& s2) {
    s1 = fma2(m[0], t1, fma2(m[1], t2, s1));
    s2 = fma2(m[2], t1, fma2(m[3], t2, s2));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dmst
