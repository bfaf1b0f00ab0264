// Argument blocks and device helpers shared by the forward and backward chain kernels.
//
// A "row" is one mono track (NCH = 1, the per-track chain gain -> EQ -> compressor of
// mst/modules.py:230-251) or one stereo master bus (NCH = 2, mst/modules.py:286-312, both
// channels share coefficients and the compressor side-chain is their sum).  Time is cut
// into tiles of NT*L samples; a (row, tile) is one work item of a persistent CTA; thread t owns
// samples [t*L, t*L+L) of the tile in registers.  Linear recurrences are solved exactly across
// threads by a scan whose combine operator is a constant matrix power (tables in RowTab),
// and across tiles by chaining work items through global memory (per-section wavefront).
#pragma once
#include "common.cuh"

namespace dmst {

constexpr int kStateStride = 32;   // floats per (row, tile): [sec][ch][2] (24) + smoother (1)
constexpr int kStateSmooth = 24;
constexpr int kTail2Stride = 32;   // floats per (row, tile): [stage 0..6][ch][2]
constexpr int kFlagSmooth = 7;
constexpr int kBwdChunk = 16;  // thread chunk of the track backward kernel = spacing of forward's state checkpoints

struct ChainArgs {
    // ---- geometry ----
    int nrows;        // B*N (tracks) or B (master)
    int N;            // tracks per batch item
    int T;            // samples per row
    int Tp;           // padded row length of internal scratch rows (multiple of 4)
    int ntiles;
    unsigned flags;   // kChain*
    int lookahead;
    int want_mixed;
    int want_grad_src;
    // ---- forward source / sinks ----
    const float* src;            // tracks: caller's (B,N,T); master: y scratch (B*N, Tp)
    long long src_batch_stride;  // elements
    long long src_row_stride;
    int src_vec_ok;
    int user_vec_ok;             // user-owned outputs 16B-aligned with T % 4 == 0
    const RowTab* tab;           // [nrows]
    const RowTab* track_tab;     // master only: [B*N], pan gains
    float* y;                    // tracks: (B*N, Tp) post-compressor mono scratch
    float* mixed;                // tracks: caller's (B,2,N,T) or null
    float* mix;                  // master: caller's (B,2,T)
    float* bus_pre;              // master: (B*2, Tp) bus before the master chain
    // ---- chain workspace (kept for backward) ----
    int* ticket;
    int* flag;                   // [nrows*ntiles]
    Mail* state;                 // [nrows*ntiles*kStateStride] end-of-tile states (mailboxes, see common.cuh)
    float* tail2;                // [nrows*ntiles*kTail2Stride] last two samples of each stage
    float* etail;                // [nrows*ntiles*NCH*lookahead] last `lookahead` EQ outputs
    // optional checkpoints that let backward skip the forward EQ recompute (tracks):
    float* esave;                // (nrows, Tp) EQ output, or null
    float* ssave;                // [nrows*ntiles][6*NCH*2][TILE/kBwdChunk] section states every kBwdChunk samples
    float* gmid;                 // tracks: smoother state every 2^gmid_shift samples inside a tile, [nrows*ntiles][(TILE >> gmid_shift) - 1]
    int gmid_shift;              //   (the track backward kernel's tile boundaries), or null
    // ---- backward only ----
    const float* gout;           // tracks: dbus (B*2, Tp); master: caller's grad_mix (B,2,T)
    const float* gmixed;         // tracks: caller's grad of mixed_tracks (B,2,N,T) or null
    float* gsrc;                 // tracks: caller's grad_tracks (B,N,T) or null; master: dbus (B*2,Tp)
    float* partial;              // [nrows*ntiles*kGradCount]
    int* bflag;                  // [nrows*ntiles] reverse-chain flags
    Mail* bstate;                // [nrows*ntiles*kStateStride] reverse states at tile start (mailboxes)
    float* dhead;                // [nrows*ntiles*NCH*lookahead] first `lookahead` of dy*G
};

__device__ __forceinline__ void mat2_apply_acc(const float* m, float t1, float t2, float& s1, float& s2) {
    s1 = fmaf(m[0], t1, fmaf(m[1], t2, s1));
    s2 = fmaf(m[2], t1, fmaf(m[3], t2, s2));
}
// transposed matrix (reverse-time all-pole recursion uses A^T)
__device__ __forceinline__ void mat2T_apply_acc(const float* m, float t1, float t2, float& s1, float& s2) {
    s1 = fmaf(m[0], t1, fmaf(m[2], t2, s1));
    s2 = fmaf(m[1], t1, fmaf(m[3], t2, s2));
}

// Store L consecutive floats starting at p[0] (global), first `valid` elements only.
template <int L>
__device__ __forceinline__ void store_chunk(float* p, int valid, bool vec_ok, const float (&v)[L]) {
    if (vec_ok && valid >= L) {
        float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
        for (int i = 0; i < L / 4; ++i) p4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < L; ++i)
            if (i < valid) p[i] = v[i];
    }
}


// ---------------------------------------------------------------------------------
// Second-level scans over the NW warp aggregates of a CTA, done redundantly by every warp
// with shuffles (no divergent loops, one barrier per recurrence).
//
// Forward: sW[u] = end state of warp u with zero warp carry-in; (sin1, sin2) = tile carry-in.
// Returns the carry-in (c1, c2) of the calling warp and the end state of the tile.
// ---------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void cross_warp_fwd(const float* sW, int stride, const float (*P2)[4], float sin1,
                                               float sin2, int lane, int warp, float& c1, float& c2,
                                               float& end1, float& end2) {
    float w1 = 0.0f, w2 = 0.0f;
    if (lane < NW) { w1 = sW[lane * stride]; w2 = sW[lane * stride + 1]; }
    if (lane == 0) mat2_apply_acc(P2[5], sin1, sin2, w1, w2);  // fold the tile carry-in into warp 0
#pragma unroll
    for (int j = 0; (1 << j) < NW; ++j) {
        const float t1 = __shfl_up_sync(0xffffffffu, w1, 1 << j);
        const float t2 = __shfl_up_sync(0xffffffffu, w2, 1 << j);
        if (lane >= (1 << j)) mat2_apply_acc(P2[5 + j], t1, t2, w1, w2);
    }
    end1 = __shfl_sync(0xffffffffu, w1, NW - 1);
    end2 = __shfl_sync(0xffffffffu, w2, NW - 1);
    const int src = warp > 0 ? warp - 1 : 0;
    c1 = __shfl_sync(0xffffffffu, w1, src);
    c2 = __shfl_sync(0xffffffffu, w2, src);
    if (warp == 0) { c1 = sin1; c2 = sin2; }
}
// Reverse: sW[u] = state at the START of warp u with zero right carry; (sin1, sin2) = state at
// the start of the successor tile.  Transposed matrices (reverse all-pole recursion).
template <int NW>
__device__ __forceinline__ void cross_warp_rev(const float* sW, int stride, const float (*P2)[4], float sin1,
                                               float sin2, int lane, int warp, float& c1, float& c2,
                                               float& start1, float& start2) {
    float w1 = 0.0f, w2 = 0.0f;
    if (lane < NW) { w1 = sW[lane * stride]; w2 = sW[lane * stride + 1]; }
    if (lane == NW - 1) mat2T_apply_acc(P2[5], sin1, sin2, w1, w2);
#pragma unroll
    for (int j = 0; (1 << j) < NW; ++j) {
        const float t1 = __shfl_down_sync(0xffffffffu, w1, 1 << j);
        const float t2 = __shfl_down_sync(0xffffffffu, w2, 1 << j);
        if (lane + (1 << j) < NW) mat2T_apply_acc(P2[5 + j], t1, t2, w1, w2);
    }
    start1 = __shfl_sync(0xffffffffu, w1, 0);
    start2 = __shfl_sync(0xffffffffu, w2, 0);
    const int src = warp < NW - 1 ? warp + 1 : NW - 1;
    c1 = __shfl_sync(0xffffffffu, w1, src);
    c2 = __shfl_sync(0xffffffffu, w2, src);
    if (warp == NW - 1) { c1 = sin1; c2 = sin2; }
}
// scalar (one-pole) versions; a2pow[j] = alpha^(L 2^j)
template <int NW>
__device__ __forceinline__ void cross_warp_fwd1(const float* sW, int stride, const float* a2pow, float sin,
                                                int lane, int warp, float& c, float& end) {
    float w = (lane < NW) ? sW[lane * stride] : 0.0f;
    if (lane == 0) w = fmaf(a2pow[5], sin, w);
#pragma unroll
    for (int j = 0; (1 << j) < NW; ++j) {
        const float t = __shfl_up_sync(0xffffffffu, w, 1 << j);
        if (lane >= (1 << j)) w = fmaf(a2pow[5 + j], t, w);
    }
    end = __shfl_sync(0xffffffffu, w, NW - 1);
    c = __shfl_sync(0xffffffffu, w, warp > 0 ? warp - 1 : 0);
    if (warp == 0) c = sin;
}
template <int NW>
__device__ __forceinline__ void cross_warp_rev1(const float* sW, int stride, const float* a2pow, float sin,
                                                int lane, int warp, float& c, float& start) {
    float w = (lane < NW) ? sW[lane * stride] : 0.0f;
    if (lane == NW - 1) w = fmaf(a2pow[5], sin, w);
#pragma unroll
    for (int j = 0; (1 << j) < NW; ++j) {
        const float t = __shfl_down_sync(0xffffffffu, w, 1 << j);
        if (lane + (1 << j) < NW) w = fmaf(a2pow[5 + j], t, w);
    }
    start = __shfl_sync(0xffffffffu, w, 0);
    c = __shfl_sync(0xffffffffu, w, warp < NW - 1 ? warp + 1 : NW - 1);
    if (warp == NW - 1) c = sin;
}

// 4 consecutive floats from global memory with bounds (zero beyond `valid`)
__device__ __forceinline__ float4 load4(const float* p, int valid, bool vec_ok) {
    if (vec_ok && valid >= 4) return __ldg(reinterpret_cast<const float4*>(p));
    float4 r;
    r.x = valid > 0 ? __ldg(p) : 0.0f;
    r.y = valid > 1 ? __ldg(p + 1) : 0.0f;
    r.z = valid > 2 ? __ldg(p + 2) : 0.0f;
    r.w = valid > 3 ? __ldg(p + 3) : 0.0f;
    return r;
}

// Static-curve gain computer of the dasp compressor, branch-free:
// t = x_db - (thr - knee/2); g_c = slope * (clamp(t,0,W)^2/(2W) + max(t-W,0)).
__device__ __forceinline__ float gain_computer(float side, const CompTab& tb, float& tc, float& lin) {
    float d = kDbPerLog2 * __log2f(fmaxf(fabsf(side), kCompEps));
    float t = d - tb.thr_lo;
    tc = fminf(fmaxf(t, 0.0f), tb.knee);
    lin = fmaxf(t - tb.knee, 0.0f);
    return tb.slope * fmaf(tc * tc, tb.inv_2knee, lin);
}

}  // namespace dmst
