// Backward pass of one chain kind of the mix console (adjoint of console_fwd.cuh): a persistent
// kernel whose CTAs claim (row, time tile) work items through an atomic ticket in REVERSE time
// order (so the successor tile of any running item is running or finished: no deadlock), and
// prefetch the next item's inputs (row table, EQ-output checkpoint, look-ahead halo, upstream
// gradient) into shared memory with cp.async while the current one is processed.  Launched once
// for the master bus (NCH = 2; produces the bus gradient) and once for the tracks (NCH = 1).
//
// Per tile the kernel
//   1. recomputes the forward chain from the tile carry-in states the forward pass left in
//      the workspace (no forward chaining needed),
//   2. runs the adjoint of the compressor (reverse one-pole recursion for the smoothed
//      gain, piecewise derivatives of the static curve, look-ahead halo taken from the
//      successor tile),
//   3. runs the adjoint of each biquad section in reverse order: reverse-time all-pole
//      recursion g = (1/A)^T u solved with the same matrix-power scan (transposed
//      matrices), coefficient gradients db_j = sum g[n] x[n-j], da_j = -sum g[n] y[n-j]
//      (accumulated in a differenced basis, see below), input gradient B^T g.  The section
//      input x is recovered from its output with the inverse recursion that shares the
//      section's state trajectory, so only two signals live in registers at a time,
//   4. writes per-tile partial sums; console_prepare.cuh's epilogue chains them through the
//      parameter Jacobian.
// Gradient definitions follow the float64 autograd of the oracle (tests/golden).
#pragma once
#include "chain.cuh"
#include "console_fwd.cuh"   // lds_chunk / sts_chunk

namespace dmst {

constexpr int kBFlagComp = 1;  // reverse smoother state + dhead halo published (section states use mailboxes)
constexpr int kBFlagDone = 2;  // master: the tile's bus gradient is stored (the track kernel, launched as a programmatic
                               // dependent, consumes it while later master tiles are still in flight)

struct BwdArgs {
    ChainArgs a;
    int total;        // rows * tiles
    int area;         // floats of the per-tile buffer area in dynamic shared memory (sE follows it)
    int* ticket;
};

// `count` (multiple of 4) consecutive floats of a global row -> shared memory (pidx4 layout), zero
// beyond `valid`; 16-byte copies when the source allows it
template <int NT>
__device__ __forceinline__ void cp_tile(float* dst, const float* src, int count, int valid, bool vec_ok, int tid) {
    static_assert(NT % 8 == 0, "constant shared-memory stride per iteration");
    if (vec_ok && valid >= count) {
        // interior tile: thread q moves float4 #q, #q+NT, ...; pidx4(4(q+NT)) - pidx4(4q) is constant
        float* d = dst + pidx4(tid << 2);
        const float* p = src + (tid << 2);
        for (int q = tid; q < (count >> 2); q += NT) {
            cp_async16(d, p);
            d += 4 * NT + (NT >> 1);
            p += 4 * NT;
        }
        return;
    }
    for (int q = tid; q < (count >> 2); q += NT) {
        const int idx = q << 2;
        float* d = dst + pidx4(idx);
        if (vec_ok && valid - idx >= 4) {
            cp_async16(d, src + idx);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (idx + e < valid) cp_async4(d + e, src + idx + e);
                else d[e] = 0.0f;
            }
        }
    }
}

template <int NT, int NCH>
struct BwdShared {
    float W[8 * (NT / 32) * NCH * 2];     // warp aggregates, slot 7 = reverse smoother
    float nb[3 * (NT / 32) * NCH * 2];    // last two samples of each warp: [section parity 0/1, chain output]
    float pre[kStateStride];              // successor's reverse states (prefetched)
    float fst[kStateStride];              // predecessor's forward end states (saved by forward)
    unsigned premask[2];                  // [0] published mailboxes at tile start, [1] successor flag
    float part[(NT / 32) * kGradCount];
    int next;                             // next ticket of this CTA
};

// Sum four values over the warp with 6 shuffles: the two halves (then quarters) of the warp exchange
// the values they do not keep.  Afterwards lanes with (lane & 7) == 0 hold the totals: lane 0 -> a0,
// lane 16 -> a1, lane 8 -> a2, lane 24 -> a3 (fixed tree => deterministic).
__device__ __forceinline__ float warp_sum4(float a0, float a1, float a2, float a3, int lane) {
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
    float p = h16 ? a1 : a0, ps = h16 ? a0 : a1;
    float q = h16 ? a3 : a2, qs = h16 ? a2 : a3;
    p += __shfl_xor_sync(0xffffffffu, ps, 16);
    q += __shfl_xor_sync(0xffffffffu, qs, 16);
    float r = h8 ? q : p, rs = h8 ? p : q;
    r += __shfl_xor_sync(0xffffffffu, rs, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}
__device__ __forceinline__ int warp_sum4_slot(int lane) { return ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1); }

// Start the asynchronous copies of the inputs of work item `ticket` into the (free) buffer area.
template <int NCH, int L, int NT, bool MASTER>
__device__ __forceinline__ void bwd_prefetch(const BwdArgs& f, int ticket, float* area, float* tabbuf, int tid) {
    constexpr int TILE = NT * L;
    if (ticket < f.total) {
        const ChainArgs& a = f.a;
        const int tile = a.ntiles - 1 - ticket / a.nrows;
        const int row = ticket % a.nrows;
        const float* tsrc = reinterpret_cast<const float*>(a.tab + row);
        for (int i = tid; i < int(sizeof(RowTab) / 16); i += NT) cp_async16(tabbuf + 4 * i, tsrc + 4 * i);
        const int LA = a.lookahead;
        const int es = pidx4(LA + TILE), gs = pidx4(TILE);
        const int tbase = tile * TILE;
        const bool saved = (a.esave != nullptr) && (a.flags & kChainEq);
        const long long rt = (long long)row * a.ntiles + tile;
        const int bb = MASTER ? row : row / a.N;
        float* gbuf = area + 2 * NCH * es;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            float* e = area + c * es;
            // chain input of the tile (EQ-output checkpoint when forward left one)
            if (saved || MASTER) {
                const float* src = (saved ? a.esave : a.src) + (long long)(row * NCH + c) * a.Tp + tbase;
                cp_tile<NT>(e + pidx4(LA), src, TILE, a.Tp - tbase, true, tid);
            } else {
                const int n = row - bb * a.N;
                const float* src = a.src + (long long)bb * a.src_batch_stride + (long long)n * a.src_row_stride + tbase;
                cp_tile<NT>(e + pidx4(LA), src, TILE, a.T - tbase, a.src_vec_ok != 0, tid);
            }
            // look-ahead halo: predecessor's last LA EQ outputs (zeros before the start of the signal)
            if (a.flags & kChainComp)
                cp_tile<NT>(e, a.etail + (rt - 1) * NCH * LA + c * LA, LA, tile > 0 ? LA : 0, true, tid);
        }
        // upstream gradient (zero beyond the end of the signal): master: caller's grad_mix (B,2,T);
        // tracks: bus gradient (B*2,Tp) written by the master launch
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const long long len = MASTER ? a.T : a.Tp;
            cp_tile<NT>(gbuf + c * gs, a.gout + (long long)(bb * 2 + c) * len + tbase, TILE, a.T - tbase,
                        MASTER ? (a.user_vec_ok != 0) : true, tid);
        }
    }
    cp_async_commit();
}

// One (row, tile) of a chain.  Returns the next ticket of this CTA (claimed on the way, inputs prefetched).
template <int NCH, int L, int NT, bool MASTER>
__device__ __forceinline__ int bwd_tile(const BwdArgs& f, const int row, const int tile, float* area,
                                        const RowTab& tb, float* tab_next, BwdShared<NT, NCH>& sh) {
    constexpr int NW = NT / 32;
    constexpr int TILE = NT * L;
    static_assert(L % 4 == 0 && L <= kMaxL && 32 % L == 0, "chunk length");
    const ChainArgs& a = f.a;
    float* s_W = sh.W;
    float* s_nb = sh.nb;
    float* s_pre = sh.pre;
    float* s_fst = sh.fst;
    unsigned* s_premask = sh.premask;
    float* s_part = sh.part;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int LA = a.lookahead;
    const int buf_stride = pidx4(LA + TILE);
    const int gbuf_stride = pidx4(TILE);
    // LA % 32 == 0 and L | 32 (checked on the host): pidx4(LA + tid*L + i) = pLA + pb + i
    const int pLA = pidx4(LA), pb = pidx4(tid * L);
    float* ebuf = area;                                         // [NCH][buf_stride]  e delay line
    float* dbuf = ebuf + NCH * buf_stride;                      // [NCH][buf_stride]  dy*G, future halo
    float* gbuf = dbuf + NCH * buf_stride;                      // [2][gbuf_stride]   upstream gradient
    float* sE = area + f.area;                                  // [6*NCH*2][NT] forward lane carry-ins (no checkpoints)

    // Claim the next work item now; the ticket stays in a register until it is handed to the other
    // threads at the barrier that frees the buffer area (never waits for the atomic's latency).
    int claimed = 0;
    if (tid == 0) claimed = atomicAdd(f.ticket, 1);
    for (int i = tid; i < NW * kGradCount; i += NT) s_part[i] = 0.0f;
    const int t0 = tile * TILE + tid * L;
    const bool has_pred = tile > 0, has_succ = tile < a.ntiles - 1;
    const long long rt = (long long)row * a.ntiles + tile;
    const Mail* state_in = a.state + (rt - 1) * kStateStride;  // forward carry-in
    const float* tail_in = a.tail2 + (rt - 1) * kTail2Stride;
    Mail* bstate_out = a.bstate + rt * kStateStride;
    const Mail* bstate_in = a.bstate + (rt + 1) * kStateStride;
    int* my_flag = a.bflag + rt;
    const int* succ_flag = my_flag + 1;
    const bool nowait = (a.flags & kChainDebugNoWait) != 0;
    if (warp == 0) {
        float pv = 0.0f;
        const bool ok = has_succ ? mail_try(bstate_in + lane, pv) : true;
        s_pre[lane] = pv;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) { s_premask[0] = m; s_premask[1] = has_succ ? (unsigned)ld_acquire(succ_flag) : 99u; }
    } else if (warp == 1) {
        s_fst[lane] = has_pred ? state_in[lane].v : 0.0f;
    }
    const int bb = MASTER ? row : row / a.N;
    const int nn = MASTER ? 0 : row - bb * a.N;
    const bool uvec = a.user_vec_ok != 0;

    // With checkpoints from forward (EQ output + section states every kBwdChunk samples) the
    // forward EQ recompute below is skipped.
    const bool saved = (a.esave != nullptr) && (a.flags & kChainEq);
    static_assert(MASTER || L == kBwdChunk, "state checkpoints are spaced by the backward chunk length");  // (master: forward is built with CHK = this L)
    const float* ssave = a.ssave ? a.ssave + rt * (kNumSections * NCH * 2) * NT : nullptr;
    (void)uvec;
    float v[NCH][L];
#pragma unroll
    for (int c = 0; c < NCH; ++c) lds_chunk<L>(ebuf + c * buf_stride + pLA + pb, v[c]);  // prefetched; landed before the caller's barrier
    if (!saved && (a.flags & kChainGain)) {
        const float g = tb.g_in;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < L; ++i) v[c][i] *= g;
    }
    __syncthreads();  // prefetched states and zeroed partial sums visible

    // ---------------- forward recompute of the EQ cascade ----------------
    if ((a.flags & kChainEq) && !saved) {
#pragma unroll 1
        for (int k = 0; k < kNumSections; ++k) {
            const SectionTab& st = tb.sec[k];
            const float b0 = st.b0, b1 = st.b1, b2 = st.b2, na1 = -st.a1, na2 = -st.a2;
            float s1[NCH], s2[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float z1 = 0.0f, z2 = 0.0f;
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    const float x = v[c][i];
                    const float yv = fmaf(b0, x, z1);
                    z1 = fmaf(b1, x, fmaf(na1, yv, z2));
                    z2 = fmaf(b2, x, na2 * yv);
                    v[c][i] = yv;
                }
                s1[c] = z1; s2[c] = z2;
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const float t1 = __shfl_up_sync(0xffffffffu, s1[c], 1 << j);
                    const float t2 = __shfl_up_sync(0xffffffffu, s2[c], 1 << j);
                    if (lane >= (1 << j)) mat2_apply_acc(st.P2[j], t1, t2, s1[c], s2[c]);
                }
            }
            float e1[NCH], e2[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                e1[c] = __shfl_up_sync(0xffffffffu, s1[c], 1);
                e2[c] = __shfl_up_sync(0xffffffffu, s2[c], 1);
                if (lane == 0) { e1[c] = 0.0f; e2[c] = 0.0f; }
                if (lane == 31) {
                    s_W[((k * NW + warp) * NCH + c) * 2 + 0] = s1[c];
                    s_W[((k * NW + warp) * NCH + c) * 2 + 1] = s2[c];
                }
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float i1 = s_fst[(k * NCH + c) * 2 + 0], i2 = s_fst[(k * NCH + c) * 2 + 1];
                float c1, c2, n1, n2;
                cross_warp_fwd<NW>(s_W + (k * NW * NCH + c) * 2, NCH * 2, st.P2, i1, i2, lane, warp, c1, c2, n1, n2);
                mat2_apply_acc(st.Ppow[lane], c1, c2, e1[c], e2[c]);
                sE[((k * NCH + c) * 2 + 0) * NT + tid] = e1[c];
                sE[((k * NCH + c) * 2 + 1) * NT + tid] = e2[c];
                float h1 = e1[c], h2 = e2[c];
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    const float t = h1;
                    v[c][i] += t;
                    h1 = fmaf(na1, t, h2);
                    h2 = na2 * t;
                }
            }
        }
    }

    float u[NCH][L];
    float acc_gout = 0.0f, acc_gl = 0.0f, acc_gr = 0.0f;
    float acc_alpha = 0.0f, acc_thr = 0.0f, acc_ratio = 0.0f, acc_knee = 0.0f, acc_makeup = 0.0f;

    // Upstream gradient of samples [i0, i0+4) of the chunk w.r.t. the chain output (tracks: gL*dbusL +
    // gR*dbusR [+ grad of mixed_tracks]; master: g_out * dmix), given the chain output `o` of those
    // samples; accumulates the gain gradients on the way.  (The prefetch zero-fills beyond the signal.)
    auto upstream4 = [&](int i0, const float (&o)[NCH][4], float (&dc)[NCH][4]) {
        if constexpr (MASTER) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float4 g4 = *reinterpret_cast<const float4*>(gbuf + c * gbuf_stride + pb + i0);
                const float gq[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float g = gq[j];
                    if (a.flags & kChainOutGain) { acc_gout = fmaf(g, o[c][j] * tb.g_out, acc_gout); g *= tb.g_out; }
                    dc[c][j] = g;
                }
            }
        } else {
            const float4 l4 = *reinterpret_cast<const float4*>(gbuf + pb + i0);
            const float4 r4 = *reinterpret_cast<const float4*>(gbuf + gbuf_stride + pb + i0);
            float bl[4] = {l4.x, l4.y, l4.z, l4.w}, br[4] = {r4.x, r4.y, r4.z, r4.w};
            if (a.gmixed) {
                const float4 ml = load4(a.gmixed + ((long long)(bb * 2 + 0) * a.N + nn) * a.T + t0 + i0, a.T - t0 - i0, uvec);
                const float4 mr = load4(a.gmixed + ((long long)(bb * 2 + 1) * a.N + nn) * a.T + t0 + i0, a.T - t0 - i0, uvec);
                bl[0] += ml.x; bl[1] += ml.y; bl[2] += ml.z; bl[3] += ml.w;
                br[0] += mr.x; br[1] += mr.y; br[2] += mr.z; br[3] += mr.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc_gl = fmaf(bl[j], o[0][j], acc_gl);
                acc_gr = fmaf(br[j], o[0][j], acc_gr);
                dc[0][j] = fmaf(tb.gL, bl[j], tb.gR * br[j]);
            }
        }
    };

    if (a.flags & kChainComp) {
        // ---- forward recompute: smoothed gain.  The look-ahead halo was prefetched into the delay line;
        // so was the EQ output when forward left a checkpoint, otherwise it was just recomputed ----
        if (!saved) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) sts_chunk<L>(ebuf + c * buf_stride + pLA + pb, v[c]);
        }
        float gs[L];
        float gz = 0.0f;
        const float alpha = tb.alpha, beta = tb.beta;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            float side = v[0][i];
            if (NCH > 1) side += v[NCH - 1][i];
            float tc, lin;
            const float gc = gain_computer(side, tb, tc, lin);
            gz = fmaf(alpha, gz, beta * gc);
            gs[i] = gz;
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float t = __shfl_up_sync(0xffffffffu, gz, 1 << j);
            if (lane >= (1 << j)) gz = fmaf(tb.a2pow[j], t, gz);
        }
        float ex = __shfl_up_sync(0xffffffffu, gz, 1);
        if (lane == 0) ex = 0.0f;
        if (lane == 31) s_W[(6 * NW + warp) * NCH * 2] = gz;
        __syncthreads();  // ebuf + smoother aggregates visible
        float cw, gend;
        cross_warp_fwd1<NW>(s_W + 6 * NW * NCH * 2, NCH * 2, tb.a2pow, s_fst[kStateSmooth], lane, warp, cw, gend);
        const float gcarry = fmaf(tb.a_lane[lane], cw, ex);  // g_s just before this chunk
#pragma unroll
        for (int i = 0; i < L; ++i) gs[i] = fmaf(tb.a_i[i], gcarry, gs[i]);

        // ---- adjoint: output -> (delayed signal path, gain path) ----
        float q[L];
        float* dhead_out = a.dhead + rt * NCH * LA;
#pragma unroll
        for (int i0 = 0; i0 < L; i0 += 4) {
            float G[4], o[NCH][4], dc[NCH][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) G[j] = fast_exp2(kLog2Per20Db * (gs[i0 + j] + tb.makeup));
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float4 e4 = *reinterpret_cast<const float4*>(ebuf + c * buf_stride + pb + i0);  // x[n - LA]
                o[c][0] = e4.x * G[0]; o[c][1] = e4.y * G[1]; o[c][2] = e4.z * G[2]; o[c][3] = e4.w * G[3];
            }
            upstream4(i0, o, dc);
            float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float dyG[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { r[j] = fmaf(dc[c][j], o[c][j], r[j]); dyG[j] = dc[c][j] * G[j]; }
                const float4 d4 = make_float4(dyG[0], dyG[1], dyG[2], dyG[3]);
                *reinterpret_cast<float4*>(dbuf + c * buf_stride + pb + i0) = d4;
                if (tid * L < LA) *reinterpret_cast<float4*>(dhead_out + c * LA + tid * L + i0) = d4;  // whole chunk: L | LA
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { q[i0 + j] = r[j] * kLn10Over20; acc_makeup += q[i0 + j]; }
        }
        // ---- reverse one-pole: p[n] = q[n] + alpha p[n+1] ----
        float pz = 0.0f;
#pragma unroll
        for (int i = L - 1; i >= 0; --i) { pz = fmaf(alpha, pz, q[i]); q[i] = pz; }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float t = __shfl_down_sync(0xffffffffu, pz, 1 << j);
            if (lane + (1 << j) < 32) pz = fmaf(tb.a2pow[j], t, pz);
        }
        float px = __shfl_down_sync(0xffffffffu, pz, 1);
        if (lane == 31) px = 0.0f;
        if (lane == 0) s_W[(7 * NW + warp) * NCH * 2] = pz;
        // successor already published its smoother state and halo (the common case): fetch the
        // halo now and save a barrier; otherwise wait for it and fetch after the barrier
        const bool halo_early = s_premask[1] >= (unsigned)kBFlagComp;
        auto fetch_halo = [&]() {
            const float* dhead_in = a.dhead + (rt + 1) * NCH * LA;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                for (int j = 4 * tid; j < LA; j += 4 * NT) {
                    const float4 h = has_succ ? __ldcg(reinterpret_cast<const float4*>(dhead_in + c * LA + j))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(dbuf + c * buf_stride + pidx4(TILE + j)) = h;
                }
        };
        if (halo_early) fetch_halo();
        else if (tid == 0) wait_flag_ge(succ_flag, kBFlagComp, nowait);
        if (tid == 0 && has_succ && !((s_premask[0] >> kStateSmooth) & 1u))
            s_pre[kStateSmooth] = mail_wait(bstate_in + kStateSmooth, nowait);
        __syncthreads();  // dbuf (own tile), reverse aggregates, successor state visible
        float pc, pstart;
        cross_warp_rev1<NW>(s_W + 7 * NW * NCH * 2, NCH * 2, tb.a2pow, s_pre[kStateSmooth], lane, warp, pc, pstart);
        if (tid == 0) {
            mail_put(bstate_out + kStateSmooth, pstart);
            // release is cumulative over the barrier: every thread's dhead stores made before the
            // __syncthreads above are visible to whoever acquires this flag
            st_release(my_flag, kBFlagComp);
        }
        const float pcarry = fmaf(tb.a_lane[31 - lane], pc, px);  // p at the first sample after this chunk
        if (!halo_early) {  // halo of dy*G from the successor tile
            fetch_halo();
            __syncthreads();
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) lds_chunk<L>(dbuf + c * buf_stride + pLA + pb, u[c]);  // dy*G of x[n] = (dy*G)[n + LA]
        // static-curve sums: S1 = sum p*tc, S2 = sum p*tc^2, S3 = sum p*lin (p = smoothed-gain adjoint)
        float S1 = 0.0f, S2 = 0.0f, S3 = 0.0f;
        const float cside = beta * tb.slope * tb.inv_knee * k20OverLn10;
        float gprev = gcarry;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const float p = fmaf(tb.a_i[L - 1 - i], pcarry, q[i]);
            float side = v[0][i];
            if (NCH > 1) side += v[NCH - 1][i];
            float tc, lin;
            const float gc = gain_computer(side, tb, tc, lin);
            acc_alpha = fmaf(p, gprev - gc, acc_alpha);
            gprev = gs[i];
            const float m = p * tc;
            S1 += m;
            S2 = fmaf(m, tc, S2);
            S3 = fmaf(p, lin, S3);
            // d g_c / d side = slope * tc / knee * 20/(ln10 * side)
            const float dside = (fabsf(side) > kCompEps) ? __fdividef(m * cside, side) : 0.0f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) u[c][i] += dside;
        }
        // d g_c/d thr = -slope*tc/knee; d g_c/d ratio = -(tc^2/(2 knee) + lin)/ratio^2;
        // d g_c/d knee = slope*tc/(2 knee) * (1 - tc/knee); each times d L/d g_c = beta * p
        acc_thr = -beta * tb.slope * tb.inv_knee * S1;
        acc_ratio = -beta * tb.inv_ratio2 * fmaf(tb.inv_2knee, S2, S3);
        acc_knee = beta * tb.slope * tb.inv_2knee * fmaf(-tb.inv_knee, S2, S1);
    } else {
        // no compressor: chain output = EQ output
#pragma unroll
        for (int i0 = 0; i0 < L; i0 += 4) {
            float o[NCH][4], dc[NCH][4];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[c][j] = v[c][i0 + j];
            upstream4(i0, o, dc);
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) u[c][i0 + j] = dc[c][j];
        }
    }
    {
        float t = warp_sum4(acc_alpha, acc_thr, acc_ratio, acc_knee, lane);
        if ((lane & 7) == 0) s_part[warp * kGradCount + kGradAlpha + warp_sum4_slot(lane)] = t;
        t = warp_sum4(acc_makeup, acc_gout, acc_gl, acc_gr, lane);
        const int sl = warp_sum4_slot(lane);
        const int idx = sl == 0 ? kGradMakeup : (sl == 1 ? kGradGout : (sl == 2 ? kGradGL : kGradGR));
        if ((lane & 7) == 0) s_part[warp * kGradCount + idx] = t;
    }

    // ---------------- adjoint of the EQ cascade, sections 5..0 ----------------
    if (a.flags & kChainEq) {
        // y[n-1], y[n-2] at the chunk start for the last section's output
        float ym1[NCH], ym2[NCH];
        {
            float* nb0 = s_nb + 2 * NW * NCH * 2;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                ym1[c] = __shfl_up_sync(0xffffffffu, v[c][L - 1], 1);
                ym2[c] = __shfl_up_sync(0xffffffffu, v[c][L - 2], 1);
                if (lane == 31) { nb0[(warp * NCH + c) * 2 + 0] = v[c][L - 1]; nb0[(warp * NCH + c) * 2 + 1] = v[c][L - 2]; }
            }
            if (tid == 0) sh.next = claimed;
            __syncthreads();
            // the buffer area is no longer read by this tile: start the copies of the next item's inputs
            bwd_prefetch<NCH, L, NT, MASTER>(f, sh.next, area, tab_next, tid);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (lane == 0) {
                    if (warp > 0) { ym1[c] = nb0[((warp - 1) * NCH + c) * 2 + 0]; ym2[c] = nb0[((warp - 1) * NCH + c) * 2 + 1]; }
                    else if (has_pred) { ym1[c] = __ldg(tail_in + (6 * NCH + c) * 2 + 1); ym2[c] = __ldg(tail_in + (6 * NCH + c) * 2 + 0); }
                    else { ym1[c] = 0.0f; ym2[c] = 0.0f; }
                }
            }
        }

        // section states from forward's checkpoints, fetched one section ahead of their use
        float sv1[NCH], sv2[NCH];
        if (saved) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                sv1[c] = __ldg(ssave + (((kNumSections - 1) * NCH + c) * 2 + 0) * NT + tid);
                sv2[c] = __ldg(ssave + (((kNumSections - 1) * NCH + c) * 2 + 1) * NT + tid);
            }
        }
#pragma unroll 1
        for (int k = kNumSections - 1; k >= 0; --k) {
            const SectionTab& st = tb.sec[k];
            float* nb = s_nb + (k & 1) * NW * NCH * 2;  // double-buffered: one barrier per section
            const float b0 = st.b0, b1 = st.b1, b2 = st.b2, na1 = -st.a1, na2 = -st.a2, inv_b0 = st.inv_b0;
            const float qb1 = -b1 * inv_b0, qb2 = -b2 * inv_b0, qa1 = st.a1 * inv_b0, qa2 = st.a2 * inv_b0;
            float xin[NCH][L];
            float r1[NCH], r2[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                // (a) recover the section input from its output along the shared state trajectory
                float s1, s2;
                if (saved) {
                    s1 = sv1[c]; s2 = sv2[c];
                    if (k > 0) {
                        sv1[c] = __ldg(ssave + (((k - 1) * NCH + c) * 2 + 0) * NT + tid);
                        sv2[c] = __ldg(ssave + (((k - 1) * NCH + c) * 2 + 1) * NT + tid);
                    }
                } else {
                    s1 = sE[((k * NCH + c) * 2 + 0) * NT + tid];
                    s2 = sE[((k * NCH + c) * 2 + 1) * NT + tid];
                }
                // (states carried as n = -s/b0: x = y/b0 + n1 is a single fma)
                float n1 = -s1 * inv_b0, n2 = -s2 * inv_b0;
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    const float yv = v[c][i];
                    const float x = fmaf(yv, inv_b0, n1);
                    n1 = fmaf(qb1, x, fmaf(qa1, yv, n2));
                    n2 = fmaf(qb2, x, qa2 * yv);
                    xin[c][i] = x;
                }
                if (lane == 31) { nb[(warp * NCH + c) * 2 + 0] = xin[c][L - 1]; nb[(warp * NCH + c) * 2 + 1] = xin[c][L - 2]; }
                // (b) reverse-time all-pole recursion, zero right-hand state
                // (only its end state is needed: the recursion is run again from the true state below)
                float g1 = 0.0f, g2 = 0.0f;
#pragma unroll
                for (int i = L - 1; i >= 0; --i) {
                    const float gh = fmaf(na1, g1, fmaf(na2, g2, u[c][i]));
                    g2 = g1; g1 = gh;
                }
                r1[c] = g1; r2[c] = g2;
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const float t1 = __shfl_down_sync(0xffffffffu, r1[c], 1 << j);
                    const float t2 = __shfl_down_sync(0xffffffffu, r2[c], 1 << j);
                    if (lane + (1 << j) < 32) mat2T_apply_acc(st.P2[j], t1, t2, r1[c], r2[c]);
                }
            }
            float x1[NCH], x2[NCH], xm1[NCH], xm2[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                x1[c] = __shfl_down_sync(0xffffffffu, r1[c], 1);
                x2[c] = __shfl_down_sync(0xffffffffu, r2[c], 1);
                if (lane == 31) { x1[c] = 0.0f; x2[c] = 0.0f; }
                if (lane == 0) {
                    s_W[((k * NW + warp) * NCH + c) * 2 + 0] = r1[c];
                    s_W[((k * NW + warp) * NCH + c) * 2 + 1] = r2[c];
                }
                xm1[c] = __shfl_up_sync(0xffffffffu, xin[c][L - 1], 1);
                xm2[c] = __shfl_up_sync(0xffffffffu, xin[c][L - 2], 1);
            }
            if (warp == 0 && lane < NCH * 2 && has_succ && !((s_premask[0] >> (k * NCH * 2 + lane)) & 1u))
                s_pre[k * NCH * 2 + lane] = mail_wait(bstate_in + k * NCH * 2 + lane, nowait);  // not prefetched
            __syncthreads();
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float c1, c2, n1, n2;
                cross_warp_rev<NW>(s_W + (k * NW * NCH + c) * 2, NCH * 2, st.P2, s_pre[(k * NCH + c) * 2 + 0],
                                   s_pre[(k * NCH + c) * 2 + 1], lane, warp, c1, c2, n1, n2);
                if (tid == 0) {  // publish right away (value + tag in one store) so the predecessor can go on
                    mail_put(bstate_out + (k * NCH + c) * 2 + 0, n1);
                    mail_put(bstate_out + (k * NCH + c) * 2 + 1, n2);
                }
                mat2T_apply_acc(st.Ppow[31 - lane], c1, c2, x1[c], x2[c]);  // (g[L], g[L+1]) of this chunk
                if (lane == 0) {
                    if (warp > 0) { xm1[c] = nb[((warp - 1) * NCH + c) * 2 + 0]; xm2[c] = nb[((warp - 1) * NCH + c) * 2 + 1]; }
                    else if (has_pred) { xm1[c] = __ldg(tail_in + (k * NCH + c) * 2 + 1); xm2[c] = __ldg(tail_in + (k * NCH + c) * 2 + 0); }
                    else { xm1[c] = 0.0f; xm2[c] = 0.0f; }
                }
            }
            float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                // reverse all-pole recursion from the true right-hand state, FIR B^T and the correlations
                float gp1 = x1[c], gp2 = x2[c];  // g[i+1], g[i+2]
                float dxp = ((L >= 2) ? xin[c][L - 2] : xm1[c]) - xin[c][L - 1];  // x[i-1] - x[i]
#pragma unroll
                for (int i = L - 1; i >= 0; --i) {
                    const float gh = fmaf(na1, gp1, fmaf(na2, gp2, u[c][i]));
                    const float xa = xin[c][i];
                    const float xb = (i >= 1) ? xin[c][i - 1] : xm1[c];
                    const float xc = (i >= 2) ? xin[c][i - 2] : (i == 1 ? xm1[c] : xm2[c]);
                    const float ya = (i >= 1) ? v[c][i - 1] : ym1[c];
                    const float yb = (i >= 2) ? v[c][i - 2] : (i == 1 ? ym1[c] : ym2[c]);
                    // Coefficient gradients in the basis r0 = b0+b1+b2, r1 = b1+2 b2, r2 = b2,
                    // p = a1+a2, q = a2: for poles/zeros near z = 1 the plain (b, a) gradients
                    // are huge and cancel in the parameter Jacobian; differencing the signals
                    // per sample before accumulating keeps float32 sums well conditioned.
                    const float dxn = xc - xb;  // first difference one sample earlier (next iteration's dxp)
                    acc[0] = fmaf(gh, xa, acc[0]);
                    acc[1] = fmaf(gh, dxp, acc[1]);
                    acc[2] = fmaf(gh, dxn - dxp, acc[2]);
                    acc[3] = fmaf(-gh, ya, acc[3]);
                    acc[4] = fmaf(-gh, yb - ya, acc[4]);
                    dxp = dxn;
                    u[c][i] = fmaf(b0, gh, fmaf(b1, gp1, b2 * gp2));
                    gp2 = gp1; gp1 = gh;
                }
#pragma unroll
                for (int i = 0; i < L; ++i) v[c][i] = xin[c][i];
                ym1[c] = xm1[c]; ym2[c] = xm2[c];
            }
            {
                const float t4 = warp_sum4(acc[0], acc[1], acc[2], acc[3], lane);
                if ((lane & 7) == 0) s_part[warp * kGradCount + kGradEq + 5 * k + warp_sum4_slot(lane)] = t4;
                const float t1 = warp_sum(acc[4]);
                if (lane == 0) s_part[warp * kGradCount + kGradEq + 5 * k + 4] = t1;
            }
        }
    }

    // ---------------- input gain and source gradient ----------------
    {
        float acc_gin = 0.0f;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < L; ++i) acc_gin = fmaf(u[c][i], v[c][i], acc_gin);
        const float t = warp_sum(acc_gin);
        if (lane == 0) s_part[warp * kGradCount + kGradGin] = t;
        const float g = (a.flags & kChainGain) ? tb.g_in : 1.0f;
        if constexpr (MASTER) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
#pragma unroll
                for (int i = 0; i < L; ++i) u[c][i] *= g;
                store_chunk<L>(a.gsrc + (long long)(row * NCH + c) * a.Tp + t0, a.Tp - t0, true, u[c]);
            }
        } else if (a.gsrc) {
#pragma unroll
            for (int i = 0; i < L; ++i) u[0][i] *= g;
            store_chunk<L>(a.gsrc + (long long)row * a.T + t0, a.T - t0, uvec, u[0]);
        }
    }
    const bool late_hand_off = !(a.flags & kChainEq);  // no EQ adjoint: the area was in use until here
    if (late_hand_off && tid == 0) sh.next = claimed;
    __syncthreads();
    if (MASTER && tid == 0) st_release(my_flag, kBFlagDone);   // (cumulative over the barrier: every thread's dbus stores)
    if (late_hand_off) bwd_prefetch<NCH, L, NT, MASTER>(f, sh.next, area, tab_next, tid);
    if (tid < kGradCount) {
        float s = 0.0f;
        for (int w = 0; w < NW; ++w) s += s_part[w * kGradCount + tid];
        a.partial[rt * kGradCount + tid] = s;
    }
    return sh.next;
}

// Floats of the per-tile buffer area: delay line + dy*G line (per channel) + upstream gradient
__host__ __device__ inline int bwd_area_floats(int nch, int tile, int la) {
    return 2 * nch * pidx4(la + tile) + 2 * pidx4(tile);
}

template <int NCH, int L, int NT, bool MASTER>
__global__ void __launch_bounds__(NT, 1) chain_bwd_kernel(BwdArgs f) {
    DMST_DYN_SMEM(smem_raw);
    float* area = reinterpret_cast<float*>(smem_raw);
    DMST_SHARED_ARRAY(float, s_tabf, 2 * (sizeof(RowTab) / 4));
    DMST_SHARED_ARRAY(int, s_first, 1);
    typedef BwdShared<NT, NCH> Shared;
    DMST_SHARED_ARRAY(Shared, sh_p, 1);
    Shared& sh = sh_p[0];
    const int tid = threadIdx.x;

    if (MASTER) griddep_launch_dependents();   // every master CTA is resident: the track kernel may fill the other SMs
    if (tid == 0) s_first[0] = atomicAdd(f.ticket, 1);
    __syncthreads();
    int cur = s_first[0];
    int par = 0;
    bwd_prefetch<NCH, L, NT, MASTER>(f, cur, area, s_tabf, tid);
    while (true) {
        cp_async_wait_all();
        __syncthreads();  // this item's inputs have landed; the previous item is completely done
        if (cur >= f.total) break;
        const int tile = f.a.ntiles - 1 - cur / f.a.nrows;
        const int row = cur % f.a.nrows;
        const RowTab& tb = *reinterpret_cast<const RowTab*>(s_tabf + par * (sizeof(RowTab) / 4));
        float* tab_next = s_tabf + (par ^ 1) * (sizeof(RowTab) / 4);
        cur = bwd_tile<NCH, L, NT, MASTER>(f, row, tile, area, tb, tab_next, sh);
        par ^= 1;
    }
}

}  // namespace dmst
