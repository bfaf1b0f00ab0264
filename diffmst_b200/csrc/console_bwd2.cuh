// Backward pass of the per-track chains when no gradient w.r.t. the audio is requested (training: the tracks are
// data; mst/system.py:280-338 differentiates the console only w.r.t. the controller's parameters).
//
// Same work decomposition as console_bwd.cuh (persistent CTAs, (row, time tile) items claimed through a ticket in
// reverse time order, next item's inputs prefetched with cp.async), same compressor adjoint, but the EQ parameter
// gradients use the fact that the six sections are LTI operators with zero initial state over the whole signal, i.e.
// lower-triangular Toeplitz operators, which COMMUTE.  With e = H5 ... H0 x, Hk = Bk / Ak, and u = dL/de:
//     de/db_{k,i} = S^i Bk^-1 e,   de/da_{k,i} = -S^i Ak^-1 e      (S = one-sample delay)
//     dL/db_{k,i} = < (Bk^-1)^T u, S^i e >,   dL/da_{k,i} = -< (Ak^-1)^T u, S^i e >
// so all 30 coefficient sums need only u and the EQ output e (the forward pass's checkpoint): no per-section
// intermediate signals, no section-state checkpoints, no inverse recursion that recovers a section's input, no
// propagation of u through the sections (there is nothing upstream that wants it).  Per section two reverse-time
// all-pole recursions on the SAME input u (independent of each other and of the other sections: 12 chains of
// instruction-level parallelism instead of a serial walk through the cascade) and five correlations with e.
//
// The recursions run in delta form (common.cuh, PairTab): state (s, v = s[n] - s[n+1]), c0 = 1 + c1 + c2 held as
// its own float32.  The coefficient gradients are accumulated in the basis {b0+b1+b2, b1+2 b2, b2, a1+a2, a2}
// (console_prepare.cuh's epilogue), whose sums are, after summation by parts, correlations of e with the state
// variables themselves:
//     d/d(b0+b1+b2) =  sum e[n] h[n],   d/d(b1+2b2) = -sum e[n] vh[n],   d/d(b2) = sum e[n] (vh[n] - vh[n+1])
//     d/d(a1+a2)    = -sum e[n-1] g[n], d/d(a2)     =  sum e[n-1] vg[n]
// (h = (Bk/b0)^-T u / b0, g = Ak^-T u): the differences are taken on the smooth filtered signal by the recursion
// itself, never on the white signal e, so no float32 cancellation enters the sums.
//
// Phase 1 (all NT threads, 16-sample chunks): compressor recompute + adjoint -> u, written to shared memory.
// Phase 2 (two thread groups x NT/2 chunks of 32 samples; group g owns sections 3g..3g+2; the two recursions of a
// section are the halves of packed float2 operations): zero-state pass over the chunk, warp scan with constant
// matrix powers, one serial pass over the warp aggregates per recursion (12 lanes, also the tile-to-tile
// hand-off through the mailboxes), true pass fused with the correlations.
//
// Tiles of this kernel (NT * 16 samples) are independent of the forward kernel's: what it needs from forward is the
// EQ-output checkpoint (whole rows, so halos are just earlier samples of it) and the smoother state at its own tile
// boundaries (forward's tile end states plus the mid-tile values forward exports, ChainArgs::gmid).  With 256
// threads and 4096-sample tiles two CTAs share an SM, so the barrier / hand-off phases of one overlap the
// arithmetic of the other.
#pragma once
#include "chain.cuh"
#include "console_bwd.cuh"

namespace dmst {

template <int NT>
struct Bwd2Shared {
    static constexpr int NW = NT / 32;
    static constexpr int NWG = NW / 2;           // warps per phase-2 group
    float W[2 * NW];                              // smoother warp aggregates: forward, reverse
    float sW[kNumRec * NWG * 2];                  // chunk-scan warp aggregates [rec][warp in group][2]
    float sC[kNumRec * NWG * 2];                  // state at each warp's right end [rec][warp in group][2]
    float pre[kStateStride];                      // successor's reverse states: [rec][2] (24), smoother at kStateSmooth
    float fst_smooth;                             // predecessor's forward smoother end state
    unsigned premask[2];                          // [0] published mailboxes at tile start, [1] successor flag
    float part[NW * kGradCount];
    int next, next_row, next_tile, next_bb;       // next work item of this CTA, decoded by the claiming thread
};

struct Bwd2Args {
    BwdArgs b;              // b.a.ntiles: tiles of THIS kernel per row; b.a.state: forward's end-of-tile states
    const EqBwdTab* etab;   // [nrows]
    int fwd_ntiles;         // forward tiles per row
    int fwd_ratio;          // tiles of this kernel per forward tile (forward tile = fwd_ratio * NT * L samples)
    // the master kernel may still be running (programmatic dependent launch): its per-tile flags say which parts
    // of the bus gradient are stored
    const int* mflag;       // [B * m_ntiles], or null: the bus gradient is complete
    int m_ntiles, m_tile_shift;   // master tiles per row, log2(master tile)
};

// Wait (one thread) until the master kernel has stored the bus gradient over the samples of work item (bb, tile)
template <int TILE>
__device__ __forceinline__ void bwd2_wait_bus_gradient(const Bwd2Args& f, int ticket, int tile, int bb) {
    if (!f.mflag || ticket >= f.b.total) return;
    const bool nowait = (f.b.a.flags & kChainDebugNoWait) != 0;
    const int m0 = (tile * TILE) >> f.m_tile_shift;
    int m1 = (tile * TILE + TILE - 1) >> f.m_tile_shift;
    if (m1 > f.m_ntiles - 1) m1 = f.m_ntiles - 1;
    for (int m = m0; m <= m1; ++m) wait_flag_ge(f.mflag + bb * f.m_ntiles + m, kBFlagDone, nowait);
}

// (row, tile) of a ticket: tiles in reverse time order, rows fastest
__device__ __forceinline__ void bwd2_decode(const Bwd2Args& f, int ticket, int& row, int& tile, int& bb) {
    const int q = ticket / f.b.a.nrows;
    row = ticket - q * f.b.a.nrows;
    tile = f.b.a.ntiles - 1 - q;
    bb = row / f.b.a.N;   // batch item
}

// Start the asynchronous copies of the inputs of work item (row, tile) into the (free) buffer area.
template <int L, int NT>
__device__ __forceinline__ void bwd2_prefetch(const Bwd2Args& f, int ticket, int row, int tile, int bb, float* area,
                                              float* ctabbuf, float* etabbuf, int tid) {
    constexpr int TILE = NT * L;
    if (ticket < f.b.total) {
        const ChainArgs& a = f.b.a;
        const float* csrc = reinterpret_cast<const float*>(static_cast<const CompTab*>(a.tab + row));
        for (int i = tid; i < int(sizeof(CompTab) / 16); i += NT) cp_async16(ctabbuf + 4 * i, csrc + 4 * i);
        const bool has_eq = (a.flags & kChainEq) != 0;
        if (has_eq) {
            const float* esrc = reinterpret_cast<const float*>(f.etab + row);
            for (int i = tid; i < int(sizeof(EqBwdTab) / 16); i += NT) cp_async16(etabbuf + 4 * i, esrc + 4 * i);
        }
        const int LA = (a.flags & kChainComp) ? a.lookahead : 0;   // halo: the LA samples before the tile
        const int es = pidx4(a.lookahead + TILE), gs = pidx4(TILE);
        const int tbase = tile * TILE;
        float* gbuf = area + 2 * es;
        float* line = area + pidx4(a.lookahead - LA);
        // chain signal before the compressor over [tbase - LA, tbase + TILE): the EQ-output checkpoint, or (no EQ)
        // the source samples (scaled by the input gain once they have landed); zeros outside the signal
        if (has_eq) {
            const float* src = a.esave + (long long)row * a.Tp + tbase;
            if (LA > 0) cp_tile<NT>(line, src - LA, LA, tile > 0 ? LA : 0, true, tid);
            cp_tile<NT>(area + pidx4(a.lookahead), src, TILE, a.Tp - tbase, true, tid);
        } else {
            const int n = row - bb * a.N;
            const float* src = a.src + (long long)bb * a.src_batch_stride + (long long)n * a.src_row_stride + tbase;
            if (LA > 0) cp_tile<NT>(line, src - LA, LA, tile > 0 ? LA : 0, a.src_vec_ok != 0, tid);
            cp_tile<NT>(area + pidx4(a.lookahead), src, TILE, a.T - tbase, a.src_vec_ok != 0, tid);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)   // bus gradient (B*2, Tp) written by the master launch; zero beyond the signal
            cp_tile<NT>(gbuf + c * gs, a.gout + (long long)(bb * 2 + c) * a.Tp + tbase, TILE, a.T - tbase, true, tid);
    }
    cp_async_commit();
}

template <int L, int NT>
__device__ __forceinline__ int bwd2_tile(const Bwd2Args& f, const int row, const int tile, float* area, const CompTab& tb,
                                         const EqBwdTab& et, float* ctab_next, float* etab_next, Bwd2Shared<NT>& sh) {
    constexpr int NW = NT / 32;
    constexpr int TILE = NT * L;
    constexpr int LC = kBwd2Chunk;             // phase-2 chunk
    constexpr int NCHUNK = TILE / LC;          // phase-2 chunks per tile = threads per group
    constexpr int NG = NT / NCHUNK;            // thread groups
    constexpr int SPG = kNumSections / NG;     // sections per group
    constexpr int NWG = NCHUNK / 32;           // warps per group
    static_assert(L % 4 == 0 && L <= kMaxL && 32 % L == 0, "chunk length");
    static_assert(NG * SPG == kNumSections && NG * NCHUNK == NT && NWG == Bwd2Shared<NT>::NWG, "phase-2 thread groups");
    const ChainArgs& a = f.b.a;
    float* s_part = sh.part;
    float* s_pre = sh.pre;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int LA = a.lookahead;
    const int buf_stride = pidx4(LA + TILE);
    const int gbuf_stride = pidx4(TILE);
    const int pLA = pidx4(LA), pb = pidx4(tid * L);
    float* ebuf = area;                       // [buf_stride]   e delay line: halo, then the tile
    float* dbuf = ebuf + buf_stride;          // [buf_stride]   dy*G and its future halo; then u for phase 2
    float* gbuf = dbuf + buf_stride;          // [2][gbuf_stride] upstream (bus) gradient

    int claimed = 0;
    if (tid == 0) claimed = atomicAdd(f.b.ticket, 1);
    for (int i = tid; i < NW * kGradCount; i += NT) s_part[i] = 0.0f;
    const int t0 = tile * TILE + tid * L;
    const bool has_pred = tile > 0, has_succ = tile < a.ntiles - 1;
    const long long rt = (long long)row * a.ntiles + tile;
    Mail* bstate_out = a.bstate + rt * kStateStride;
    const Mail* bstate_in = a.bstate + (rt + 1) * kStateStride;
    int* my_flag = a.bflag + rt;
    const int* succ_flag = my_flag + 1;
    const bool nowait = (a.flags & kChainDebugNoWait) != 0;
    const bool has_eq = (a.flags & kChainEq) != 0;
    if (warp == 0) {
        float pv = 0.0f;
        const bool ok = has_succ ? mail_try(bstate_in + lane, pv) : true;
        s_pre[lane] = pv;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) { sh.premask[0] = m; sh.premask[1] = has_succ ? (unsigned)ld_acquire(succ_flag) : 99u; }
    } else if (warp == 1 && lane == 0) {
        // forward smoother state just before this tile: a forward tile's end state, or a mid-tile export
        float g0 = 0.0f;
        if (has_pred && (a.flags & kChainComp)) {
            const int ft = tile / f.fwd_ratio, sub = tile - ft * f.fwd_ratio;
            const long long frt = (long long)row * f.fwd_ntiles + ft;
            g0 = (sub == 0) ? a.state[(frt - 1) * kStateStride + kStateSmooth].v
                            : a.gmid[frt * (f.fwd_ratio - 1) + sub - 1];
        }
        sh.fst_smooth = g0;
    }
    const bool uvec = a.user_vec_ok != 0;

    float v[L];
    lds_chunk<L>(ebuf + pLA + pb, v);   // prefetched; landed before the caller's barrier
    if (!has_eq && (a.flags & kChainGain)) {
        const float g = tb.g_in;
#pragma unroll
        for (int i = 0; i < L; ++i) v[i] *= g;
    }
    __syncthreads();  // prefetched states and zeroed partial sums visible

    float u[L];
    float acc_gl = 0.0f, acc_gr = 0.0f;
    float acc_alpha = 0.0f, acc_thr = 0.0f, acc_ratio = 0.0f, acc_knee = 0.0f, acc_makeup = 0.0f;

    // upstream gradient of samples [i0, i0+4) w.r.t. the chain output: gL*dbusL + gR*dbusR [+ grad of mixed_tracks]
    auto upstream4 = [&](int i0, const float (&o)[4], float (&dc)[4]) {
        const float4 l4 = *reinterpret_cast<const float4*>(gbuf + pb + i0);
        const float4 r4 = *reinterpret_cast<const float4*>(gbuf + gbuf_stride + pb + i0);
        float bl[4] = {l4.x, l4.y, l4.z, l4.w}, br[4] = {r4.x, r4.y, r4.z, r4.w};
        if (a.gmixed) {   // (rare: a loss on the returned mixed_tracks)
            const int bb = row / a.N, nn = row - bb * a.N;
            const float4 ml = load4(a.gmixed + ((long long)(bb * 2 + 0) * a.N + nn) * a.T + t0 + i0, a.T - t0 - i0, uvec);
            const float4 mr = load4(a.gmixed + ((long long)(bb * 2 + 1) * a.N + nn) * a.T + t0 + i0, a.T - t0 - i0, uvec);
            bl[0] += ml.x; bl[1] += ml.y; bl[2] += ml.z; bl[3] += ml.w;
            br[0] += mr.x; br[1] += mr.y; br[2] += mr.z; br[3] += mr.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc_gl = fmaf(bl[j], o[j], acc_gl);
            acc_gr = fmaf(br[j], o[j], acc_gr);
            dc[j] = fmaf(tb.gL, bl[j], tb.gR * br[j]);
        }
    };

    // ------------------------------ phase 1: compressor adjoint ------------------------------
    if (a.flags & kChainComp) {
        if (!has_eq) {   // (with EQ the delay line already holds the checkpoint)
            sts_chunk<L>(ebuf + pLA + pb, v);
            if (a.flags & kChainGain)
                for (int j = 4 * tid; j < LA; j += 4 * NT) {
                    float4 h = *reinterpret_cast<float4*>(ebuf + pidx4(j));
                    h.x *= tb.g_in; h.y *= tb.g_in; h.z *= tb.g_in; h.w *= tb.g_in;
                    *reinterpret_cast<float4*>(ebuf + pidx4(j)) = h;
                }
        }
        float gs[L];
        float gz = 0.0f;
        const float alpha = tb.alpha, beta = tb.beta;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            float tc, lin;
            const float gc = gain_computer(v[i], tb, tc, lin);
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
        if (lane == 31) sh.W[warp] = gz;
        __syncthreads();  // ebuf + smoother aggregates visible
        float cw, gend;
        cross_warp_fwd1<NW>(sh.W, 1, tb.a2pow, sh.fst_smooth, lane, warp, cw, gend);
        const float gcarry = fmaf(tb.a_lane[lane], cw, ex);  // g_s just before this chunk
#pragma unroll
        for (int i = 0; i < L; ++i) gs[i] = fmaf(tb.a_i[i], gcarry, gs[i]);

        // adjoint: output -> (delayed signal path, gain path)
        float q[L];
        float* dhead_out = a.dhead + rt * LA;
#pragma unroll
        for (int i0 = 0; i0 < L; i0 += 4) {
            float G[4], o[4], dc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) G[j] = fast_exp2(kLog2Per20Db * (gs[i0 + j] + tb.makeup));
            const float4 e4 = *reinterpret_cast<const float4*>(ebuf + pb + i0);  // x[n - LA]
            o[0] = e4.x * G[0]; o[1] = e4.y * G[1]; o[2] = e4.z * G[2]; o[3] = e4.w * G[3];
            upstream4(i0, o, dc);
            float dyG[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                q[i0 + j] = dc[j] * o[j] * kLn10Over20;
                acc_makeup += q[i0 + j];
                dyG[j] = dc[j] * G[j];
            }
            const float4 d4 = make_float4(dyG[0], dyG[1], dyG[2], dyG[3]);
            *reinterpret_cast<float4*>(dbuf + pb + i0) = d4;
            if (tid * L < LA) *reinterpret_cast<float4*>(dhead_out + tid * L + i0) = d4;  // whole chunk: L | LA
        }
        // reverse one-pole: p[n] = q[n] + alpha p[n+1]
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
        if (lane == 0) sh.W[NW + warp] = pz;
        const bool halo_early = sh.premask[1] >= (unsigned)kBFlagComp;
        auto fetch_halo = [&]() {
            const float* dhead_in = a.dhead + (rt + 1) * LA;
            for (int j = 4 * tid; j < LA; j += 4 * NT) {
                const float4 h = has_succ ? __ldcg(reinterpret_cast<const float4*>(dhead_in + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(dbuf + pidx4(TILE + j)) = h;
            }
        };
        if (halo_early) fetch_halo();
        else if (tid == 0) wait_flag_ge(succ_flag, kBFlagComp, nowait);
        if (tid == 0 && has_succ && !((sh.premask[0] >> kStateSmooth) & 1u))
            s_pre[kStateSmooth] = mail_wait(bstate_in + kStateSmooth, nowait);
        __syncthreads();  // dbuf (own tile), reverse aggregates, successor state visible
        float pc, pstart;
        cross_warp_rev1<NW>(sh.W + NW, 1, tb.a2pow, s_pre[kStateSmooth], lane, warp, pc, pstart);
        if (tid == 0) {
            mail_put(bstate_out + kStateSmooth, pstart);
            st_release(my_flag, kBFlagComp);   // cumulative over the barrier: every thread's dhead stores are visible
        }
        const float pcarry = fmaf(tb.a_lane[31 - lane], pc, px);  // p at the first sample after this chunk
        if (!halo_early) {
            fetch_halo();
            __syncthreads();
        }
        lds_chunk<L>(dbuf + pLA + pb, u);  // dy*G of x[n] = (dy*G)[n + LA]
        float S1 = 0.0f, S2 = 0.0f, S3 = 0.0f;
        const float cside = beta * tb.slope * tb.inv_knee * k20OverLn10;
        float gprev = gcarry;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const float p = fmaf(tb.a_i[L - 1 - i], pcarry, q[i]);
            const float side = v[i];
            float tc, lin;
            const float gc = gain_computer(side, tb, tc, lin);
            acc_alpha = fmaf(p, gprev - gc, acc_alpha);
            gprev = gs[i];
            const float m = p * tc;
            S1 += m;
            S2 = fmaf(m, tc, S2);
            S3 = fmaf(p, lin, S3);
            // d g_c / d side = slope * tc / knee * 20 / (ln10 * side); tc = 0 (so m = 0) wherever |side| is below the
            // knee, in particular near 0: only an exact zero needs the guard
            const float dside = __fdividef(m * cside, side == 0.0f ? 1.0f : side);
            u[i] += dside;
        }
        acc_thr = -beta * tb.slope * tb.inv_knee * S1;
        acc_ratio = -beta * tb.inv_ratio2 * fmaf(tb.inv_2knee, S2, S3);
        acc_knee = beta * tb.slope * tb.inv_2knee * fmaf(-tb.inv_knee, S2, S1);
    } else {
#pragma unroll
        for (int i0 = 0; i0 < L; i0 += 4) {
            float o[4], dc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = v[i0 + j];
            upstream4(i0, o, dc);
#pragma unroll
            for (int j = 0; j < 4; ++j) u[i0 + j] = dc[j];
        }
    }
    // input gain: e is linear in g_in, so dL/dg_in * g_in = <u, e> (u = gradient at the EQ output)
    float acc_gin = 0.0f;
#pragma unroll
    for (int i = 0; i < L; ++i) acc_gin = fmaf(u[i], v[i], acc_gin);
    {
        float t = warp_sum4(acc_alpha, acc_thr, acc_ratio, acc_knee, lane);
        if ((lane & 7) == 0) s_part[warp * kGradCount + kGradAlpha + warp_sum4_slot(lane)] = t;
        t = warp_sum4(acc_makeup, acc_gin, acc_gl, acc_gr, lane);
        const int sl = warp_sum4_slot(lane);
        const int idx = sl == 0 ? kGradMakeup : (sl == 1 ? kGradGin : (sl == 2 ? kGradGL : kGradGR));
        if ((lane & 7) == 0) s_part[warp * kGradCount + idx] = t;
    }
    if (tid == 0) {
        sh.next = claimed;
        bwd2_decode(f, claimed, sh.next_row, sh.next_tile, sh.next_bb);
        bwd2_wait_bus_gradient<TILE>(f, claimed, sh.next_tile, sh.next_bb);   // before the barrier that precedes the prefetch
    }

    // ------------------------------ phase 2: EQ coefficient gradients ------------------------------
    if (has_eq) {
        __syncthreads();                   // every thread has read its dy*G from dbuf
        sts_chunk<L>(dbuf + pb, u);        // u of the tile, pidx4 layout
        // e[tbase - 1]: the sample before the tile (zero before the start of the signal)
        const float e_before_tile = has_pred ? __ldg(a.esave + (long long)row * a.Tp + tile * TILE - 1) : 0.0f;
        __syncthreads();
        const int gi = tid / NCHUNK, ct = tid - gi * NCHUNK;
        const int wg = ct >> 5;
        const PairTab* ptab = et.sec + gi * SPG;
        float e[LC];
        lds_chunk<LC>(ebuf + pLA + pidx4(ct * LC), e);
        float em1;   // e[n - 1] at the chunk start
        if (ct > 0) em1 = ebuf[pLA + pidx4(ct * LC - 1)];
        else em1 = e_before_tile;
        const float* ub = dbuf + pidx4(ct * LC);
        float2 nc0[SPG], k2[SPG];
#pragma unroll
        for (int k = 0; k < SPG; ++k) { nc0[k] = ptab[k].nc0; k2[k] = ptab[k].k2; }
        // zero-state pass over the chunk, last sample first (.x: poles, .y: zeros)
        float2 s[SPG], vv[SPG];
#pragma unroll
        for (int k = 0; k < SPG; ++k) { s[k] = make_float2(0.f, 0.f); vv[k] = s[k]; }
#pragma unroll
        for (int i4 = LC / 4 - 1; i4 >= 0; --i4) {
            const float4 u4 = *reinterpret_cast<const float4*>(ub + 4 * i4);
            const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                const float2 u2 = make_float2(uu[j], uu[j]);
#pragma unroll
                for (int k = 0; k < SPG; ++k) {   // v = k2 v + (u - c0 s); s += v   (3 dependent operations per sample)
                    vv[k] = fma2(k2[k], vv[k], fma2(nc0[k], s[k], u2));
                    s[k] = add2(s[k], vv[k]);
                }
            }
        }
        // reverse warp scan of the chunk start states, combine operator P^(2^j)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
#pragma unroll
            for (int k = 0; k < SPG; ++k) {
                const float2 t1 = shfl_down2(s[k], 1 << j), t2 = shfl_down2(vv[k], 1 << j);
                if (lane + (1 << j) < 32) mat2_apply_acc2(ptab[k].P2[j], t1, t2, s[k], vv[k]);
            }
        }
        float2 x1[SPG], x2[SPG];   // state at the right end of this chunk (zero at the warp's right end so far)
#pragma unroll
        for (int k = 0; k < SPG; ++k) {
            x1[k] = shfl_down2(s[k], 1);
            x2[k] = shfl_down2(vv[k], 1);
            if (lane == 31) { x1[k] = make_float2(0.f, 0.f); x2[k] = x1[k]; }
            if (lane == 0) {
                const int r0 = 2 * (gi * SPG + k);   // recursion index = 2 * section + type
                sh.sW[(r0 * NWG + wg) * 2 + 0] = s[k].x;        sh.sW[(r0 * NWG + wg) * 2 + 1] = vv[k].x;
                sh.sW[((r0 + 1) * NWG + wg) * 2 + 0] = s[k].y;  sh.sW[((r0 + 1) * NWG + wg) * 2 + 1] = vv[k].y;
            }
        }
        if (warp == 0 && lane < 2 * kNumRec && has_succ && !((sh.premask[0] >> lane) & 1u))
            s_pre[lane] = mail_wait(bstate_in + lane, nowait);   // not yet published at tile start
        __syncthreads();
        // e is in registers and u in dbuf: the prefetch targets (delay line, upstream gradient) are free
        bwd2_prefetch<L, NT>(f, sh.next, sh.next_row, sh.next_tile, sh.next_bb, area, ctab_next, etab_next, tid);
        if (tid < kNumRec) {   // serial pass over the warp aggregates of recursion `tid`, right to left
            const float* pm = reinterpret_cast<const float*>(et.sec[tid >> 1].P2[5]) + (tid & 1);
            const float m[4] = {pm[0], pm[2], pm[4], pm[6]};   // P^32 of this recursion
            float c1 = has_succ ? s_pre[2 * tid] : 0.0f, c2 = has_succ ? s_pre[2 * tid + 1] : 0.0f;
#pragma unroll
            for (int w = NWG - 1; w >= 0; --w) {
                sh.sC[(tid * NWG + w) * 2 + 0] = c1;
                sh.sC[(tid * NWG + w) * 2 + 1] = c2;
                float a1 = sh.sW[(tid * NWG + w) * 2 + 0], a2 = sh.sW[(tid * NWG + w) * 2 + 1];
                mat2_apply_acc(m, c1, c2, a1, a2);   // state at the start of warp w
                c1 = a1; c2 = a2;
            }
            mail_put(bstate_out + 2 * tid, c1);   // state at the tile start: the predecessor tile's right end
            mail_put(bstate_out + 2 * tid + 1, c2);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SPG; ++k) {
            const int r0 = 2 * (gi * SPG + k);
            const float2 c1 = make_float2(sh.sC[(r0 * NWG + wg) * 2 + 0], sh.sC[((r0 + 1) * NWG + wg) * 2 + 0]);
            const float2 c2 = make_float2(sh.sC[(r0 * NWG + wg) * 2 + 1], sh.sC[((r0 + 1) * NWG + wg) * 2 + 1]);
            mat2_apply_acc2(ptab[k].Ppow[31 - lane], c1, c2, x1[k], x2[k]);
        }
        // true pass fused with the correlations: accS = sum (e[n-1] g, e[n] h), accV = sum (e[n-1] vg, e[n] vh),
        // accW = sum e[n] wh with wh[n] = vh[n] - vh[n+1] the second difference.  Summed by parts it is
        // sum (e[n] - e[n-1]) vh[n]: the boundary terms e[last] vh[last+1] - e[first-1] vh[first] of neighbouring
        // chunks (and tiles) cancel, e[-1] = 0 and vh = 0 after the end of the signal, so wh is never formed
        float2 accS[SPG], accV[SPG];
        float accW[SPG];
#pragma unroll
        for (int k = 0; k < SPG; ++k) { accS[k] = make_float2(0.f, 0.f); accV[k] = accS[k]; accW[k] = 0.0f; }
#pragma unroll
        for (int i4 = LC / 4 - 1; i4 >= 0; --i4) {
            const float4 u4 = *reinterpret_cast<const float4*>(ub + 4 * i4);
            const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                const int i = 4 * i4 + j;
                const float2 u2 = make_float2(uu[j], uu[j]);
                const float2 ee = make_float2((i > 0) ? e[i > 0 ? i - 1 : 0] : em1, e[i]);
                const float de = ee.y - ee.x;
#pragma unroll
                for (int k = 0; k < SPG; ++k) {
                    x2[k] = fma2(k2[k], x2[k], fma2(nc0[k], x1[k], u2));
                    x1[k] = add2(x1[k], x2[k]);
                    accS[k] = fma2(ee, x1[k], accS[k]);
                    accV[k] = fma2(ee, x2[k], accV[k]);
                    accW[k] = fmaf(de, x2[k].y, accW[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < SPG; ++k) {
            const float sb = ptab[k].scale.y;   // 1 / b0
            const float t4 = warp_sum4(accS[k].y * sb, -accV[k].y * sb, accW[k] * sb, -accS[k].x, lane);
            const int sec = gi * SPG + k;
            if ((lane & 7) == 0) s_part[warp * kGradCount + kGradEq + 5 * sec + warp_sum4_slot(lane)] = t4;
            const float t1 = warp_sum(accV[k].x);
            if (lane == 0) s_part[warp * kGradCount + kGradEq + 5 * sec + 4] = t1;
        }
    }
    __syncthreads();
    if (!has_eq) bwd2_prefetch<L, NT>(f, sh.next, sh.next_row, sh.next_tile, sh.next_bb, area, ctab_next, etab_next, tid);   // (no phase 2: the area was in use until here)
    if (tid < kGradCount) {
        float t = 0.0f;
        for (int w = 0; w < NW; ++w) t += s_part[w * kGradCount + tid];
        a.partial[rt * kGradCount + tid] = t;
    }
    return sh.next;
}

template <int L, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 2 : 1) track_bwd2_kernel(Bwd2Args f) {
    DMST_DYN_SMEM(smem_raw);
    float* area = reinterpret_cast<float*>(smem_raw);
    DMST_SHARED_ARRAY(float, s_ctab, 2 * (sizeof(CompTab) / 4));
    DMST_SHARED_ARRAY(float, s_etab, 2 * (sizeof(EqBwdTab) / 4));
    typedef Bwd2Shared<NT> Shared;
    DMST_SHARED_ARRAY(Shared, sh_p, 1);
    Shared& sh = sh_p[0];
    const int tid = threadIdx.x;

    if (tid == 0) {
        sh.next = atomicAdd(f.b.ticket, 1);
        bwd2_decode(f, sh.next, sh.next_row, sh.next_tile, sh.next_bb);
        bwd2_wait_bus_gradient<NT * L>(f, sh.next, sh.next_tile, sh.next_bb);
    }
    __syncthreads();
    int cur = sh.next, row = sh.next_row, tile = sh.next_tile;
    int par = 0;
    bwd2_prefetch<L, NT>(f, cur, row, tile, sh.next_bb, area, s_ctab, s_etab, tid);
    while (true) {
        cp_async_wait_all();
        __syncthreads();  // this item's inputs have landed; the previous item is completely done
        if (cur >= f.b.total) break;
        const CompTab& tb = *reinterpret_cast<const CompTab*>(s_ctab + par * (sizeof(CompTab) / 4));
        const EqBwdTab& et = *reinterpret_cast<const EqBwdTab*>(s_etab + par * (sizeof(EqBwdTab) / 4));
        cur = bwd2_tile<L, NT>(f, row, tile, area, tb, et, s_ctab + (par ^ 1) * (sizeof(CompTab) / 4),
                               s_etab + (par ^ 1) * (sizeof(EqBwdTab) / 4), sh);
        row = sh.next_row; tile = sh.next_tile;   // (stable until the next item's hand-off, several barriers away)
        par ^= 1;
    }
}

}  // namespace dmst
