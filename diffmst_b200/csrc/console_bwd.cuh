// Backward chain kernel (adjoint of console_fwd.cuh), one CTA per (row, time tile), tiles
// claimed in REVERSE time order.  Per tile it
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

namespace dmst {

constexpr int kBFlagComp = 1;  // reverse smoother state + dhead halo published (section states use mailboxes)

template <int NCH, int L, int NT, bool MASTER, int MINB>
__global__ void __launch_bounds__(NT, MINB) chain_bwd_kernel(ChainArgs a) {
    constexpr int NW = NT / 32;
    constexpr int TILE = NT * L;
    static_assert(L % 4 == 0 && L <= kMaxL, "chunk length");

    DMST_DYN_SMEM(smem_raw);
    DMST_SHARED_ARRAY(float, s_W, 8 * NW * NCH * 2);       // warp aggregates, slot 7 = reverse smoother
    DMST_SHARED_ARRAY(float, s_nb, 3 * NW * NCH * 2);       // last two samples of each warp: [section parity 0/1, chain output]
    DMST_SHARED_ARRAY(float, s_pre, kStateStride);   // successor's reverse states (prefetched)
    DMST_SHARED_ARRAY(float, s_fst, kStateStride);   // predecessor's forward end states (saved by forward)
    DMST_SHARED_ARRAY(unsigned, s_premask, 2);       // [0] published mailboxes at tile start, [1] successor flag
    DMST_SHARED_ARRAY(float, s_part, NW * kGradCount);
    DMST_SHARED_ARRAY(int, s_ticket, 1);
    DMST_SHARED_ARRAY(float, s_tabf, sizeof(RowTab) / 4);
    const RowTab& tb = *reinterpret_cast<const RowTab*>(s_tabf);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket[0] = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int ticket = s_ticket[0];
    const int tile = a.ntiles - 1 - ticket / a.nrows;
    const int row = ticket % a.nrows;
    const int LA = a.lookahead;
    const int buf_stride = pidx(LA + TILE) + 1;
    // LA % 32 == 0 and L | 32 (checked on the host): pidx(LA + tid*L + i) = pLA + pb + i
    const int pLA = pidx(LA), pb = pidx(tid * L);
    float* ebuf = reinterpret_cast<float*>(smem_raw);          // [NCH][buf_stride]  e delay line
    float* dbuf = ebuf + NCH * buf_stride;                      // [NCH][buf_stride]  dy*G, future halo
    float* sE = dbuf + NCH * buf_stride;                        // [6*NCH*2][NT] forward lane carry-ins

    {
        const float* src = reinterpret_cast<const float*>(a.tab + row);
        for (int i = tid; i < int(sizeof(RowTab) / 4); i += NT) s_tabf[i] = __ldg(src + i);
        for (int i = tid; i < NW * kGradCount; i += NT) s_part[i] = 0.0f;
    }
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
    static_assert(MASTER || L == kBwdChunk, "state checkpoints are spaced by the backward chunk length");
    const float* ssave = a.ssave ? a.ssave + rt * (kNumSections * NCH * 2) * NT : nullptr;
    float v[NCH][L];
    if (saved) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            load_chunk<L>(a.esave + (long long)(row * NCH + c) * a.Tp + t0, a.Tp - t0, true, v[c]);
    } else if constexpr (!MASTER) {
        load_chunk<L>(a.src + (long long)bb * a.src_batch_stride + (long long)nn * a.src_row_stride + t0,
                      a.T - t0, a.src_vec_ok != 0, v[0]);
    } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            load_chunk<L>(a.src + (long long)(row * NCH + c) * a.Tp + t0, a.Tp - t0, true, v[c]);
    }
    __syncthreads();
    if (!saved && (a.flags & kChainGain)) {
        const float g = tb.g_in;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < L; ++i) v[c][i] *= g;
    }

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
    float acc_comp[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // alpha, thr, ratio, knee, makeup

    // Pass over the chunk that forms, per sample, the gradient w.r.t. the compressor output
    // (tracks: gL*dbusL + gR*dbusR [+ grad of mixed_tracks]; master: g_out * dmix) and hands it
    // with the compressor output `o` (or the EQ output when the compressor is off) to `fn`.
    auto for_each_upstream = [&](auto&& out_of, auto&& fn) {
#pragma unroll
        for (int i0 = 0; i0 < L; i0 += 4) {
            float d[NCH][4];
            float bl[4] = {0.f, 0.f, 0.f, 0.f}, br[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (MASTER) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const float4 g4 = load4(a.gout + (long long)(row * NCH + c) * a.T + t0 + i0, a.T - t0 - i0, uvec);
                    d[c][0] = g4.x; d[c][1] = g4.y; d[c][2] = g4.z; d[c][3] = g4.w;
                }
            } else {
                const float4 l4 = load4(a.gout + (long long)(bb * 2 + 0) * a.Tp + t0 + i0, a.Tp - t0 - i0, true);
                const float4 r4 = load4(a.gout + (long long)(bb * 2 + 1) * a.Tp + t0 + i0, a.Tp - t0 - i0, true);
                bl[0] = l4.x; bl[1] = l4.y; bl[2] = l4.z; bl[3] = l4.w;
                br[0] = r4.x; br[1] = r4.y; br[2] = r4.z; br[3] = r4.w;
                if (a.gmixed) {
                    const float4 ml = load4(a.gmixed + ((long long)(bb * 2 + 0) * a.N + nn) * a.T + t0 + i0, a.T - t0 - i0, uvec);
                    const float4 mr = load4(a.gmixed + ((long long)(bb * 2 + 1) * a.N + nn) * a.T + t0 + i0, a.T - t0 - i0, uvec);
                    bl[0] += ml.x; bl[1] += ml.y; bl[2] += ml.z; bl[3] += ml.w;
                    br[0] += mr.x; br[1] += mr.y; br[2] += mr.z; br[3] += mr.w;
                }
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j) d[c][j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j;
                const bool live = (t0 + i) < a.T;
                float dc[NCH], o[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    o[c] = out_of(c, i);
                    if constexpr (MASTER) {
                        float g = live ? d[c][j] : 0.0f;
                        if (a.flags & kChainOutGain) { acc_gout = fmaf(g, o[c] * tb.g_out, acc_gout); g *= tb.g_out; }
                        dc[c] = g;
                    } else {
                        const float l = live ? bl[j] : 0.0f, r = live ? br[j] : 0.0f;
                        acc_gl = fmaf(l, o[c], acc_gl);
                        acc_gr = fmaf(r, o[c], acc_gr);
                        dc[c] = fmaf(tb.gL, l, tb.gR * r);
                    }
                }
                fn(i, dc, o);
            }
        }
    };

    if (a.flags & kChainComp) {
        // ---- forward recompute: delay line, smoothed gain ----
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < L; ++i) ebuf[c * buf_stride + pLA + pb + i] = v[c][i];
        {
            const float* etail_in = a.etail + (rt - 1) * NCH * LA;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                for (int j = tid; j < LA; j += NT)
                    ebuf[c * buf_stride + pidx(j)] = has_pred ? __ldg(etail_in + c * LA + j) : 0.0f;
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
        float Gi = 0.0f;
        for_each_upstream(
            [&](int c, int i) {
                if (c == 0) Gi = fast_exp2(kLog2Per20Db * (gs[i] + tb.makeup));
                return ebuf[c * buf_stride + pb + i] * Gi;
            },
            [&](int i, const float* dc, const float* o) {
                float r = 0.0f;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    r = fmaf(dc[c], o[c], r);
                    const float dyG = dc[c] * Gi;
                    dbuf[c * buf_stride + pb + i] = dyG;
                    if (tid * L < LA) dhead_out[c * LA + tid * L + i] = dyG;  // whole chunk: L | LA
                }
                q[i] = r * kLn10Over20;
                acc_comp[4] += q[i];
            });
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
        if (halo_early) {
            const float* dhead_in = a.dhead + (rt + 1) * NCH * LA;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                for (int j = tid; j < LA; j += NT)
                    dbuf[c * buf_stride + pidx(TILE + j)] = has_succ ? __ldcg(dhead_in + c * LA + j) : 0.0f;
        } else if (tid == 0) {
            wait_flag_ge(succ_flag, kBFlagComp, nowait);
        }
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
            const float* dhead_in = a.dhead + (rt + 1) * NCH * LA;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                for (int j = tid; j < LA; j += NT)
                    dbuf[c * buf_stride + pidx(TILE + j)] = has_succ ? __ldcg(dhead_in + c * LA + j) : 0.0f;
            __syncthreads();
        }
        float gprev = gcarry;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const float p = fmaf(tb.a_i[L - 1 - i], pcarry, q[i]);
            float side = v[0][i];
            if (NCH > 1) side += v[NCH - 1][i];
            float tc, lin;
            const float gc = gain_computer(side, tb, tc, lin);
            acc_comp[0] = fmaf(p, gprev - gc, acc_comp[0]);
            gprev = gs[i];
            const float dgc = beta * p;
            const float dcurve = tb.slope * tc * tb.inv_knee;                 // d g_c / d x_db
            acc_comp[1] = fmaf(dgc, -dcurve, acc_comp[1]);                    // threshold
            acc_comp[2] = fmaf(dgc, -fmaf(tc * tc, tb.inv_2knee, lin) * tb.inv_ratio2, acc_comp[2]);  // ratio
            acc_comp[3] = fmaf(dgc, tb.slope * tc * tb.inv_2knee * (1.0f - tc * tb.inv_knee), acc_comp[3]);  // knee
            const float dside = (fabsf(side) > kCompEps) ? __fdividef(dgc * dcurve * k20OverLn10, side) : 0.0f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) u[c][i] = dbuf[c * buf_stride + pLA + pb + i] + dside;
        }
    } else {
        // no compressor: chain output = EQ output
        for_each_upstream([&](int c, int i) { return v[c][i]; },
                          [&](int i, const float* dc, const float*) {
#pragma unroll
                              for (int c = 0; c < NCH; ++c) u[c][i] = dc[c];
                          });
    }
    {
        float t;
        t = warp_sum(acc_comp[0]); if (lane == 0) s_part[warp * kGradCount + kGradAlpha] = t;
        t = warp_sum(acc_comp[1]); if (lane == 0) s_part[warp * kGradCount + kGradThr] = t;
        t = warp_sum(acc_comp[2]); if (lane == 0) s_part[warp * kGradCount + kGradRatio] = t;
        t = warp_sum(acc_comp[3]); if (lane == 0) s_part[warp * kGradCount + kGradKnee] = t;
        t = warp_sum(acc_comp[4]); if (lane == 0) s_part[warp * kGradCount + kGradMakeup] = t;
        t = warp_sum(acc_gout); if (lane == 0) s_part[warp * kGradCount + kGradGout] = t;
        t = warp_sum(acc_gl); if (lane == 0) s_part[warp * kGradCount + kGradGL] = t;
        t = warp_sum(acc_gr); if (lane == 0) s_part[warp * kGradCount + kGradGR] = t;
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
            __syncthreads();
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
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    const float yv = v[c][i];
                    const float x = (yv - s1) * inv_b0;
                    s1 = fmaf(b1, x, fmaf(na1, yv, s2));
                    s2 = fmaf(b2, x, na2 * yv);
                    xin[c][i] = x;
                }
                if (lane == 31) { nb[(warp * NCH + c) * 2 + 0] = xin[c][L - 1]; nb[(warp * NCH + c) * 2 + 1] = xin[c][L - 2]; }
                // (b) reverse-time all-pole recursion, zero right-hand state
                float g1 = 0.0f, g2 = 0.0f;
#pragma unroll
                for (int i = L - 1; i >= 0; --i) {
                    const float gh = fmaf(na1, g1, fmaf(na2, g2, u[c][i]));
                    g2 = g1; g1 = gh;
                    u[c][i] = gh;
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
                // add the homogeneous reverse response, then FIR B^T and the correlations
                float h1 = x1[c], h2 = x2[c];
#pragma unroll
                for (int i = L - 1; i >= 0; --i) {
                    const float t = fmaf(na1, h1, na2 * h2);
                    u[c][i] += t;
                    h2 = h1; h1 = t;
                }
                float gp1 = x1[c], gp2 = x2[c];  // g[i+1], g[i+2]
#pragma unroll
                for (int i = L - 1; i >= 0; --i) {
                    const float gh = u[c][i];
                    const float xa = xin[c][i];
                    const float xb = (i >= 1) ? xin[c][i - 1] : xm1[c];
                    const float xc = (i >= 2) ? xin[c][i - 2] : (i == 1 ? xm1[c] : xm2[c]);
                    const float ya = (i >= 1) ? v[c][i - 1] : ym1[c];
                    const float yb = (i >= 2) ? v[c][i - 2] : (i == 1 ? ym1[c] : ym2[c]);
                    // Coefficient gradients in the basis r0 = b0+b1+b2, r1 = b1+2 b2, r2 = b2,
                    // p = a1+a2, q = a2: for poles/zeros near z = 1 the plain (b, a) gradients
                    // are huge and cancel in the parameter Jacobian; differencing the signals
                    // per sample before accumulating keeps float32 sums well conditioned.
                    acc[0] = fmaf(gh, xa, acc[0]);
                    acc[1] = fmaf(gh, xb - xa, acc[1]);
                    acc[2] = fmaf(gh, (xa - xb) - (xb - xc), acc[2]);
                    acc[3] = fmaf(-gh, ya, acc[3]);
                    acc[4] = fmaf(-gh, yb - ya, acc[4]);
                    u[c][i] = fmaf(b0, gh, fmaf(b1, gp1, b2 * gp2));
                    gp2 = gp1; gp1 = gh;
                }
#pragma unroll
                for (int i = 0; i < L; ++i) v[c][i] = xin[c][i];
                ym1[c] = xm1[c]; ym2[c] = xm2[c];
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float t = warp_sum(acc[j]);
                if (lane == 0) s_part[warp * kGradCount + kGradEq + 5 * k + j] = t;
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
    __syncthreads();
    if (tid < kGradCount) {
        float s = 0.0f;
        for (int w = 0; w < NW; ++w) s += s_part[w * kGradCount + tid];
        a.partial[rt * kGradCount + tid] = s;
    }
}

}  // namespace dmst
