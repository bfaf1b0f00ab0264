// Forward chain kernel: gain -> 6-section biquad cascade -> compressor (-> output gain),
// one CTA per (row, time tile).  Replaces the dasp_pytorch calls at mst/modules.py:230-251
// (tracks, NCH = 1) and :286-312 (master bus, NCH = 2, fed by the pan + bus sum of
// :262-272).  See chain.cuh for the decomposition.
#pragma once
#include "chain.cuh"

namespace dmst {

template <int NCH, int L, int NT, bool MASTER>
__global__ void __launch_bounds__(NT) chain_fwd_kernel(ChainArgs a) {
    constexpr int NW = NT / 32;
    constexpr int TILE = NT * L;
    static_assert(L % 4 == 0 && L <= kMaxL, "chunk length");

    DMST_DYN_SMEM(smem_raw);
    float* ebuf = reinterpret_cast<float*>(smem_raw);  // [NCH][pidx(LA + TILE) + 1]
    DMST_SHARED_ARRAY(float, s_W, 7 * NW * NCH * 2);
    DMST_SHARED_ARRAY(float, s_pre, kStateStride);  // predecessor's end states ([sec][ch][2], smoother at 24)
    DMST_SHARED_ARRAY(unsigned, s_premask, 1);      // which of them were already published at tile start
    DMST_SHARED_ARRAY(int, s_ticket, 1);
    DMST_SHARED_ARRAY(float, s_tabf, sizeof(RowTab) / 4);
    const RowTab& tb = *reinterpret_cast<const RowTab*>(s_tabf);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // Tiles are claimed in time-major order so that the predecessor tile of any running
    // CTA has already been claimed by a CTA that is running or finished (no deadlock).
    if (tid == 0) s_ticket[0] = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int ticket = s_ticket[0];
    const int tile = ticket / a.nrows;
    const int row = ticket - tile * a.nrows;
    const int LA = a.lookahead;
    const int ebuf_stride = pidx(LA + TILE) + 1;
    // LA % 32 == 0 and L | 32 (checked on the host): pidx(LA + tid*L + i) = pLA + pb + i
    const int pLA = pidx(LA), pb = pidx(tid * L);

    float* tail2 = a.tail2 + ((long long)row * a.ntiles + tile) * kTail2Stride;
    Mail* state_out = a.state + ((long long)row * a.ntiles + tile) * kStateStride;
    const Mail* state_in = a.state + ((long long)row * a.ntiles + tile - 1) * kStateStride;
    int* my_flag = a.flag + (long long)row * a.ntiles + tile;
    const int* pred_flag = my_flag - 1;
    // Prefetch whatever the predecessor tile has already published (usually everything: it was
    // claimed nrows tickets earlier), so the per-section waits below rarely touch global memory.
    const bool nowait = (a.flags & kChainDebugNoWait) != 0;
    if (warp == 0) {
        float pv = 0.0f;
        const bool ok = (tile > 0) ? mail_try(state_in + lane, pv) : true;
        s_pre[lane] = pv;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_premask[0] = m;
    }
    {
        const float* src = reinterpret_cast<const float*>(a.tab + row);
        for (int i = tid; i < int(sizeof(RowTab) / 4); i += NT) s_tabf[i] = __ldg(src + i);
    }

    const int t0 = tile * TILE + tid * L;  // first sample of this thread's chunk
    const int tbase = tile * TILE;          // first sample of the tile
    float* ybuf = ebuf + NCH * ebuf_stride; // [NCH][pidx(TILE) + 1] staging of the chain output
    const int ybuf_stride = pidx(TILE) + 1;
    float v[NCH][L];

    // Source tile -> shared memory with coalesced accesses (staged in the part of the delay line
    // that later receives the EQ output of the very same samples).
    if constexpr (!MASTER) {
        const int b = row / a.N, n = row - b * a.N;
        const float* p = a.src + (long long)b * a.src_batch_stride + (long long)n * a.src_row_stride + tbase;
        stage_in<NT, TILE>(ebuf + pLA, p, a.T - tbase, a.src_vec_ok != 0, tid);
    } else {
        // pan + bus sum (mst/modules.py:262-272): bus_c = sum_n g_c[n] * y[n], fixed order;
        // each thread owns TILE/4/NT float4 columns of the tile, all their loads in flight per track
        constexpr int NQ = TILE / 4 / NT;
        float4 accl[NQ], accr[NQ];
#pragma unroll
        for (int j = 0; j < NQ; ++j) { accl[j] = make_float4(0.f, 0.f, 0.f, 0.f); accr[j] = accl[j]; }
        for (int n = 0; n < a.N; ++n) {
            const int trow = row * a.N + n;
            const float gl = __ldg(&a.track_tab[trow].gL), gr = __ldg(&a.track_tab[trow].gR);
            float4 yv[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int idx = 4 * (tid + j * NT);
                yv[j] = load4(a.src + (long long)trow * a.Tp + tbase + idx, a.Tp - tbase - idx, true);
            }
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                accl[j].x = fmaf(gl, yv[j].x, accl[j].x); accl[j].y = fmaf(gl, yv[j].y, accl[j].y);
                accl[j].z = fmaf(gl, yv[j].z, accl[j].z); accl[j].w = fmaf(gl, yv[j].w, accl[j].w);
                accr[j].x = fmaf(gr, yv[j].x, accr[j].x); accr[j].y = fmaf(gr, yv[j].y, accr[j].y);
                accr[j].z = fmaf(gr, yv[j].z, accr[j].z); accr[j].w = fmaf(gr, yv[j].w, accr[j].w);
            }
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            const int idx = 4 * (tid + j * NT);
            const int pp = pLA + pidx(idx);
            ebuf[pp] = accl[j].x; ebuf[pp + 1] = accl[j].y; ebuf[pp + 2] = accl[j].z; ebuf[pp + 3] = accl[j].w;
            float* er = ebuf + (NCH - 1) * ebuf_stride;
            er[pp] = accr[j].x; er[pp + 1] = accr[j].y; er[pp + 2] = accr[j].z; er[pp + 3] = accr[j].w;
            if (tbase + idx < a.Tp) {  // Tp % 4 == 0: whole float4 or nothing
                *reinterpret_cast<float4*>(a.bus_pre + (long long)(row * NCH + 0) * a.Tp + tbase + idx) = accl[j];
                *reinterpret_cast<float4*>(a.bus_pre + (long long)(row * NCH + NCH - 1) * a.Tp + tbase + idx) = accr[j];
            }
        }
    }
    __syncthreads();  // table and source tile in shared memory
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < L; ++i) v[c][i] = ebuf[c * ebuf_stride + pLA + pb + i];

    if (a.flags & kChainGain) {
        const float g = tb.g_in;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < L; ++i) v[c][i] *= g;
    }
    if (tid == NT - 1) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            tail2[(0 * NCH + c) * 2 + 0] = v[c][L - 2];
            tail2[(0 * NCH + c) * 2 + 1] = v[c][L - 1];
        }
    }

    // ------------------------------ EQ cascade ------------------------------
    if (a.flags & kChainEq) {
#pragma unroll 1
        for (int k = 0; k < kNumSections; ++k) {
            const SectionTab& st = tb.sec[k];
            const float b0 = st.b0, b1 = st.b1, b2 = st.b2, na1 = -st.a1, na2 = -st.a2;
            float s1[NCH], s2[NCH];
            constexpr int NSUB = L / kBwdChunk;           // state checkpoints per thread chunk
            float zm1[NCH][NSUB], zm2[NCH][NSUB];         // zero-state states at the checkpoints
            // zero-state pass over the thread chunk (transposed direct form II)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float z1 = 0.0f, z2 = 0.0f;
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    if (i % kBwdChunk == 0) { zm1[c][i / kBwdChunk] = z1; zm2[c][i / kBwdChunk] = z2; }
                    const float x = v[c][i];
                    const float yv = fmaf(b0, x, z1);
                    z1 = fmaf(b1, x, fmaf(na1, yv, z2));
                    z2 = fmaf(b2, x, na2 * yv);
                    v[c][i] = yv;
                }
                s1[c] = z1; s2[c] = z2;
            }
            // warp inclusive scan of end states, combine operator = P^(2^j)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float* m = st.P2[j];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const float t1 = __shfl_up_sync(0xffffffffu, s1[c], 1 << j);
                    const float t2 = __shfl_up_sync(0xffffffffu, s2[c], 1 << j);
                    if (lane >= (1 << j)) mat2_apply_acc(m, t1, t2, s1[c], s2[c]);
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
            if (warp == 0 && lane < NCH * 2 && !((s_premask[0] >> (k * NCH * 2 + lane)) & 1u))
                s_pre[k * NCH * 2 + lane] = mail_wait(state_in + k * NCH * 2 + lane, nowait);  // not prefetched
            __syncthreads();
            // second-level scan over warps -> warp carry-in; publish the tile end state
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float c1, c2, n1, n2;
                cross_warp_fwd<NW>(s_W + (k * NW * NCH + c) * 2, NCH * 2, st.P2, s_pre[(k * NCH + c) * 2 + 0],
                                   s_pre[(k * NCH + c) * 2 + 1], lane, warp, c1, c2, n1, n2);
                if (tid == 0) {  // publish: value + tag in one 8-byte store, no fence needed
                    mail_put(state_out + (k * NCH + c) * 2 + 0, n1);
                    mail_put(state_out + (k * NCH + c) * 2 + 1, n2);
                }
                // lane carry-in = exclusive prefix + P^lane * C_w
                mat2_apply_acc(st.Ppow[lane], c1, c2, e1[c], e2[c]);
            }
            // add the homogeneous response to the carried-in state
            float* ssave = a.ssave ? a.ssave + ((long long)row * a.ntiles + tile) * (kNumSections * NCH * 2) * (TILE / kBwdChunk)
                                   : nullptr;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float h1 = e1[c], h2 = e2[c];
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    if (i % kBwdChunk == 0 && ssave) {  // true state at this checkpoint
                        const int tb = tid * NSUB + i / kBwdChunk;
                        ssave[((k * NCH + c) * 2 + 0) * (TILE / kBwdChunk) + tb] = zm1[c][i / kBwdChunk] + h1;
                        ssave[((k * NCH + c) * 2 + 1) * (TILE / kBwdChunk) + tb] = zm2[c][i / kBwdChunk] + h2;
                    }
                    const float t = h1;
                    v[c][i] += t;
                    h1 = fmaf(na1, t, h2);
                    h2 = na2 * t;
                }
            }
            if (tid == NT - 1) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    tail2[((k + 1) * NCH + c) * 2 + 0] = v[c][L - 2];
                    tail2[((k + 1) * NCH + c) * 2 + 1] = v[c][L - 1];
                }
            }
        }
    }

    if (a.esave && !(a.flags & kChainComp)) {  // (with the compressor on, e is stored from the delay line below)
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            store_chunk<L>(a.esave + (long long)(row * NCH + c) * a.Tp + t0, a.Tp - t0, true, v[c]);
    }

    // ------------------------------ compressor ------------------------------
    if (a.flags & kChainComp) {
        // EQ output into the delay line (its last LA samples go to the successor tile, below)
        float* etail_out = a.etail + ((long long)row * a.ntiles + tile) * NCH * LA;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int i = 0; i < L; ++i) ebuf[c * ebuf_stride + pLA + pb + i] = v[c][i];
            // halo for the successor: written by the owning threads before the barrier so that the
            // flag can be released right after it (keeps the inter-tile chain short)
            const int off = tid * L - (TILE - LA);
            if (off >= 0) store_chunk<L>(etail_out + c * LA + off, L, true, v[c]);  // L | LA: whole chunks
        }
        // side-chain level -> static gain curve -> zero-state one-pole smoothing
        float g[L];
        float gs = 0.0f;
        {
            const float alpha = tb.alpha, beta = tb.beta;
#pragma unroll
            for (int i = 0; i < L; ++i) {
                float side = v[0][i];
                if (NCH > 1) side += v[NCH - 1][i];
                float tc, lin;
                const float gc = gain_computer(side, tb, tc, lin);
                gs = fmaf(alpha, gs, beta * gc);
                g[i] = gs;
            }
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float t = __shfl_up_sync(0xffffffffu, gs, 1 << j);
            if (lane >= (1 << j)) gs = fmaf(tb.a2pow[j], t, gs);
        }
        float ex = __shfl_up_sync(0xffffffffu, gs, 1);
        if (lane == 0) ex = 0.0f;
        if (lane == 31) s_W[(6 * NW + warp) * NCH * 2] = gs;
        if (tid == 0) {
            if (tile > 0) {
                wait_flag_ge(pred_flag, kFlagSmooth, nowait);  // predecessor's halo (etail) is complete
                if (!((s_premask[0] >> kStateSmooth) & 1u)) s_pre[kStateSmooth] = mail_wait(state_in + kStateSmooth, nowait);
            }
        }
        __syncthreads();
        float cw, gend;
        cross_warp_fwd1<NW>(s_W + 6 * NW * NCH * 2, NCH * 2, tb.a2pow, s_pre[kStateSmooth], lane, warp, cw, gend);
        if (tid == 0) {
            mail_put(state_out + kStateSmooth, gend);
            // release is cumulative over the barrier: every thread's etail stores made before the
            // __syncthreads above are visible to whoever acquires this flag
            st_release(my_flag, kFlagSmooth);
        }
        const float carry = fmaf(tb.a_lane[lane], cw, ex);
        // checkpoint of the EQ output for backward: coalesced store from the delay line
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (a.esave)
                stage_out<NT, TILE>(a.esave + (long long)(row * NCH + c) * a.Tp + tbase, ebuf + c * ebuf_stride + pLA,
                                    a.Tp - tbase, true, tid);
        }
        // halo: predecessor's last LA EQ outputs (zeros before the start of the signal)
        {
            const float* etail_in = a.etail + ((long long)row * a.ntiles + tile - 1) * NCH * LA;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                for (int j = tid; j < LA; j += NT)
                    ebuf[c * ebuf_stride + pidx(j)] = (tile > 0) ? __ldcg(etail_in + c * LA + j) : 0.0f;
        }
        __syncthreads();
        const float makeup = tb.makeup;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const float gtrue = fmaf(tb.a_i[i], carry, g[i]);
            const float G = fast_exp2(kLog2Per20Db * (gtrue + makeup));
#pragma unroll
            for (int c = 0; c < NCH; ++c) v[c][i] = ebuf[c * ebuf_stride + pb + i] * G;
        }
    }

    // ------------------------------ sinks ------------------------------
    if constexpr (!MASTER) {
        // tracks: stage the chain output and store it (and the panned copies) fully coalesced
#pragma unroll
        for (int i = 0; i < L; ++i) ybuf[pb + i] = v[0][i];
        __syncthreads();
        stage_out<NT, TILE>(a.y + (long long)row * a.Tp + tbase, ybuf, a.Tp - tbase, true, tid);
        if (a.want_mixed) {
            const int b = row / a.N, n = row - b * a.N;
            stage_out<NT, TILE>(a.mixed + ((long long)(b * 2 + 0) * a.N + n) * a.T + tbase, ybuf, a.T - tbase,
                                a.user_vec_ok != 0, tid, tb.gL);
            stage_out<NT, TILE>(a.mixed + ((long long)(b * 2 + 1) * a.N + n) * a.T + tbase, ybuf, a.T - tbase,
                                a.user_vec_ok != 0, tid, tb.gR);
        }
    } else {
        // master: few rows, chain-latency bound; keep the footprint small (more tiles resident)
        if (a.flags & kChainOutGain) {
            const float go = tb.g_out;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int i = 0; i < L; ++i) v[c][i] *= go;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            store_chunk<L>(a.mix + (long long)(row * NCH + c) * a.T + t0, a.T - t0, a.user_vec_ok != 0, v[c]);
    }
}

}  // namespace dmst
