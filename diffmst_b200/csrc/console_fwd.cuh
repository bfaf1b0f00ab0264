// Forward pass of the mix console as ONE persistent kernel: per-track chain (gain -> 6-section
// biquad cascade -> compressor) of mst/modules.py:230-251, then pan + bus sum (:262-272) and
// the master-bus chain (:286-312, + output fader) on the stereo bus.  See chain.cuh for the
// decomposition of a chain into (row, time tile) work items.
//
// Work items are claimed through an atomic ticket in an order that interleaves both kinds:
// group g holds the master tiles that cover track time-tile g-lag (for every batch item), then the
// track tiles of time-tile g (for every track); the lag is chosen so that the track tiles a master
// tile sums have normally finished by the time it is claimed.  A work item only ever waits for items with a
// lower ticket (its predecessor tile in time; for a master tile also the N track tiles it sums),
// so whatever a running CTA waits for is running or finished: no deadlock, no co-residency
// assumption.  The master bus is chain-latency bound (few rows); interleaved like this its
// latency hides under the track work and its bus sum reads the track outputs from L2.
//
// CTAs are persistent: while a tile is being processed the next ticket is claimed and its
// inputs (row table; for track tiles the source samples) are prefetched into shared memory
// with cp.async, so the HBM latency of a tile's first touch is off the critical path.
#pragma once
#include "chain.cuh"

namespace dmst {

struct FwdArgs {
    ChainArgs t, m;   // per-track chains / master bus chains
    int B;            // batch items
    int R;            // master tiles per track tile
    int group;        // tickets per group = B*R + B*N
    int lag;          // the master tiles in group g cover track time-tile g - lag
    int total;        // (track tiles + lag) * group
    int* ticket;
    int* done;        // [B * track tiles]: tracks of the item that finished the time tile
};

struct FwdWork { int role, row, tile, b; };  // role 0: nothing, 1: track tile, 2: master tile; b = batch item of the row

__device__ __forceinline__ FwdWork fwd_decode(const FwdArgs& f, int ticket) {
    FwdWork w{0, 0, 0, 0};
    if (ticket >= f.total) return w;
    const int g = ticket / f.group, r = ticket - g * f.group;
    const int nm = f.B * f.R;
    if (r < nm) {  // master tiles of track time-tile g-lag, earlier tile first
        const int mi = r / f.B, b = r - mi * f.B;
        const int mt = (g - f.lag) * f.R + mi;
        if (g >= f.lag && mt < f.m.ntiles) { w.role = 2; w.row = b; w.tile = mt; w.b = b; }
    } else if (g < f.t.ntiles) {
        w.role = 1; w.row = r - nm; w.tile = g; w.b = w.row / f.t.N;
    }
    return w;
}

template <int L>
__device__ __forceinline__ void lds_chunk(const float* p, float (&v)[L]) {
#pragma unroll
    for (int i = 0; i < L / 4; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(p + 4 * i);
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
}
template <int L>
__device__ __forceinline__ void sts_chunk(float* p, const float (&v)[L]) {
#pragma unroll
    for (int i = 0; i < L / 4; ++i)
        *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// 4 consecutive floats written by another CTA of this launch (L2-coherent load), zero beyond `valid`
__device__ __forceinline__ float4 load4_cg(const float* p, int valid) {
    if (valid >= 4) return __ldcg(reinterpret_cast<const float4*>(p));
    float4 r;
    r.x = valid > 0 ? __ldcg(p) : 0.0f;
    r.y = valid > 1 ? __ldcg(p + 1) : 0.0f;
    r.z = valid > 2 ? __ldcg(p + 2) : 0.0f;
    r.w = valid > 3 ? __ldcg(p + 3) : 0.0f;
    return r;
}
// Coalesced tile output from the pidx4 layout: the CTA moves TILE consecutive floats with 128-bit accesses
template <int NT, int TILE>
__device__ __forceinline__ void stage_out4(float* g, const float* stage, int valid, bool vec_ok, int tid, float scale = 1.0f) {
#pragma unroll
    for (int q = tid; q < TILE / 4; q += NT) {
        const int idx = 4 * q, left = valid - idx;
        float4 val = *reinterpret_cast<const float4*>(stage + pidx4(idx));
        val.x *= scale; val.y *= scale; val.z *= scale; val.w *= scale;
        if (vec_ok && left >= 4) {
            *reinterpret_cast<float4*>(g + idx) = val;
        } else {
            if (left > 0) g[idx] = val.x;
            if (left > 1) g[idx + 1] = val.y;
            if (left > 2) g[idx + 2] = val.z;
            if (left > 3) g[idx + 3] = val.w;
        }
    }
}

// Start the asynchronous copies of a work item's inputs: its row table and, for a track tile, the
// TILE source samples (zero beyond the end of the signal) into `inbuf` (pidx4 layout).
template <int NT, int TILE_T>
__device__ __forceinline__ void fwd_prefetch(const FwdArgs& f, const FwdWork& w, float* inbuf, float* tabbuf, int tid) {
    if (w.role != 0) {
        const ChainArgs& a = (w.role == 1) ? f.t : f.m;
        const float* src = reinterpret_cast<const float*>(a.tab + w.row);
        for (int i = tid; i < int(sizeof(RowTab) / 16); i += NT) cp_async16(tabbuf + 4 * i, src + 4 * i);
    }
    if (w.role == 1) {
        const ChainArgs& a = f.t;
        const int b = w.b, n = w.row - b * a.N;
        const int tbase = w.tile * TILE_T;
        const float* p = a.src + (long long)b * a.src_batch_stride + (long long)n * a.src_row_stride + tbase;
        const int valid = a.T - tbase;
        const bool vec = a.src_vec_ok != 0;
        if (vec && valid >= TILE_T) {
            // interior tile: pidx4(4(q+NT)) - pidx4(4q) is constant, no per-copy index arithmetic
            float* d = inbuf + pidx4(tid << 2);
            const float* s = p + (tid << 2);
#pragma unroll
            for (int q = 0; q < TILE_T / 4 / NT; ++q) {
                cp_async16(d, s);
                d += 4 * NT + (NT >> 1);
                s += 4 * NT;
            }
        } else {
            for (int q = tid; q < TILE_T / 4; q += NT) {
                const int idx = 4 * q;
                float* d = inbuf + pidx4(idx);
                if (vec && valid - idx >= 4) {
                    cp_async16(d, p + idx);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (idx + e < valid) cp_async4(d + e, p + idx + e);
                        else d[e] = 0.0f;
                    }
                }
            }
        }
    }
    cp_async_commit();
}

// Shared-memory scratch common to both roles
template <int NT>
struct FwdShared {
    float W[7 * (NT / 32) * 2 * 2];   // warp aggregates: [stage 0..6][warp][ch][2]
    float pre[kStateStride];           // predecessor's end states ([sec][ch][2], smoother at 24)
    unsigned premask;                  // which of them were already published at tile start
    int next;                          // next ticket of this CTA
    FwdWork next_work;                 // ... decoded by the claiming thread (the integer divisions run once per tile)
};

// One (row, tile) of a chain.  Returns the next ticket of this CTA (claimed on the way, inputs prefetched).
// CHK: spacing of the section-state checkpoints left for backward = thread chunk of the backward kernel of this role
template <int NCH, int L, int NT, bool MASTER, int TILE_T, int CHK>
__device__ __forceinline__ int fwd_tile(const FwdArgs& f, const ChainArgs& a, const int row, const int tile, const int row_b,
                                        float* ebuf, float* inbuf, const RowTab& tb, float* tab_next,
                                        FwdShared<NT>& sh) {
    constexpr int NW = NT / 32;
    constexpr int TILE = NT * L;
    static_assert(L % 4 == 0 && L <= kMaxL && 32 % L == 0, "chunk length");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* s_W = sh.W;
    float* s_pre = sh.pre;

    const int LA = a.lookahead;
    const int ebuf_stride = pidx4(LA + TILE);
    // LA % 32 == 0 and L | 32 (checked on the host): pidx4(LA + tid*L + i) = pLA + pb + i
    const int pLA = pidx4(LA), pb = pidx4(tid * L);

    float* tail2 = a.tail2 + ((long long)row * a.ntiles + tile) * kTail2Stride;
    Mail* state_out = a.state + ((long long)row * a.ntiles + tile) * kStateStride;
    const Mail* state_in = a.state + ((long long)row * a.ntiles + tile - 1) * kStateStride;
    int* my_flag = a.flag + (long long)row * a.ntiles + tile;
    const int* pred_flag = my_flag - 1;
    const bool nowait = (a.flags & kChainDebugNoWait) != 0;
    const bool has_eq = (a.flags & kChainEq) != 0;

    // Prefetch whatever the predecessor tile has already published (usually everything), so the
    // per-section waits below rarely touch global memory.  (Read after the first barrier below.)
    if (warp == 0) {
        float pv = 0.0f;
        const bool ok = (tile > 0) ? mail_try(state_in + lane, pv) : true;
        s_pre[lane] = pv;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) sh.premask = m;
        __syncwarp();   // section 0 reads premask / s_pre in this warp before the first CTA barrier (racecheck)
    }
    // Claim the next work item early in the tile; the ticket stays in a register until it is handed to
    // the other threads a few barriers later, so the atomic's latency is never waited for.  (Not at the
    // very start of the tile: items claimed long before they run unbalance the tail of the launch.)
    int claimed = 0;
    if (!MASTER && !has_eq && tid == 0) claimed = atomicAdd(f.ticket, 1);

    const int t0 = tile * TILE + tid * L;  // first sample of this thread's chunk
    const int tbase = tile * TILE;          // first sample of the tile
    float v[NCH][L];

    if constexpr (!MASTER) {
        lds_chunk<L>(inbuf + pb, v[0]);  // prefetched source samples (landed before the caller's barrier)
    } else {
        // the N track tiles this bus tile sums must be complete
        if (tid == 0) wait_flag_ge(f.done + (long long)row * f.t.ntiles + tbase / TILE_T, a.N, nowait);
        __syncthreads();
        // pan + bus sum (mst/modules.py:262-272): bus_c = sum_n g_c[n] * y[n], fixed order;
        // each thread owns TILE/4/NT float4 columns of the tile
        constexpr int NQ = TILE / 4 / NT;
        float4 accl[NQ], accr[NQ];
#pragma unroll
        for (int j = 0; j < NQ; ++j) { accl[j] = make_float4(0.f, 0.f, 0.f, 0.f); accr[j] = accl[j]; }
#pragma unroll 2
        for (int n = 0; n < a.N; ++n) {
            const int trow = row * a.N + n;
            const float gl = __ldg(&a.track_tab[trow].gL), gr = __ldg(&a.track_tab[trow].gR);
            float4 yv[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int idx = 4 * (tid + j * NT);
                yv[j] = load4_cg(a.src + (long long)trow * a.Tp + tbase + idx, a.Tp - tbase - idx);
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
            const int pp = pLA + pidx4(idx);
            *reinterpret_cast<float4*>(ebuf + pp) = accl[j];
            *reinterpret_cast<float4*>(ebuf + (NCH - 1) * ebuf_stride + pp) = accr[j];
            if (tbase + idx < a.Tp) {  // Tp % 4 == 0: whole float4 or nothing
                *reinterpret_cast<float4*>(a.bus_pre + (long long)(row * NCH + 0) * a.Tp + tbase + idx) = accl[j];
                *reinterpret_cast<float4*>(a.bus_pre + (long long)(row * NCH + NCH - 1) * a.Tp + tbase + idx) = accr[j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < NCH; ++c) lds_chunk<L>(ebuf + c * ebuf_stride + pLA + pb, v[c]);
        // a bus tile claims its next item only now: while it waited for its tracks and summed them, an item
        // claimed earlier would have been withheld from the CTAs that are free to run it
        if (tid == 0) claimed = atomicAdd(f.ticket, 1);
    }

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
    if (has_eq) {
#pragma unroll 1
        for (int k = 0; k < kNumSections; ++k) {
            const SectionTab& st = tb.sec[k];
            const float b0 = st.b0, b1 = st.b1, b2 = st.b2, na1 = -st.a1, na2 = -st.a2;
            float s1[NCH], s2[NCH];
            static_assert(L % CHK == 0, "checkpoint spacing divides the forward chunk");
            constexpr int NSUB = L / CHK;  // state checkpoints per thread chunk
            float zm1[NCH][NSUB], zm2[NCH][NSUB];                  // zero-state states at the checkpoints
            if (tid == 0) {
                if (!MASTER && k == 1) claimed = atomicAdd(f.ticket, 1);
                if (k == kNumSections - 2) { sh.next = claimed; sh.next_work = fwd_decode(f, claimed); }  // visible after this section's barrier
            }
            // zero-state pass over the thread chunk (transposed direct form II)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float z1 = 0.0f, z2 = 0.0f;
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    if (i % CHK == 0) { zm1[c][i / CHK] = z1; zm2[c][i / CHK] = z2; }
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
            // (premask/pre of sections k >= 1 were made visible by the barrier of section k-1; for
            // k == 0 warp 0 itself wrote them)
            if (warp == 0 && lane < NCH * 2 && !((sh.premask >> (k * NCH * 2 + lane)) & 1u))
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
            float* ssave = a.ssave
                               ? a.ssave + ((long long)row * a.ntiles + tile) * (kNumSections * NCH * 2) * (TILE / CHK)
                               : nullptr;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float h1 = e1[c], h2 = e2[c];
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    if (i % CHK == 0 && ssave) {  // true state at this checkpoint
                        const int tbx = tid * NSUB + i / CHK;
                        ssave[((k * NCH + c) * 2 + 0) * (TILE / CHK) + tbx] = zm1[c][i / CHK] + h1;
                        ssave[((k * NCH + c) * 2 + 1) * (TILE / CHK) + tbx] = zm2[c][i / CHK] + h2;
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
    } else {
        if (tid == 0) { sh.next = claimed; sh.next_work = fwd_decode(f, claimed); }
        __syncthreads();  // the claimed ticket (and warp 0's prefetched states) visible to all
    }
    // Every thread is past a barrier that follows its reads of `inbuf`: start the next item's copies.
    const int next_ticket = sh.next;
    fwd_prefetch<NT, TILE_T>(f, sh.next_work, inbuf, tab_next, tid);

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
            sts_chunk<L>(ebuf + c * ebuf_stride + pLA + pb, v[c]);
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
                if (!((sh.premask >> kStateSmooth) & 1u)) s_pre[kStateSmooth] = mail_wait(state_in + kStateSmooth, nowait);
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
        const float carry = fmaf(tb.a_lane[lane], cw, ex);   // g_s just before this thread's chunk
        if (a.gmid && tid > 0 && ((tid * L) & ((1 << a.gmid_shift) - 1)) == 0)   // the backward kernel's tile boundaries inside this tile
            a.gmid[((long long)row * a.ntiles + tile) * ((TILE >> a.gmid_shift) - 1) + ((tid * L) >> a.gmid_shift) - 1] = carry;
        // checkpoint of the EQ output for backward: coalesced store from the delay line
        if (a.esave) {
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                stage_out4<NT, TILE>(a.esave + (long long)(row * NCH + c) * a.Tp + tbase, ebuf + c * ebuf_stride + pLA,
                                     a.Tp - tbase, true, tid);
        }
        // halo: predecessor's last LA EQ outputs (zeros before the start of the signal)
        {
            const float* etail_in = a.etail + ((long long)row * a.ntiles + tile - 1) * NCH * LA;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                for (int j = 4 * tid; j < LA; j += 4 * NT) {
                    const float4 h = (tile > 0) ? __ldcg(reinterpret_cast<const float4*>(etail_in + c * LA + j))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(ebuf + c * ebuf_stride + pidx4(j)) = h;
                }
        }
        __syncthreads();
        const float makeup = tb.makeup;
        float G[L];
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const float gtrue = fmaf(tb.a_i[i], carry, g[i]);
            G[i] = fast_exp2(kLog2Per20Db * (gtrue + makeup));
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            lds_chunk<L>(ebuf + c * ebuf_stride + pb, v[c]);  // x[n - LA]
#pragma unroll
            for (int i = 0; i < L; ++i) v[c][i] *= G[i];
        }
    }

    // ------------------------------ sinks ------------------------------
    if constexpr (!MASTER) {
        // tracks: stage the chain output in the (now consumed) delay line and store it (and the
        // panned copies) fully coalesced
        __syncthreads();  // every thread has read its delayed samples
        sts_chunk<L>(ebuf + pb, v[0]);
        __syncthreads();
        stage_out4<NT, TILE>(a.y + (long long)row * a.Tp + tbase, ebuf, a.Tp - tbase, true, tid);
        if (a.want_mixed) {
            const int b = row_b, n = row - b * a.N;
            stage_out4<NT, TILE>(a.mixed + ((long long)(b * 2 + 0) * a.N + n) * a.T + tbase, ebuf, a.T - tbase,
                                 a.user_vec_ok != 0, tid, tb.gL);
            stage_out4<NT, TILE>(a.mixed + ((long long)(b * 2 + 1) * a.N + n) * a.T + tbase, ebuf, a.T - tbase,
                                 a.user_vec_ok != 0, tid, tb.gR);
        }
    } else {
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
    return next_ticket;
}

// Dynamic shared memory of the forward kernel (floats): delay line(s) + prefetch buffer
__host__ __device__ inline int fwd_ebuf_floats(int tile_t, int la_t, int tile_m, int la_m) {
    const int t = pidx4(la_t + tile_t), m = 2 * pidx4(la_m + tile_m);
    return t > m ? t : m;
}

template <int L_T, int L_M, int NT, int CHK_T, int CHK_M>
__global__ void __launch_bounds__(NT, 2) console_fwd_kernel(FwdArgs f) {
    constexpr int TILE_T = NT * L_T;
    DMST_DYN_SMEM(smem_raw);
    float* ebuf = reinterpret_cast<float*>(smem_raw);
    float* inbuf = ebuf + fwd_ebuf_floats(TILE_T, f.t.lookahead, NT * L_M, f.m.lookahead);
    DMST_SHARED_ARRAY(float, s_tabf, 2 * (sizeof(RowTab) / 4));
    DMST_SHARED_ARRAY(FwdShared<NT>, sh_p, 1);
    FwdShared<NT>& sh = sh_p[0];
    const int tid = threadIdx.x;

    if (tid == 0) { sh.next = atomicAdd(f.ticket, 1); sh.next_work = fwd_decode(f, sh.next); }
    __syncthreads();
    int cur = sh.next;
    FwdWork w = sh.next_work;
    int par = 0;
    fwd_prefetch<NT, TILE_T>(f, w, inbuf, s_tabf, tid);
    int* signal = nullptr;  // completion counter of the track tile just processed
    while (true) {
        cp_async_wait_all();
        __syncthreads();  // this item's inputs have landed; the previous item is completely done
        if (signal != nullptr && tid == 0) red_release_add(signal, 1);  // (cumulative over the barrier)
        signal = nullptr;
        if (cur >= f.total) break;
        const RowTab& tb = *reinterpret_cast<const RowTab*>(s_tabf + par * (sizeof(RowTab) / 4));
        float* tab_next = s_tabf + (par ^ 1) * (sizeof(RowTab) / 4);
        int nxt;
        if (w.role == 1) {
            nxt = fwd_tile<1, L_T, NT, false, TILE_T, CHK_T>(f, f.t, w.row, w.tile, w.b, ebuf, inbuf, tb, tab_next, sh);
            signal = f.done + (long long)w.b * f.t.ntiles + w.tile;
        } else if (w.role == 2) {
            nxt = fwd_tile<2, L_M, NT, true, TILE_T, CHK_M>(f, f.m, w.row, w.tile, w.b, ebuf, inbuf, tb, tab_next, sh);
        } else {
            if (tid == 0) { sh.next = atomicAdd(f.ticket, 1); sh.next_work = fwd_decode(f, sh.next); }
            __syncthreads();
            nxt = sh.next;
            fwd_prefetch<NT, TILE_T>(f, sh.next_work, inbuf, tab_next, tid);
        }
        cur = nxt;
        w = sh.next_work;   // (stable until the next item's hand-off, several barriers away)
        par ^= 1;
    }
}

}  // namespace dmst
