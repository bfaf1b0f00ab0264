// Host-side launch logic of the mix console behind the C ABI (include/diffmst_b200.h).
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/diffmst_b200.h"
#include "console_bwd.cuh"
#include "console_bwd2.cuh"
#include "console_fwd.cuh"
#include "console_prepare.cuh"

namespace dmst {

// Tile geometry.  Forward and backward must agree on NT*L per row kind because backward
// restarts each tile from the carry-in states the forward pass saved.
// (Measured on B200: 2048...8192-sample tiles and 128...512-thread CTAs are within 5 % of each other for
// the track chain; occupancy above 16 warps per SM does not pay)
constexpr int kTrackFwdL = 32, kTrackFwdNT = 256;   // 8192-sample tiles
constexpr int kTrackBwdL = 16, kTrackBwdNT = 512;
constexpr int kMasterL = 16, kMasterNT = 256;       // 4096-sample tiles, 2 channels per thread
// The master bus has few rows and is bound by the tile-to-tile chain: backward spreads a tile over
// more threads (shorter chunks) to shorten every stage of that chain.
#ifndef DMST_MASTER_BWD_NT
#define DMST_MASTER_BWD_NT 256
#endif
constexpr int kMasterBwdNT = DMST_MASTER_BWD_NT;
constexpr int kMasterBwdL = kMasterL * kMasterNT / kMasterBwdNT;
static_assert(kMasterBwdL * kMasterBwdNT == kMasterL * kMasterNT && kMasterBwdL % 4 == 0, "master tile in the backward CTA shape");
// track backward without audio gradient (console_bwd2.cuh): its own, smaller tiles, two CTAs per SM
constexpr int kTrackBwd2L = 16, kTrackBwd2NT = 256;
constexpr int kTrackBwd2Tile = kTrackBwd2L * kTrackBwd2NT;
constexpr int kTrackBwd2Shift = 12;
static_assert((1 << kTrackBwd2Shift) == kTrackBwd2Tile, "power-of-two tile");
constexpr int kTrackTile = kTrackFwdL * kTrackFwdNT;
static_assert(kTrackTile % kTrackBwd2Tile == 0, "backward tiles nest in forward tiles");
constexpr int kMasterTile = kMasterL * kMasterNT;
constexpr int kMasterTileShift = 12;
static_assert((1 << kMasterTileShift) == kMasterTile, "power-of-two master tile");
// CTAs of the master backward kernel while the track kernel runs beside it (see console_backward).  Both kernels are
// work bound once the master chain has a few tiles in flight: a master tile costs about 25 us of one CTA, a track tile
// 7.3 us of one SM (measured on B200), so the two finish together when (SMs - C) / C = 0.293 N; C is set 15 % above
// that, because a master kernel that finishes last stalls the track kernel, while one that finishes early only idles
// its own SMs.  Measured, ms per step at B = 8, N = 16 (graph replay): 16 CTAs 1.46, 24 1.236, 32 1.198, 48 1.204,
// 64 1.207, all 148 1.217, no overlap 1.239; console forward + backward alone at B = 16, N = 4: 32 CTAs 1.04 ms,
// 48 0.80, 64 0.69, 96 0.73, 148 0.76.
inline int master_bwd_overlap_ctas(int sms, int B, int N) {
    int c = (int)(sms * 1.15 / (1.0 + 0.293 * N) + 0.5);
    if (c > 24 * B) c = 24 * B;   // (more than ~24 tiles of one chain in flight only wait for each other)
    if (c < 8) c = 8;
    return c;
}
static_assert(kTrackFwdL * kTrackFwdNT == kTrackBwdL * kTrackBwdNT, "forward/backward tiles must agree");
static_assert(kTrackBwdL == kBwdChunk && kTrackFwdL % kBwdChunk == 0, "checkpoint spacing");

#ifdef DMST_EMULATE
#define DMST_MEMSET_ASYNC(ptr, val, bytes, stream) (memset((ptr), (val), (bytes)), 0)
#define DMST_LAST_ERROR() 0
#define DMST_SET_SMEM(kernel, bytes) 0
#else
#define DMST_MEMSET_ASYNC(ptr, val, bytes, stream) (int)cudaMemsetAsync((ptr), (val), (bytes), (stream))
#define DMST_LAST_ERROR() (int)cudaGetLastError()
#define DMST_SET_SMEM(kernel, bytes) \
    (int)cudaFuncSetAttribute((kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif

// Optional per-kernel timing (bench.py's roofline figure): when enabled, every launch of a chain
// kernel is bracketed by a pair of CUDA events recorded on the launching stream; nothing
// synchronises until dmst_profile_read.
struct Profiler {
    static constexpr int kKinds = 4;  // 0 track fwd, 1 master fwd, 2 master bwd, 3 track bwd
#ifndef DMST_EMULATE
    cudaEvent_t* ev[kKinds] = {nullptr, nullptr, nullptr, nullptr};
    int count[kKinds] = {0, 0, 0, 0};
    int cap = 0;
#endif
    bool enabled = false;
};
inline Profiler& profiler() { static Profiler p; return p; }
#ifndef DMST_EMULATE
struct ScopedTimer {
    int kind; cudaStream_t s; bool on;
    ScopedTimer(int k, cudaStream_t st) : kind(k), s(st) {
        Profiler& p = profiler();
        on = p.enabled && p.count[k] < p.cap;
        if (on) cudaEventRecord(p.ev[k][2 * p.count[k]], s);
    }
    ~ScopedTimer() {
        if (!on) return;
        Profiler& p = profiler();
        cudaEventRecord(p.ev[kind][2 * p.count[kind] + 1], s);
        p.count[kind]++;
    }
};
#else
struct ScopedTimer { ScopedTimer(int, cudaStream_t) {} };
#endif

struct Carver {
    unsigned char* base;
    size_t off;
    template <class T>
    T* take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct ConsoleWs {
    int* header;          // [0..3] tickets (fwd track, fwd master, bwd master, bwd track)
    RowTab *track_tab, *track_tab_b, *master_tab, *master_tab_b;  // *_b: tables for the backward chunk length
    EqBwdTab* track_etab;  // recursion tables of track_bwd2_kernel
    float *y, *bus_pre, *dbus, *esave, *ssave, *m_esave, *m_ssave, *gmid;
    // forward chain (kept for backward)
    int *t_flag, *m_flag, *t_done;
    Mail *t_state, *m_state;
    float *t_tail2, *m_tail2, *t_etail, *m_etail;
    // backward chain
    int *t_bflag, *m_bflag;
    Mail *t_bstate, *m_bstate;
    float *t_dhead, *m_dhead, *t_partial, *m_partial;
    size_t flags_begin, flags_end;    // byte range of forward flags (zeroed per forward)
    size_t bflags_begin, bflags_end;  // byte range of backward flags
    int Tp, nt_track, nt_master, nt_track_b2;
    size_t total;
};

// flags: DMST_WANT_GRAD_TRACKS -> the per-section state checkpoints the classic track backward needs;
//        DMST_FORWARD_ONLY     -> no checkpoints at all (no backward will follow)
inline ConsoleWs carve_console(void* base, int B, int N, int T, int la_t, int la_m, unsigned flags) {
    const bool fwd_only = (flags & DMST_FORWARD_ONLY) != 0;
    const bool classic_track_bwd = (flags & DMST_WANT_GRAD_TRACKS) != 0 && !fwd_only;
    ConsoleWs w;
    Carver c{reinterpret_cast<unsigned char*>(base), 0};
    w.Tp = (T + 3) & ~3;
    w.nt_track = (T + kTrackTile - 1) / kTrackTile;
    w.nt_master = (T + kMasterTile - 1) / kMasterTile;
    w.nt_track_b2 = (T + kTrackBwd2Tile - 1) / kTrackBwd2Tile;
    const size_t rows = (size_t)B * N, rt = rows * w.nt_track, rm = (size_t)B * w.nt_master;
    const size_t rtb = rows * (size_t)w.nt_track_b2;   // backward work items (the finer of the two tilings)
    w.header = c.take<int>(64);
    w.track_tab = c.take<RowTab>(rows);
    w.track_tab_b = (kTrackBwdL == kTrackFwdL) ? w.track_tab : c.take<RowTab>(rows);
    w.master_tab = c.take<RowTab>(B);
    w.master_tab_b = (kMasterBwdL == kMasterL) ? w.master_tab : c.take<RowTab>(B);
    w.y = c.take<float>(rows * w.Tp);
    w.bus_pre = c.take<float>((size_t)B * 2 * w.Tp);
    w.dbus = c.take<float>((size_t)B * 2 * w.Tp);
    w.track_etab = fwd_only ? nullptr : c.take<EqBwdTab>(rows);
    w.esave = fwd_only ? nullptr : c.take<float>(rows * w.Tp);
    w.gmid = fwd_only ? nullptr : c.take<float>(rt * (kTrackTile / kTrackBwd2Tile - 1));
    w.ssave = classic_track_bwd ? c.take<float>(rows * w.nt_track * (size_t)(kNumSections * 2) * (kTrackTile / kBwdChunk)) : nullptr;
    w.m_esave = fwd_only ? nullptr : c.take<float>((size_t)B * 2 * w.Tp);
    w.m_ssave = fwd_only ? nullptr : c.take<float>((size_t)B * w.nt_master * (size_t)(kNumSections * 2 * 2) * (kMasterTile / kMasterBwdL));
    w.flags_begin = (c.off + 255) & ~size_t(255);   // zeroed before every forward
    w.t_flag = c.take<int>(rt);
    w.m_flag = c.take<int>(rm);
    w.t_done = c.take<int>((size_t)B * w.nt_track);
    w.t_state = c.take<Mail>(rt * kStateStride);
    w.m_state = c.take<Mail>(rm * kStateStride);
    w.flags_end = c.off;
    w.bflags_begin = (c.off + 255) & ~size_t(255);  // zeroed before every backward
    w.t_bflag = c.take<int>(rtb);
    w.m_bflag = c.take<int>(rm);
    w.t_bstate = c.take<Mail>(rtb * kStateStride);
    w.m_bstate = c.take<Mail>(rm * kStateStride);
    w.bflags_end = c.off;
    w.t_tail2 = c.take<float>(rt * kTail2Stride);
    w.m_tail2 = c.take<float>(rm * kTail2Stride);
    w.t_etail = c.take<float>(rt * (size_t)la_t + 4);
    w.m_etail = c.take<float>(rm * 2 * (size_t)la_m + 4);
    w.t_dhead = c.take<float>(rtb * (size_t)la_t + 4);
    w.m_dhead = c.take<float>(rm * 2 * (size_t)la_m + 4);
    w.t_partial = c.take<float>(rtb * kGradCount);
    w.m_partial = c.take<float>(rm * kGradCount);
    w.total = (c.off + 255) & ~size_t(255);
    return w;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline unsigned track_chain_flags(unsigned f) {
    unsigned c = 0;
    if (f & DMST_USE_TRACK_INPUT_FADER) c |= kChainGain;
    if (!(f & DMST_BASIC_CONSOLE)) {
        if (f & DMST_USE_TRACK_EQ) c |= kChainEq;
        if (f & DMST_USE_TRACK_COMPRESSOR) c |= kChainComp;
    }
    return c;
}
inline unsigned master_chain_flags(unsigned f) {
    unsigned c = 0;
    if (f & DMST_BASIC_CONSOLE) return 0;
    if (f & DMST_USE_MASTER_BUS) c |= kChainGain | kChainEq | kChainComp;
    if (f & DMST_USE_OUTPUT_FADER) c |= kChainOutGain;
    return c;
}

inline size_t fwd_smem_bytes(int la_t, int la_m) {
    // delay line(s) of either role + the prefetch buffer of a track tile
    return (size_t)(fwd_ebuf_floats(kTrackTile, la_t, kMasterTile, la_m) + pidx4(kTrackTile)) * 4;
}
static_assert(kTrackTile % kMasterTile == 0 && kMasterNT == kTrackFwdNT, "forward kernel runs both roles in one CTA shape");

// CTAs of a persistent kernel: as many as are co-resident on the device
template <class Kernel>
inline int persistent_ctas(Kernel kern, int threads, size_t smem) {
#ifdef DMST_EMULATE
    (void)kern; (void)threads; (void)smem;
    return 2;
#else
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess) return 0;
    return sms * per_sm;
#endif
}
// Launch `kern` as a programmatic dependent of the kernel launched just before it on `stream`: it may start as soon
// as every CTA of that kernel has executed griddep_launch_dependents() (or exited) and runs beside it.  Everything
// launched before that kernel has completed; what that kernel itself produces must be consumed through flags.
template <class Kernel, class Arg>
inline void launch_dependent(Kernel kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const Arg& arg) {
#ifdef DMST_EMULATE
    DMST_LAUNCH(kern, grid, block, smem, stream, arg);
#else
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, arg);
#endif
}
inline size_t bwd_smem_bytes(int nch, int tile, int la, int nt, bool master) {
    // per-tile buffer area (+ the forward lane carry-ins of a chain without checkpoints)
    return (size_t)(bwd_area_floats(nch, tile, la) + (master ? 12 * nch * nt : 0)) * 4;
}

struct ConsoleCall {
    const float* tracks; long long tbs, trs;
    const float* track_params; const float* master_params;
    const dmst_ranges* ranges; float sr;
    int B, N, T; unsigned flags; int la_t, la_m;
};

inline int check_call(const ConsoleCall& k) {
    if (!k.tracks || !k.track_params || !k.ranges) return DMST_EINVAL;
    if (k.B <= 0 || k.N <= 0 || k.T <= 0) return DMST_EINVAL;
    if (k.flags & DMST_USE_FX_BUS) return DMST_EINVAL;              // out of scope
    if (!(k.flags & DMST_USE_TRACK_PANNER)) return DMST_EINVAL;     // broken upstream (modules.py:269)
    if (!(k.flags & DMST_BASIC_CONSOLE) && !k.master_params) return DMST_EINVAL;
    // (limits: the backward kernels keep two look-ahead + tile lines and the prefetched gradient in shared memory)
    if (k.la_t < 0 || k.la_t > kTrackBwd2Tile || k.la_m < 0 || k.la_m > kMasterTile) return DMST_EINVAL;
    if ((k.la_t & 31) || (k.la_m & 31)) return DMST_EINVAL;  // look-ahead must be a multiple of 32 samples
    return 0;
}

inline void fill_prepare(PrepareArgs& p, const ConsoleCall& k, bool master, ConsoleWs& w, int* status) {
    memset(&p, 0, sizeof(p));
    const bool basic = (k.flags & DMST_BASIC_CONSOLE) != 0;
    if (!master) {
        p.params = k.track_params; p.rows = k.B * k.N;
        p.np = basic ? 2 : DMST_NUM_TRACK_PARAMS; p.kind = basic ? 1 : 0;
        if (basic) {
            p.lo[0] = k.ranges->track_lo[0]; p.hi[0] = k.ranges->track_hi[0];
            p.lo[1] = k.ranges->track_lo[25]; p.hi[1] = k.ranges->track_hi[25];
        } else {
            for (int i = 0; i < DMST_NUM_TRACK_PARAMS; ++i) { p.lo[i] = k.ranges->track_lo[i]; p.hi[i] = k.ranges->track_hi[i]; }
        }
        p.L[0] = kTrackFwdL; p.tab[0] = w.track_tab;
        if (w.track_tab_b != w.track_tab) { p.L[1] = kTrackBwdL; p.tab[1] = w.track_tab_b; }
        p.status_base = 0;
    } else {
        p.params = k.master_params; p.rows = k.B;
        p.np = DMST_NUM_MASTER_PARAMS; p.kind = 2;
        for (int i = 0; i < DMST_NUM_MASTER_PARAMS; ++i) { p.lo[i] = k.ranges->master_lo[i]; p.hi[i] = k.ranges->master_hi[i]; }
        p.L[0] = kMasterL; p.tab[0] = w.master_tab; p.status_base = 1000;
        if (w.master_tab_b != w.master_tab) { p.L[1] = kMasterBwdL; p.tab[1] = w.master_tab_b; }
    }
    p.sr = (double)k.sr; p.status = status;
}

inline unsigned debug_flags() {
    static const unsigned f = (getenv("DMST_DEBUG_NOWAIT") && getenv("DMST_DEBUG_NOWAIT")[0] == '1') ? kChainDebugNoWait : 0u;
    return f;
}

inline void fill_chain(ChainArgs& a, const ConsoleCall& k, bool master, ConsoleWs& w) {
    memset(&a, 0, sizeof(a));
    a.N = k.N; a.T = k.T; a.Tp = w.Tp;
    if (!master) {
        a.nrows = k.B * k.N; a.ntiles = w.nt_track; a.flags = track_chain_flags(k.flags) | debug_flags();
        a.lookahead = k.la_t;
        a.src = k.tracks; a.src_batch_stride = k.tbs; a.src_row_stride = k.trs;
        a.src_vec_ok = aligned16(k.tracks) && (k.tbs % 4 == 0) && (k.trs % 4 == 0);
        a.tab = w.track_tab; a.track_tab = w.track_tab; a.y = w.y;
        if (a.flags & kChainEq) { a.esave = w.esave; a.ssave = w.ssave; }
        if (a.flags & kChainComp) { a.gmid = w.gmid; a.gmid_shift = kTrackBwd2Shift; }
        a.ticket = w.header + 0; a.flag = w.t_flag; a.state = w.t_state; a.tail2 = w.t_tail2; a.etail = w.t_etail;
        a.partial = w.t_partial; a.bflag = w.t_bflag; a.bstate = w.t_bstate; a.dhead = w.t_dhead;
    } else {
        a.nrows = k.B; a.ntiles = w.nt_master; a.flags = master_chain_flags(k.flags) | debug_flags();
        a.lookahead = k.la_m;
        a.src = w.y; a.tab = w.master_tab; a.track_tab = w.track_tab; a.bus_pre = w.bus_pre;
        if (a.flags & kChainEq) { a.esave = w.m_esave; a.ssave = w.m_ssave; }  // checkpoints: backward skips the EQ recompute
        a.ticket = w.header + 1; a.flag = w.m_flag; a.state = w.m_state; a.tail2 = w.m_tail2; a.etail = w.m_etail;
        a.partial = w.m_partial; a.bflag = w.m_bflag; a.bstate = w.m_bstate; a.dhead = w.m_dhead;
    }
}

#define DMST_CHECK(expr)            \
    do {                            \
        int _e = (expr);            \
        if (_e != 0) return _e;     \
    } while (0)

inline int console_forward(const ConsoleCall& k, float* mix, float* mixed, int* status, void* ws,
                           size_t ws_bytes, cudaStream_t stream) {
    DMST_CHECK(check_call(k));
    if (!mix || !status || !ws) return DMST_EINVAL;
    if ((k.flags & DMST_WANT_MIXED_TRACKS) && !mixed) return DMST_EINVAL;
    ConsoleWs w = carve_console(ws, k.B, k.N, k.T, k.la_t, k.la_m, k.flags);
    if (ws_bytes < w.total || !aligned16(ws)) return DMST_EINVAL;
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    DMST_CHECK(DMST_MEMSET_ASYNC(status, 0x7f, 4, stream));  // 0x7f7f7f7f = "no offender yet"

    PrepareArgs pt, pm;
    fill_prepare(pt, k, false, w, status);
    // the track prepare kernel also clears the tickets and the chain flags / mailboxes of forward AND backward
    // (contiguous in the workspace): no memset nodes in front of the chain kernels
    pt.zero[0] = reinterpret_cast<int4*>(w.header); pt.zero_n16[0] = 64 * sizeof(int) / 16;
    const size_t zend = w.bflags_end > w.flags_end ? w.bflags_end : w.flags_end;
    pt.zero[1] = reinterpret_cast<int4*>(base + w.flags_begin); pt.zero_n16[1] = (long long)((zend - w.flags_begin + 15) / 16);
    fill_prepare(pm, k, true, w, status);
    if (!k.master_params) { pm.kind = 3; pm.np = 0; }
    DMST_LAUNCH(prepare2_kernel, dim3(pt.rows + pm.rows), dim3(256), 0, stream, pt, pm);   // tracks and master bus, one launch

    FwdArgs f;
    memset(&f, 0, sizeof(f));
    fill_chain(f.t, k, false, w);
    f.t.want_mixed = (k.flags & DMST_WANT_MIXED_TRACKS) ? 1 : 0;
    f.t.mixed = mixed;
    f.t.user_vec_ok = mixed && aligned16(mixed) && (k.T % 4 == 0);
    fill_chain(f.m, k, true, w);
    f.m.mix = mix;
    f.m.user_vec_ok = aligned16(mix) && (k.T % 4 == 0);
    f.B = k.B; f.R = kTrackTile / kMasterTile;
    f.group = k.B * f.R + k.B * k.N;
    f.ticket = w.header + 0; f.done = w.t_done;
    {
        auto kern = console_fwd_kernel<kTrackFwdL, kMasterL, kTrackFwdNT, kBwdChunk, kMasterBwdL>;
        const size_t smem = fwd_smem_bytes(k.la_t, k.la_m);
        DMST_CHECK(DMST_SET_SMEM(kern, smem));
        int ctas = persistent_ctas(kern, kTrackFwdNT, smem);
        if (ctas <= 0) return DMST_EINVAL;
        // a master tile claimed `lag` groups after the track tiles it sums normally finds them finished
        // (measured optimum on B200: one more group than the CTAs in flight span)
        f.lag = (ctas + f.group - 1) / f.group;
        if (f.lag < 1) f.lag = 1;
        if (const char* e = getenv("DMST_FWD_LAG")) f.lag = atoi(e) > 0 ? atoi(e) : f.lag;  // tuning aid
        f.total = (w.nt_track + f.lag) * f.group;
        if (ctas > f.total) ctas = f.total;
        ScopedTimer tm(0, stream);
        DMST_LAUNCH(kern, dim3(ctas), dim3(kTrackFwdNT), smem, stream, f);
    }
    return DMST_LAST_ERROR();
}

inline int console_backward(const ConsoleCall& k, const float* gmix, const float* gmixed, float* gtp,
                            float* gmp, float* gtracks, void* ws, size_t ws_bytes, cudaStream_t stream) {
    DMST_CHECK(check_call(k));
    if (!gmix || !gtp || !ws) return DMST_EINVAL;
    if (k.flags & DMST_FORWARD_ONLY) return DMST_EINVAL;   // the forward call kept no checkpoints
    if ((k.flags & DMST_WANT_GRAD_TRACKS) && !gtracks) return DMST_EINVAL;
    ConsoleWs w = carve_console(ws, k.B, k.N, k.T, k.la_t, k.la_m, k.flags);
    if (ws_bytes < w.total || !aligned16(ws)) return DMST_EINVAL;
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    BwdArgs fm, ft;
    memset(&fm, 0, sizeof(fm));
    memset(&ft, 0, sizeof(ft));
    ChainArgs& am = fm.a;
    ChainArgs& at = ft.a;
    fill_chain(am, k, true, w);
    am.src = w.bus_pre; am.tab = w.master_tab_b;
    am.gout = gmix; am.gsrc = w.dbus;
    am.user_vec_ok = aligned16(gmix) && (k.T % 4 == 0);
    fm.total = am.nrows * am.ntiles; fm.ticket = w.header + 2;
    fm.area = bwd_area_floats(2, kMasterTile, k.la_m);
    fill_chain(at, k, false, w);
    at.gout = w.dbus; at.gmixed = gmixed;
    at.gsrc = (k.flags & DMST_WANT_GRAD_TRACKS) ? gtracks : nullptr;
    at.user_vec_ok = (k.T % 4 == 0) && (!gmixed || aligned16(gmixed)) && (!at.gsrc || aligned16(at.gsrc));
    at.tab = w.track_tab_b;
    ft.total = at.nrows * at.ntiles; ft.ticket = w.header + 3;
    ft.area = bwd_area_floats(1, kTrackTile, k.la_t);
    const bool params_only = !(k.flags & DMST_WANT_GRAD_TRACKS);
    // tickets and reverse-chain flags / mailboxes: cleared by the recursion-table kernel when it runs (first kernel
    // of the parameter-gradient path), else by memsets
    if (!(params_only && (at.flags & kChainEq))) {
        DMST_CHECK(DMST_MEMSET_ASYNC(w.header, 0, 64 * sizeof(int), stream));
        DMST_CHECK(DMST_MEMSET_ASYNC(base + w.bflags_begin, 0, w.bflags_end - w.bflags_begin, stream));
    }
    // Parameter gradients only (training): the track kernel is launched as a programmatic dependent of the master
    // kernel.  The master chain is latency bound (B chains of tiles, each hop a round trip through L2): it gets a
    // fraction of the SMs, the track kernel starts on the others at once and follows the master's per-tile flags.
    static const bool overlap_env = !(getenv("DMST_BWD_OVERLAP") && getenv("DMST_BWD_OVERLAP")[0] == '0');   // tuning aid
    const bool overlap = overlap_env && !profiler().enabled;   // (per-kernel timing brackets each kernel with events: serial)
    if (params_only && (at.flags & kChainEq)) {   // (before the master kernel: nothing may sit between it and its dependent)
        PrepareBwdArgs pb;
        memset(&pb, 0, sizeof(pb));
        pb.params = k.track_params; pb.rows = at.nrows; pb.np = DMST_NUM_TRACK_PARAMS; pb.sr = (double)k.sr; pb.tab = w.track_etab;
        for (int i = 0; i < DMST_NUM_TRACK_PARAMS; ++i) { pb.lo[i] = k.ranges->track_lo[i]; pb.hi[i] = k.ranges->track_hi[i]; }
        pb.zero[0] = reinterpret_cast<int4*>(w.header); pb.zero_n16[0] = 64 * sizeof(int) / 16;
        pb.zero[1] = reinterpret_cast<int4*>(base + w.bflags_begin); pb.zero_n16[1] = (long long)((w.bflags_end - w.bflags_begin + 15) / 16);
        DMST_LAUNCH(prepare_bwd_kernel, dim3(pb.rows), dim3(kNumRec * 32), 0, stream, pb);
    }
    {
        auto kern = chain_bwd_kernel<2, kMasterBwdL, kMasterBwdNT, true>;
        const size_t smem = bwd_smem_bytes(2, kMasterTile, k.la_m, kMasterBwdNT, true);
        DMST_CHECK(DMST_SET_SMEM(kern, smem));
        const int max_ctas = persistent_ctas(kern, kMasterBwdNT, smem);
        if (max_ctas <= 0) return DMST_EINVAL;
        int ctas = max_ctas;
        if (params_only && overlap) {
            ctas = master_bwd_overlap_ctas(max_ctas, k.B, k.N);   // (one CTA per SM: max_ctas = SMs)
            if (ctas > max_ctas) ctas = max_ctas;
        }
        if (const char* e = getenv("DMST_MASTER_BWD_CTAS")) { const int c = atoi(e); if (c > 0) ctas = c < max_ctas ? c : max_ctas; }  // tuning aid
        if (ctas > fm.total) ctas = fm.total;
        ScopedTimer tm(2, stream);
        DMST_LAUNCH(kern, dim3(ctas), dim3(kMasterBwdNT), smem, stream, fm);
    }
    int epilogue_tiles = w.nt_track;
    if (params_only) {
        // commuting-sections formulation, needs only the EQ-output checkpoint
        Bwd2Args f2;
        memset(&f2, 0, sizeof(f2));
        f2.b = ft;
        f2.b.a.ntiles = w.nt_track_b2;
        f2.b.total = at.nrows * w.nt_track_b2;
        f2.etab = w.track_etab;
        f2.fwd_ntiles = w.nt_track;
        f2.fwd_ratio = kTrackTile / kTrackBwd2Tile;
        f2.mflag = w.m_bflag; f2.m_ntiles = w.nt_master; f2.m_tile_shift = kMasterTileShift;
        epilogue_tiles = w.nt_track_b2;
        auto kern = track_bwd2_kernel<kTrackBwd2L, kTrackBwd2NT>;
        const size_t smem = bwd_smem_bytes(1, kTrackBwd2Tile, k.la_t, kTrackBwd2NT, false);
        DMST_CHECK(DMST_SET_SMEM(kern, smem));
        int ctas = persistent_ctas(kern, kTrackBwd2NT, smem);
        if (ctas <= 0) return DMST_EINVAL;
        if (ctas > f2.b.total) ctas = f2.b.total;
        ScopedTimer tm(3, stream);
        if (overlap) launch_dependent(kern, dim3(ctas), dim3(kTrackBwd2NT), smem, stream, f2);
        else DMST_LAUNCH(kern, dim3(ctas), dim3(kTrackBwd2NT), smem, stream, f2);
    } else {
        auto kern = chain_bwd_kernel<1, kTrackBwdL, kTrackBwdNT, false>;
        const size_t smem = bwd_smem_bytes(1, kTrackTile, k.la_t, kTrackBwdNT, false);
        DMST_CHECK(DMST_SET_SMEM(kern, smem));
        int ctas = persistent_ctas(kern, kTrackBwdNT, smem);
        if (ctas <= 0) return DMST_EINVAL;
        if (ctas > ft.total) ctas = ft.total;
        ScopedTimer tm(3, stream);
        DMST_LAUNCH(kern, dim3(ctas), dim3(kTrackBwdNT), smem, stream, ft);
    }
    {
        EpilogueArgs em, e;
        memset(&em, 0, sizeof(em));
        memset(&e, 0, sizeof(e));
        const bool with_master = gmp && k.master_params;
        if (with_master) {
            em.params = k.master_params; em.rows = k.B; em.np = DMST_NUM_MASTER_PARAMS; em.kind = 2;
            for (int i = 0; i < DMST_NUM_MASTER_PARAMS; ++i) { em.lo[i] = k.ranges->master_lo[i]; em.hi[i] = k.ranges->master_hi[i]; }
            em.sr = (double)k.sr; em.partial = w.m_partial; em.ntiles = w.nt_master; em.flags = am.flags; em.grad = gmp;
        }
        const bool basic = (k.flags & DMST_BASIC_CONSOLE) != 0;
        e.params = k.track_params; e.rows = k.B * k.N;
        e.np = basic ? 2 : DMST_NUM_TRACK_PARAMS; e.kind = basic ? 1 : 0;
        if (basic) {
            e.lo[0] = k.ranges->track_lo[0]; e.hi[0] = k.ranges->track_hi[0];
            e.lo[1] = k.ranges->track_lo[25]; e.hi[1] = k.ranges->track_hi[25];
        } else {
            for (int i = 0; i < DMST_NUM_TRACK_PARAMS; ++i) { e.lo[i] = k.ranges->track_lo[i]; e.hi[i] = k.ranges->track_hi[i]; }
        }
        e.sr = (double)k.sr; e.partial = w.t_partial; e.ntiles = epilogue_tiles; e.flags = at.flags; e.grad = gtp;
        // master-bus rows and track rows in one launch
        if (with_master) DMST_LAUNCH(grad_epilogue2_kernel, dim3(em.rows + e.rows), dim3(64), 0, stream, em, e);
        else DMST_LAUNCH(grad_epilogue_kernel, dim3(e.rows), dim3(64), 0, stream, e);
    }
    return DMST_LAST_ERROR();
}

}  // namespace dmst
