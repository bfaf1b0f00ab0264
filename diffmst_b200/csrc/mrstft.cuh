// Multi-resolution STFT loss, forward and gradient w.r.t. the first argument.
// Replaces auraloss.freq.MultiResolutionSTFTLoss.forward (instantiated
// at configs/models/naive.yaml:54-68 and mst/system.py:61-69, called at mst/system.py:332):
//   |X| = sqrt(clamp(re^2 + im^2, eps));
//   L_sc = mean_rows ||Y|-|X||_F / ||Y||_F;  L_log = mean |log|X| - log|Y||;  L_lin = mean ||X|-|Y||
//   loss = mean_res (w_sc L_sc + w_log L_log + w_lin L_lin)
// The spectra come from cuFFT (stft.cuh); one fused kernel reads X and Y once and produces all
// reductions (two-stage, fixed order => deterministic); a second fused kernel turns them into
// the half-spectrum gradient that the C2R + overlap-add adjoint consumes.
#pragma once
#include "../../include/diffmst_b200.h"
#include "stft.cuh"
#include "stft_fused.cuh"
#ifndef DMST_EMULATE
#include <vector>
#endif

namespace dmst {

#ifndef DMST_EMULATE   // (cuFFT-fed: not part of the host-emulated build)
constexpr int kMrBlock = 256;
constexpr int kMrLoadsInFlight = 4;      // independent 8-byte loads per thread and operand
// Spectrum elements per block of the loss kernel: the (blocks_per_row, rows) grid is sized to ONE wave of
// 148 SMs x 8 resident blocks (a second, partial wave cost 25 % of the kernel at the headline shape)
inline int mr_items_per_block(int per_row, int rows) {
    constexpr int kStep = kMrBlock * kMrLoadsInFlight;
    int bpr = (148 * 8) / (rows > 0 ? rows : 1);
    if (bpr < 1) bpr = 1;
    int items = (per_row + bpr - 1) / bpr;
    items = ((items + kStep - 1) / kStep) * kStep;
    return items < 4 * kStep ? 4 * kStep : items;
}

// ---------------------------------------------------------------------------------
// Framing with 128-bit accesses: thread g handles samples [4*i4, 4*i4+4) of frame f of one row.
// grid: (ceil(frames*n/4 / 256), rows, 2 signals)
// ---------------------------------------------------------------------------------
struct Frame4Args {
    const float* x[2];        // the two signals, rows x T each
    long long row_stride[2];
    int vec_ok[2];            // 16-byte aligned base and row stride % 4 == 0
    float* out[2];            // rows x frames x n
    int rows, T, n, hop, win, frames;
    const float* window;      // win
};
__global__ void frame4_kernel(Frame4Args a) {
    const int row = blockIdx.y, z = blockIdx.z;
    const int nq = a.n >> 2;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.frames * nq) return;
    const int f = g / nq, i = (g - f * nq) << 2;
    // (kernel-parameter arrays are only ever indexed with constants: a dynamic index would force a local copy)
    const float* x = (z ? a.x[1] : a.x[0]) + (long long)row * (z ? a.row_stride[1] : a.row_stride[0]);
    float* o = (z ? a.out[1] : a.out[0]) + ((long long)row * a.frames + f) * a.n + i;
    const bool vec_ok = (z ? a.vec_ok[1] : a.vec_ok[0]) != 0;
    const int pad = a.n >> 1, wl = (a.n - a.win) >> 1;
    const int t = f * a.hop + i - pad;
    float4 xv, wv;
    if (vec_ok && t >= 0 && t + 3 < a.T && (t & 3) == 0) {
        xv = __ldg(reinterpret_cast<const float4*>(x + t));
    } else {
        xv.x = __ldg(x + reflect_index(t, a.T)); xv.y = __ldg(x + reflect_index(t + 1, a.T));
        xv.z = __ldg(x + reflect_index(t + 2, a.T)); xv.w = __ldg(x + reflect_index(t + 3, a.T));
    }
    if (wl == 0) {
        wv = __ldg(reinterpret_cast<const float4*>(a.window + i));
    } else {
        float w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { const int wi = i + e - wl; w[e] = (wi >= 0 && wi < a.win) ? __ldg(a.window + wi) : 0.0f; }
        wv = make_float4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<float4*>(o) = make_float4(xv.x * wv.x, xv.y * wv.y, xv.z * wv.z, xv.w * wv.w);
}

// MUFU reciprocal square root and log2 (arguments here are clamped at eps > 0, never subnormal):
// |X| = p * rsqrt(p) and 1/|X| = rsqrt(p) are good to 2 ulp, log2 to 2^-22 absolute near 1 (2 ulp elsewhere):
// three orders of magnitude inside the 1e-4 of the loss, and these two kernels stay bound by HBM
// instead of by the issue rate (the IEEE sqrtf / logf / division sequences cost ~80 instructions per bin)
__device__ __forceinline__ float fast_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_log2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct MrLossArgs {
    const float2* X;  // rows x frames x bins
    const float2* Y;
    int rows, per_row;   // per_row = frames * bins
    int blocks_per_row, items_per_block;
    float eps;
    float* partial;      // [rows][blocks_per_row][4]: sum (|Y|-|X|)^2, sum |Y|^2, sum |log|, sum |lin|
    unsigned* done;      // blocks finished (zero at launch; the last block resets it)
    MrFinalArgs fin;
};
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid: (blocks_per_row, rows)
__global__ void mr_loss_kernel(MrLossArgs a) {
    DMST_SHARED_ARRAY(float, sh, 33);
    const int row = blockIdx.y;
    const long long base = (long long)row * a.per_row;
    const int begin = blockIdx.x * a.items_per_block;
    const int end = min(begin + a.items_per_block, a.per_row);
    float s_d2 = 0.f, s_y2 = 0.f, s_log = 0.f, s_lin = 0.f;
    constexpr int U = kMrLoadsInFlight;
    for (int i0 = begin + threadIdx.x; i0 < end; i0 += U * kMrBlock) {
        float2 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * kMrBlock;
            if (i < end) { x[u] = __ldg(a.X + base + i); y[u] = __ldg(a.Y + base + i); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * kMrBlock >= end) continue;
            const float px = fmaxf(fmaf(x[u].x, x[u].x, x[u].y * x[u].y), a.eps);
            const float py = fmaxf(fmaf(y[u].x, y[u].x, y[u].y * y[u].y), a.eps);
            // (__fmul_rn: no contraction into an FMA, so that |Y| - |X| is exactly 0 for identical spectra)
            const float d = __fmul_rn(py, fast_rsqrt(py)) - __fmul_rn(px, fast_rsqrt(px));   // |Y| - |X|
            s_d2 = fmaf(d, d, s_d2);
            s_y2 += py;
            s_log += fabsf(fast_log2(px) - fast_log2(py));
            s_lin += fabsf(d);
        }
    }
    s_log *= 0.5f * 0.6931471805599453f;   // log|X| - log|Y| = ln2 / 2 * (log2 px - log2 py)
    float* out = a.partial + ((long long)row * a.blocks_per_row + blockIdx.x) * 4;
    float r;
    r = block_sum(s_d2, sh); if (threadIdx.x == 0) out[0] = r;
    r = block_sum(s_y2, sh); if (threadIdx.x == 0) out[1] = r;
    r = block_sum(s_log, sh); if (threadIdx.x == 0) out[2] = r;
    r = block_sum(s_lin, sh); if (threadIdx.x == 0) out[3] = r;
    // the block that finishes last reduces the partials (threadFenceReduction pattern)
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned prev = atomicAdd(a.done, 1u);
        sh[32] = (prev == total - 1) ? 1.0f : 0.0f;
        if (prev == total - 1) { *a.done = 0u; __threadfence(); }   // ready for the next launch
    }
    __syncthreads();
    if (sh[32] != 0.0f) mr_final(a.fin);
}

__device__ void mr_final(const MrFinalArgs& a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int row = warp; row < a.rows; row += nwarps) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int b = lane; b < a.blocks_per_row; b += 32) {
            const float4 p = __ldcg(reinterpret_cast<const float4*>(a.partial + ((long long)row * a.blocks_per_row + b) * 4));
            s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += p.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s[j] = warp_sum_f64(s[j]);
            if (lane == 0) a.rowsum[row * 4 + j] = s[j];
        }
    }
    __syncthreads();
    if (warp != 0) return;
    // lanes over rows (fixed assignment), then a fixed shuffle tree: deterministic
    double sc = 0.0, slog = 0.0, slin = 0.0;
    for (int row = lane; row < a.rows; row += 32) {
        const double d2 = a.rowsum[row * 4 + 0], y2 = a.rowsum[row * 4 + 1];
        slog += a.rowsum[row * 4 + 2]; slin += a.rowsum[row * 4 + 3];
        const double num = sqrt(d2), den = sqrt(y2);
        sc += num / den;
        // d/d|X| of w_sc * (1/rows) * ||Y|-|X||_F / ||Y||_F = w_sc/(rows) * (|X|-|Y|) / (num*den)
        a.row_coef[row] = (num > 0.0) ? (float)(a.w_sc / (a.rows * (double)a.n_res * num * den)) : 0.0f;
    }
    sc = warp_sum_f64(sc); slog = warp_sum_f64(slog); slin = warp_sum_f64(slin);
    if (lane != 0) return;
    const double cnt = (double)a.rows * (double)a.per_row;
    const double l_sc = sc / a.rows, l_log = slog / cnt, l_lin = slin / cnt;
    const double lr = a.w_sc * l_sc + a.w_log * l_log + a.w_lin * l_lin;
    a.res_loss[0] = (float)(lr / a.n_res);
    a.res_loss[1] = (float)l_sc; a.res_loss[2] = (float)l_log; a.res_loss[3] = (float)l_lin;
    a.scal[0] = (float)(a.w_log / (a.n_res * cnt));
    a.scal[1] = (float)(a.w_lin / (a.n_res * cnt));
}

struct MrGradArgs {
    float2* X;        // in: spectrum of x; out: Z, the C2R-ready half-spectrum gradient
    const float2* Y;  // spectrum of y, or null when PY is given
    const float* PY;  // max(|Y|^2, eps) (the fused front end keeps only this of the target), or null
    int rows, frames, bins;
    float eps;
    const float* row_coef;
    const float* scal;
    int use_log, use_lin;
};

// grid: (ceil(per_row/(256*U)), rows)
constexpr int kMrGradU = 4;
__global__ void mr_grad_kernel(MrGradArgs a) {
    const int row = blockIdx.y;
    const int per_row = a.frames * a.bins;
    const long long base = (long long)row * per_row;
    const int i0 = blockIdx.x * (kMrBlock * kMrGradU) + threadIdx.x;
    const float rc = __ldg(a.row_coef + row), c_log = a.use_log ? __ldg(a.scal) : 0.0f, c_lin = a.use_lin ? __ldg(a.scal + 1) : 0.0f;
    float2 x[kMrGradU], y[kMrGradU];
#pragma unroll
    for (int u = 0; u < kMrGradU; ++u) {
        const int i = i0 + u * kMrBlock;
        if (i < per_row) {
            x[u] = a.X[base + i];
            if (a.PY) y[u] = make_float2(__ldg(a.PY + base + i), 0.0f);
            else y[u] = __ldg(a.Y + base + i);
        }
    }
    int bin = i0 % a.bins;   // one division per thread; the other elements step by the block size
#pragma unroll
    for (int u = 0; u < kMrGradU; ++u) {
        const int i = i0 + u * kMrBlock;
        if (i < per_row) {
            const float px_raw = fmaf(x[u].x, x[u].x, x[u].y * x[u].y);
            float2 z = make_float2(0.0f, 0.0f);
            if (px_raw >= a.eps) {  // clamp passes gradient only where it is inactive
                const float py = a.PY ? y[u].x : fmaxf(fmaf(y[u].x, y[u].x, y[u].y * y[u].y), a.eps);
                const float rx = fast_rsqrt(px_raw);            // 1 / |X|
                const float d = __fmul_rn(px_raw, rx) - __fmul_rn(py, fast_rsqrt(py));   // |X| - |Y| (no FMA contraction)
                // sign(log|X| - log|Y|) = sign(|X|^2 - |Y|^2) (monotone), so the gradient needs no logarithm
                const float sg_log = (px_raw > py) ? 1.0f : ((px_raw < py) ? -1.0f : 0.0f);
                const float sg_lin = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
                const float g = fmaf(rc, d, fmaf(c_log * sg_log, rx, c_lin * sg_lin));
                const float s = g * rx;
                z.x = s * x[u].x; z.y = s * x[u].y;
            }
            // adjoint of the onesided real FFT expressed through an unnormalised C2R transform:
            // DC and Nyquist keep their real part, interior bins are halved
            if (bin == 0 || bin == a.bins - 1) z.y = 0.0f;
            else { z.x *= 0.5f; z.y *= 0.5f; }
            a.X[base + i] = z;
        }
        bin += kMrBlock;
        while (bin >= a.bins) bin -= a.bins;
    }
}

// ---------------------------------------------------------------------------------
// Overlap-add adjoint of all resolutions in one pass over the gradient (written once), plus the
// total loss.  Gather form, fixed order => deterministic.  grid: (ceil(T/4 / 256), rows)
// ---------------------------------------------------------------------------------
struct OlaMultiArgs {
    int n_res, rows, T;
    const float* dframes[DMST_MRSTFT_MAX_RES];   // rows x frames x n
    const float* window[DMST_MRSTFT_MAX_RES];
    int n[DMST_MRSTFT_MAX_RES], hop[DMST_MRSTFT_MAX_RES], win[DMST_MRSTFT_MAX_RES], frames[DMST_MRSTFT_MAX_RES];
    int hop_shift[DMST_MRSTFT_MAX_RES];   // log2(hop) when hop is a power of two (the frame range needs no division), else -1
    int batch_ok[DMST_MRSTFT_MAX_RES];    // half-overlapping frames, full window, power-of-two hop: the straight-line path
    long long row_elems[DMST_MRSTFT_MAX_RES];   // frames * n
    float* gx;                 // rows x T (contiguous), or null
    int gx_vec_ok;
    const float* gscale;       // device scalar multiplying the gradient (the upstream gradient of the loss), or null
    const float* res_loss;     // [n_res][4], or null (gradient-only launch)
    float* loss;               // [0] total
    float* terms;              // [3*r ...] sc, log, lin, or null
};
__global__ void ola_multi_kernel(OlaMultiArgs a) {
    if (a.res_loss && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        float tot = 0.0f;
        for (int r = 0; r < a.n_res; ++r) {
            tot += a.res_loss[4 * r];
            if (a.terms) {
                a.terms[3 * r + 0] = a.res_loss[4 * r + 1];
                a.terms[3 * r + 1] = a.res_loss[4 * r + 2];
                a.terms[3 * r + 2] = a.res_loss[4 * r + 3];
            }
        }
        a.loss[0] = tot;
    }
    if (!a.gx) return;
    const int row = blockIdx.y;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) << 2;
    if (t >= a.T) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // Half-overlapping frames (the configurations of the reference): exactly the frames f1 = j / hop and f1 - 1 cover the
    // quad.  The frame-gradient loads of the first resolutions are all issued before the first use (one DRAM latency
    // per thread instead of one per frame).
    constexpr int kBatched = 4;
    float4 da[kBatched], db[kBatched];
    bool batched[kBatched];
#pragma unroll
    for (int r = 0; r < kBatched; ++r) {
        batched[r] = false;
        da[r] = make_float4(0.f, 0.f, 0.f, 0.f); db[r] = da[r];
        if (r < a.n_res) {
            const int hop = a.hop[r], frames = a.frames[r], hs = a.hop_shift[r];
            if (a.batch_ok[r] && (t + 3 < a.T)) {   // (host: win == n, hop a power of two >= 4, n == 2 hop, frames * n < 2^31)
                batched[r] = true;
                const float* df = a.dframes[r] + (long long)row * a.row_elems[r];
                const int j = t + hop;                          // padded position (pad = n/2 = hop)
                const int f1 = j >> hs, i1 = j & (hop - 1);     // i1 in [0, hop), a multiple of 4
                const int o1 = (f1 << (hs + 1)) + i1;           // f1 * n + i1 (32-bit: the host checked frames * n < 2^31)
                if (f1 < frames) da[r] = __ldg(reinterpret_cast<const float4*>(df + o1));
                if (f1 >= 1 && f1 - 1 < frames) db[r] = __ldg(reinterpret_cast<const float4*>(df + o1 - hop));   // (f1-1) n + i1 + hop
            }
        }
    }
#pragma unroll
    for (int r = 0; r < DMST_MRSTFT_MAX_RES; ++r) {  // fully unrolled: kernel-parameter arrays indexed with constants
        if (r >= a.n_res) break;
        const int n = a.n[r], hop = a.hop[r], pad = n >> 1, frames = a.frames[r];
        const float* df = a.dframes[r] + (long long)row * frames * n;
        OlaArgs oa{df, a.rows, a.T, n, hop, a.win[r], frames, a.window[r], nullptr, 0, 1.0f};
        const bool fast = (a.win[r] == n) && ((hop & 3) == 0) && ((n & 7) == 0) && (t + 3 < a.T);
        const int hs = a.hop_shift[r];
        if (r < kBatched && batched[r < kBatched ? r : 0]) {
            const int j = t + pad;
            const int i1 = j & (hop - 1);
            const float4 wa = __ldg(reinterpret_cast<const float4*>(a.window[r] + i1));
            const float4 wb = __ldg(reinterpret_cast<const float4*>(a.window[r] + i1 + hop));
            const float4 d0 = da[r < kBatched ? r : 0], d1 = db[r < kBatched ? r : 0];   // (zero where the frame does not exist)
            acc[0] = fmaf(d0.x, wa.x, acc[0]); acc[1] = fmaf(d0.y, wa.y, acc[1]);
            acc[2] = fmaf(d0.z, wa.z, acc[2]); acc[3] = fmaf(d0.w, wa.w, acc[3]);
            acc[0] = fmaf(d1.x, wb.x, acc[0]); acc[1] = fmaf(d1.y, wb.y, acc[1]);
            acc[2] = fmaf(d1.z, wb.z, acc[2]); acc[3] = fmaf(d1.w, wb.w, acc[3]);
        } else if (fast && hs >= 0 && n == 2 * hop) {
            // half-overlapping frames (the configurations of the reference): exactly the frames f1 = j / hop and f1 - 1
            // cover the quad, straight-line code
            const int j = t + pad;
            const int f1 = j >> hs, i1 = j - (f1 << hs);   // i1 in [0, hop), a multiple of 4
            const float* wv = a.window[r];
            if (f1 < frames) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(df + (long long)f1 * n + i1));
                const float4 w = __ldg(reinterpret_cast<const float4*>(wv + i1));
                acc[0] = fmaf(d.x, w.x, acc[0]); acc[1] = fmaf(d.y, w.y, acc[1]);
                acc[2] = fmaf(d.z, w.z, acc[2]); acc[3] = fmaf(d.w, w.w, acc[3]);
            }
            if (f1 >= 1 && f1 - 1 < frames) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(df + (long long)(f1 - 1) * n + i1 + hop));
                const float4 w = __ldg(reinterpret_cast<const float4*>(wv + i1 + hop));
                acc[0] = fmaf(d.x, w.x, acc[0]); acc[1] = fmaf(d.y, w.y, acc[1]);
                acc[2] = fmaf(d.z, w.z, acc[2]); acc[3] = fmaf(d.w, w.w, acc[3]);
            }
        } else if (fast) {
            const int j = t + pad;  // padded position of the first of the 4 samples (multiple of 4)
            int f_hi = hs >= 0 ? (j >> hs) : j / hop;
            if (f_hi > frames - 1) f_hi = frames - 1;
            int f_lo = 0;
            if (j + 3 - n + 1 > 0) f_lo = hs >= 0 ? ((j + 3 - n + hop) >> hs) : (j + 3 - n + hop) / hop;   // (numerator >= hop here)
            for (int f = f_lo; f <= f_hi; ++f) {
                const int i = j - f * hop;  // multiple of 4; the 4 samples lie in [0, n) by the choice of f_lo, f_hi
                if (i < 0 || i + 3 >= n) {  // frame covers only part of the quad (cannot happen when 4 | hop, kept for safety)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ie = i + e;
                        if (ie >= 0 && ie < n) acc[e] = fmaf(__ldg(df + (long long)f * n + ie), __ldg(a.window[r] + ie), acc[e]);
                    }
                    continue;
                }
                const float4 d = __ldg(reinterpret_cast<const float4*>(df + (long long)f * n + i));
                const float4 w = __ldg(reinterpret_cast<const float4*>(a.window[r] + i));
                acc[0] = fmaf(d.x, w.x, acc[0]); acc[1] = fmaf(d.y, w.y, acc[1]);
                acc[2] = fmaf(d.z, w.z, acc[2]); acc[3] = fmaf(d.w, w.w, acc[3]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (t + e < a.T) acc[e] += ola_at(oa, df, t + e + pad);
        }
        // reflected padding folds back onto the first / last `pad` samples
        if (t <= pad || t + 3 >= a.T - 1 - pad) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int te = t + e;
                if (te >= 1 && te <= pad && te < a.T) acc[e] += ola_at(oa, df, pad - te);
                const int jr = pad + 2 * a.T - 2 - te;
                if (te <= a.T - 2 && jr >= pad + a.T && jr < a.T + 2 * pad) acc[e] += ola_at(oa, df, jr);
            }
        }
    }
    if (a.gscale) {
        const float gs = __ldg(a.gscale);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] *= gs;
    }
    float* g = a.gx + (long long)row * a.T + t;
    if (a.gx_vec_ok && t + 3 < a.T) {
        *reinterpret_cast<float4*>(g) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (t + e < a.T) g[e] = acc[e];
    }
}

// Twiddle tables of the fused front end, built in float64 on the host once per (device, fft size)
struct SfTables { float2* tw_m; float2* tw_n; };
inline bool sf_tables(int n, SfTables* out) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, SfTables> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(dev, n);
    auto it = cache.find(key);
    if (it == cache.end()) {
        const int M = n / 2, K = M / 2 + 1;
        std::vector<float2> h(M + K);
        const double pi = 3.14159265358979323846;
        for (int j = 0; j < M; ++j) h[j] = make_float2((float)cos(2.0 * pi * j / M), (float)-sin(2.0 * pi * j / M));
        for (int k = 0; k < K; ++k) h[M + k] = make_float2((float)cos(2.0 * pi * k / n), (float)-sin(2.0 * pi * k / n));
        float2* d = nullptr;
        if (cudaMalloc(&d, sizeof(float2) * (M + K)) != cudaSuccess) return false;
        if (cudaMemcpy(d, h.data(), sizeof(float2) * (M + K), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return false; }
        it = cache.emplace(key, SfTables{d, d + M}).first;
    }
    *out = it->second;
    return true;
}
// The fused front end serves power-of-two FFT sizes (DMST_MRSTFT_FUSED=0: framing + cuFFT + loss kernels, tuning aid)
inline bool mr_use_fused(int n) {
    static const bool off = getenv("DMST_MRSTFT_FUSED") && getenv("DMST_MRSTFT_FUSED")[0] == '0';
    return !off && sf_supported(n);
}

// Per-resolution slice of the workspace: the resolutions run concurrently on side streams
struct MrResWs {
    float* frames;    // 2*rows*frames*n
    float2* spec;     // 2*rows*frames*bins
    float* partial; double* rowsum; float* row_coef; float* scal; void* fft_work;
};
struct MrWs {
    MrResWs res[DMST_MRSTFT_MAX_RES];
    float* res_loss;  // [n_res][4]
    unsigned* done;   // [MAX_RES] finished-block counters of mr_loss_kernel (zeroed once per call)
    size_t total;
};
inline int mr_carve(void* base, const dmst_mrstft_cfg* c, int rows, int T, MrWs* w) {
    unsigned char* b = reinterpret_cast<unsigned char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 255) & ~size_t(255); void* p = b ? b + off : nullptr; off += bytes; return p; };
    w->res_loss = (float*)take(sizeof(float) * 4 * DMST_MRSTFT_MAX_RES);
    w->done = (unsigned*)take(sizeof(unsigned) * DMST_MRSTFT_MAX_RES);
    for (int r = 0; r < c->n_res; ++r) {
        const int n = c->fft_size[r], hop = c->hop_size[r], win = c->win_length[r];
        if (n <= 0 || hop <= 0 || win <= 0 || win > n || (n & 7) || n / 2 >= T) return DMST_EINVAL;
        const size_t frames = 1 + T / hop, bins = n / 2 + 1;
        const size_t ipb = (size_t)mr_items_per_block((int)(frames * bins), rows);
        size_t bpr = (frames * bins + ipb - 1) / ipb;
        if (mr_use_fused(n)) {
            bpr = max(bpr, (size_t)sf_blocks_per_row(n, (int)frames));
            SfTables tb;
            if (!sf_tables(n, &tb)) return 1003;
        }
        size_t w1 = 0, w2 = 0;   // (the fused kernels need no cuFFT plan)
        if (!mr_use_fused(n)) {
            w1 = plan_work_bytes(n, 2 * rows * (int)frames); w2 = plan_work_bytes(n, rows * (int)frames);
            if (w1 == (size_t)-1 || w2 == (size_t)-1) return 1002;
        }
        MrResWs& s = w->res[r];
        s.frames = (float*)take((size_t)2 * rows * frames * n * 4);
        s.spec = (float2*)take((size_t)2 * rows * frames * bins * 8);
        s.partial = (float*)take((size_t)rows * bpr * 4 * 4);
        s.rowsum = (double*)take((size_t)rows * 4 * 8);
        s.row_coef = (float*)take((size_t)rows * 4);
        s.scal = (float*)take(16);
        s.fft_work = take(max(w1, w2));
    }
    w->total = (off + 255) & ~size_t(255);
    return 0;
}

// Side streams / events for the fork-join over resolutions (created once per device)
struct MrStreams {
    cudaStream_t side[DMST_MRSTFT_MAX_RES];
    cudaEvent_t start, done[DMST_MRSTFT_MAX_RES];
    bool ok = false;
};
inline MrStreams* mr_streams() {
    static std::mutex mu;
    static std::map<int, MrStreams> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    MrStreams& s = cache[dev];
    if (!s.ok) {
        if (cudaEventCreateWithFlags(&s.start, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int r = 0; r < DMST_MRSTFT_MAX_RES; ++r) {
            if (cudaStreamCreateWithFlags(&s.side[r], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&s.done[r], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        s.ok = true;
    }
    return &s;
}

// Launch of the overlap-add (+ loss totals) over the frame gradients the workspace holds
inline int mr_launch_ola(OlaMultiArgs& oa, const dmst_mrstft_cfg* c, int rows, int T, float* grad_x, cudaStream_t stream) {
    oa.n_res = c->n_res; oa.rows = rows; oa.T = T;
    for (int r = 0; r < DMST_MRSTFT_MAX_RES; ++r) {
        const int h = r < c->n_res ? c->hop_size[r] : 0;
        oa.hop_shift[r] = -1;
        if (h > 0 && (h & (h - 1)) == 0) { int sft = 0; while ((1 << sft) < h) ++sft; oa.hop_shift[r] = sft; }
        oa.row_elems[r] = (long long)oa.frames[r] * oa.n[r];
        oa.batch_ok[r] = r < c->n_res && oa.win[r] == oa.n[r] && oa.hop_shift[r] >= 2 && oa.n[r] == 2 * h &&
                         oa.row_elems[r] < (1ll << 31);
    }
    oa.gx = grad_x; oa.gx_vec_ok = grad_x && ((reinterpret_cast<uintptr_t>(grad_x) & 15) == 0) && (T % 4 == 0);
    const dim3 grid(grad_x ? ((T + 3) / 4 + 255) / 256 : 1, grad_x ? rows : 1);
    ola_multi_kernel<<<grid, 256, 0, stream>>>(oa);
    return (int)cudaGetLastError();
}

// keep_frames: compute the per-frame gradients and leave them in the workspace for mrstft_backward_run (the
// two-call form an autograd node uses: the upstream gradient of the loss only exists at backward time)
inline int mrstft_run(const float* x, long long xs, const float* y, long long ys, const float* windows,
                      const dmst_mrstft_cfg* c, int rows, int T, float* loss, float* terms, float* grad_x,
                      bool keep_frames, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!x || !y || !windows || !c || !loss || !ws || rows <= 0 || T <= 0) return DMST_EINVAL;
    if (c->n_res <= 0 || c->n_res > DMST_MRSTFT_MAX_RES) return DMST_EINVAL;
    MrWs w;
    int e = mr_carve(ws, c, rows, T, &w);
    if (e) return e;
    if (ws_bytes < w.total) return DMST_EINVAL;
    MrStreams* st = mr_streams();
    if (!st) return (int)cudaGetLastError();
    static const bool serial = getenv("DMST_MRSTFT_SERIAL") && getenv("DMST_MRSTFT_SERIAL")[0] == '1';  // tuning aid
    if (cudaMemsetAsync(w.done, 0, sizeof(unsigned) * DMST_MRSTFT_MAX_RES, stream) != cudaSuccess) return (int)cudaGetLastError();
    // fork: resolution 0 stays on the caller's stream, the others run on side streams
    if (c->n_res > 1 && !serial && cudaEventRecord(st->start, stream) != cudaSuccess) return (int)cudaGetLastError();
    OlaMultiArgs oa;
    memset(&oa, 0, sizeof(oa));
    const float* win = windows;
    for (int r = 0; r < c->n_res; ++r) {
        cudaStream_t s = (r == 0 || serial) ? stream : st->side[r];
        if (r > 0 && !serial && cudaStreamWaitEvent(s, st->start, 0) != cudaSuccess) return (int)cudaGetLastError();
        const int n = c->fft_size[r], hop = c->hop_size[r], wl = c->win_length[r];
        const int frames = 1 + T / hop, bins = n / 2 + 1;
        const int per_row = frames * bins;
        MrResWs& q = w.res[r];
        float* fx = q.frames;
        float* fy = q.frames + (size_t)rows * frames * n;
        float2* X = q.spec;
        float2* Y = q.spec + (size_t)rows * per_row;
        const bool fused = mr_use_fused(n);
        const bool want_spectra = grad_x || keep_frames;
        if (fused) {
            SfTables tb;
            if (!sf_tables(n, &tb)) return 1003;
            SfArgs sa;
            memset(&sa, 0, sizeof(sa));
            sa.x[0] = x; sa.x[1] = y; sa.row_stride[0] = xs; sa.row_stride[1] = ys;
            sa.vec_ok[0] = ((reinterpret_cast<uintptr_t>(x) & 7) == 0) && (xs % 2 == 0) && (hop % 2 == 0);
            sa.vec_ok[1] = ((reinterpret_cast<uintptr_t>(y) & 7) == 0) && (ys % 2 == 0) && (hop % 2 == 0);
            sa.rows = rows; sa.T = T; sa.n = n; sa.hop = hop; sa.win = wl; sa.frames = frames;
            sa.window = win; sa.win_vec_ok = (wl == n) && ((reinterpret_cast<uintptr_t>(win) & 7) == 0);
            sa.tw_m = tb.tw_m; sa.tw_n = tb.tw_n;
            sa.X = want_spectra ? X : nullptr;
            sa.PY = want_spectra ? reinterpret_cast<float*>(Y) : nullptr;   // (the target's slot of the spectrum area)
            sa.eps = c->eps; sa.partial = q.partial; sa.done = w.done + r;
            const int bpr = sf_blocks_per_row(n, frames);
            sa.fin = MrFinalArgs{q.partial, bpr, q.rowsum, rows, per_row, c->w_sc, c->w_log_mag, c->w_lin_mag, c->n_res,
                                 w.res_loss + 4 * r, q.row_coef, q.scal};
            if (!sf_launch(sa, s)) return DMST_EINVAL;
        } else {
            Frame4Args fa;
            fa.x[0] = x; fa.x[1] = y; fa.row_stride[0] = xs; fa.row_stride[1] = ys;
            fa.vec_ok[0] = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (xs % 4 == 0) && (hop % 4 == 0);
            fa.vec_ok[1] = ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && (ys % 4 == 0) && (hop % 4 == 0);
            fa.out[0] = fx; fa.out[1] = fy;
            fa.rows = rows; fa.T = T; fa.n = n; fa.hop = hop; fa.win = wl; fa.frames = frames; fa.window = win;
            frame4_kernel<<<dim3((frames * (n / 4) + 255) / 256, rows, 2), 256, 0, s>>>(fa);
            e = exec_r2c(n, 2 * rows * frames, q.frames, q.spec, q.fft_work, s);
            if (e) return e;
            const int ipb = mr_items_per_block(per_row, rows);
            const int bpr = (per_row + ipb - 1) / ipb;
            MrFinalArgs fa2{q.partial, bpr, q.rowsum, rows, per_row, c->w_sc, c->w_log_mag, c->w_lin_mag, c->n_res,
                            w.res_loss + 4 * r, q.row_coef, q.scal};
            MrLossArgs la{X, Y, rows, per_row, bpr, ipb, c->eps, q.partial, w.done + r, fa2};
            mr_loss_kernel<<<dim3(bpr, rows), kMrBlock, 0, s>>>(la);
        }
        if (want_spectra && fused) {
            // spectrum -> half-spectrum gradient -> inverse real FFT of every frame, one kernel
            SfTables tb;
            if (!sf_tables(n, &tb)) return 1003;
            SfGradArgs ga;
            memset(&ga, 0, sizeof(ga));
            ga.X = X; ga.PY = reinterpret_cast<const float*>(Y); ga.dframes = fx;
            ga.rows = rows; ga.n = n; ga.frames = frames; ga.tw_m = tb.tw_m; ga.tw_n = tb.tw_n; ga.eps = c->eps;
            ga.row_coef = q.row_coef; ga.scal = q.scal; ga.use_log = c->w_log_mag != 0.0f; ga.use_lin = c->w_lin_mag != 0.0f;
            if (!sf_grad_launch(ga, s)) return DMST_EINVAL;
        } else if (want_spectra) {
            MrGradArgs ga{X, Y, nullptr, rows, frames, bins, c->eps, q.row_coef, q.scal, c->w_log_mag != 0.0f,
                          c->w_lin_mag != 0.0f};
            mr_grad_kernel<<<dim3((per_row + kMrBlock * kMrGradU - 1) / (kMrBlock * kMrGradU), rows), kMrBlock, 0, s>>>(ga);
            e = exec_c2r(n, rows * frames, X, fx, q.fft_work, s);
            if (e) return e;
        }
        oa.dframes[r] = fx; oa.window[r] = win; oa.n[r] = n; oa.hop[r] = hop; oa.win[r] = wl; oa.frames[r] = frames;
        if (r > 0 && !serial) {
            if (cudaEventRecord(st->done[r], s) != cudaSuccess) return (int)cudaGetLastError();
            if (cudaStreamWaitEvent(stream, st->done[r], 0) != cudaSuccess) return (int)cudaGetLastError();
        }
        win += wl;
    }
    // join: overlap-add of every resolution's frame gradients + the total loss, one pass
    oa.res_loss = w.res_loss; oa.loss = loss; oa.terms = terms;
    return mr_launch_ola(oa, c, rows, T, grad_x, stream);
}

// Second call of the two-call form: grad_x = grad_loss * d loss / d x from the frame gradients mrstft_run left
inline int mrstft_backward_run(const float* windows, const dmst_mrstft_cfg* c, int rows, int T, const float* grad_loss,
                               float* grad_x, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!windows || !c || !grad_x || !ws || rows <= 0 || T <= 0) return DMST_EINVAL;
    if (c->n_res <= 0 || c->n_res > DMST_MRSTFT_MAX_RES) return DMST_EINVAL;
    MrWs w;
    int e = mr_carve(ws, c, rows, T, &w);
    if (e) return e;
    if (ws_bytes < w.total) return DMST_EINVAL;
    OlaMultiArgs oa;
    memset(&oa, 0, sizeof(oa));
    const float* win = windows;
    for (int r = 0; r < c->n_res; ++r) {
        oa.dframes[r] = w.res[r].frames; oa.window[r] = win; oa.n[r] = c->fft_size[r]; oa.hop[r] = c->hop_size[r];
        oa.win[r] = c->win_length[r]; oa.frames[r] = 1 + T / c->hop_size[r];
        win += c->win_length[r];
    }
    oa.gscale = grad_loss;
    return mr_launch_ola(oa, c, rows, T, grad_x, stream);
}
#endif

}  // namespace dmst
