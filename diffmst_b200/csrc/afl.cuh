// Audio feature loss (mst/loss.py:127-260): rms, crest factor, stereo width, stereo imbalance
// and mid/side bark spectrum of input vs target (B,2,T), each an MSE times its weight, in the
// reference's key order.  One streaming pass per signal produces every time-domain statistic
// (two-stage deterministic reduction); the bark spectrum is STFT 32768/8192 (stft.cuh) ->
// frame-mean |X| -> 24-band filterbank (passed in, built once on the host from
// mst/filter.py:107-161's definition) -> log.  Backward (gradient w.r.t. input) reuses the
// statistics and spectra kept in the workspace.
#pragma once
#include "../../include/diffmst_b200.h"
#include "mrstft.cuh"

namespace dmst {

constexpr int kAflChunk = 8192;   // samples per stats block
constexpr int kAflStats = 8;      // sumL2, sumR2, sumS2, sumD2, maxL, maxR, argmaxL, argmaxR

struct AflStatsArgs {
    const float* x;  // (B,2,T)
    long long batch_stride, ch_stride;
    int B, T, chunks;
    float* partial;  // [B][chunks][kAflStats]
};

// grid (chunks, B)
__global__ void afl_stats_kernel(AflStatsArgs a) {
    DMST_SHARED_ARRAY(float, sh, 32);
    DMST_SHARED_ARRAY(float, shm, 64);
    DMST_SHARED_ARRAY(int, shi, 64);
    const int b = blockIdx.y;
    const float* L = a.x + (long long)b * a.batch_stride;
    const float* R = L + a.ch_stride;
    const int begin = blockIdx.x * kAflChunk, end = min(begin + kAflChunk, a.T);
    float sl = 0.f, sr = 0.f, ss = 0.f, sd = 0.f, ml = -1.f, mr = -1.f;
    int il = 0, ir = 0;
    for (int t = begin + threadIdx.x; t < end; t += blockDim.x) {
        const float l = __ldg(L + t), r = __ldg(R + t);
        sl = fmaf(l, l, sl); sr = fmaf(r, r, sr);
        const float s = l + r, d = l - r;
        ss = fmaf(s, s, ss); sd = fmaf(d, d, sd);
        if (fabsf(l) > ml) { ml = fabsf(l); il = t; }
        if (fabsf(r) > mr) { mr = fabsf(r); ir = t; }
    }
    float* out = a.partial + ((long long)b * a.chunks + blockIdx.x) * kAflStats;
    float v;
    v = block_sum(sl, sh); if (threadIdx.x == 0) out[0] = v;
    v = block_sum(sr, sh); if (threadIdx.x == 0) out[1] = v;
    v = block_sum(ss, sh); if (threadIdx.x == 0) out[2] = v;
    v = block_sum(sd, sh); if (threadIdx.x == 0) out[3] = v;
    // arg-max (first occurrence): warp then block
    for (int ch = 0; ch < 2; ++ch) {
        float m = ch ? mr : ml; int idx = ch ? ir : il;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
            if (m2 > m || (m2 == m && i2 < idx)) { m = m2; idx = i2; }
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { shm[warp] = m; shi[warp] = idx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
                if (shm[w] > m || (shm[w] == m && shi[w] < idx)) { m = shm[w]; idx = shi[w]; }
            out[4 + ch] = fmaxf(m, 0.0f);
            out[6 + ch] = __int_as_float(idx);
        }
        __syncthreads();
    }
}

// per item: [0..1] ms (mean square) L,R; [2] mean S2; [3] mean D2; [4..5] peak L,R; [6..7] argmax L,R
struct AflReduceArgs {
    const float* partial; int B, T, chunks;
    float* stats;  // [B][kAflStats]
};
__global__ void afl_reduce_kernel(AflReduceArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    double s[4] = {0, 0, 0, 0};
    float m[2] = {-1.f, -1.f}; int idx[2] = {0, 0};
    for (int c = 0; c < a.chunks; ++c) {
        const float* p = a.partial + ((long long)b * a.chunks + c) * kAflStats;
        for (int j = 0; j < 4; ++j) s[j] += p[j];
        for (int ch = 0; ch < 2; ++ch)
            if (p[4 + ch] > m[ch]) { m[ch] = p[4 + ch]; idx[ch] = __float_as_int(p[6 + ch]); }
    }
    float* o = a.stats + (long long)b * kAflStats;
    for (int j = 0; j < 4; ++j) o[j] = (float)(s[j] / a.T);
    o[4] = m[0]; o[5] = m[1]; o[6] = __int_as_float(idx[0]); o[7] = __int_as_float(idx[1]);
}

// mid/side framing: rows 2b = L+R, 2b+1 = L-R
struct MsFrameArgs {
    const float* x; long long batch_stride, ch_stride;
    int B, T, n, hop, frames;
    const float* window;
    float* out;  // (2B) x frames x n
};
__global__ void ms_frame_kernel(MsFrameArgs a) {
    const int f = blockIdx.x, row = blockIdx.y;
    const int b = row >> 1, side = row & 1;
    const float* L = a.x + (long long)b * a.batch_stride;
    const float* R = L + a.ch_stride;
    float* o = a.out + ((long long)row * a.frames + f) * a.n;
    const int pad = a.n / 2;
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        const int t = reflect_index(f * a.hop + i - pad, a.T);
        const float l = __ldg(L + t), r = __ldg(R + t);
        o[i] = __ldg(a.window + i) * (side ? (l - r) : (l + r));
    }
}

// frame-mean magnitude: m[row][bin] = mean_f |X[row][f][bin]|
struct MagMeanArgs { const float2* X; int rows, frames, bins; float* m; };
__global__ void mag_mean_kernel(MagMeanArgs a) {
    const int row = blockIdx.y, bin = blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= a.bins) return;
    float s = 0.0f;
    for (int f = 0; f < a.frames; ++f) {
        const float2 x = a.X[((long long)row * a.frames + f) * a.bins + bin];
        s += sqrtf(fmaf(x.x, x.x, x.y * x.y));
    }
    a.m[(long long)row * a.bins + bin] = s / a.frames;
}

// bark[row][band] = log(sum_bin fb[band][bin] * m[row][bin] + 1e-8); also z = the sum
struct BarkArgs { const float* m; const float* fb; int rows, bins, bands; float* z; };
__global__ void bark_kernel(BarkArgs a) {  // grid (bands, rows), block 256
    DMST_SHARED_ARRAY(float, sh, 32);
    const int band = blockIdx.x, row = blockIdx.y;
    float s = 0.0f;
    for (int k = threadIdx.x; k < a.bins; k += blockDim.x)
        s = fmaf(__ldg(a.fb + (long long)band * a.bins + k), a.m[(long long)row * a.bins + k], s);
    const float r = block_sum(s, sh);
    if (threadIdx.x == 0) a.z[row * a.bands + band] = r;
}

// Losses from the per-item statistics and the bark sums; also the per-item coefficients the
// backward pass needs.  Single thread: B is small and everything here is O(B * bands).
struct AflFinalArgs {
    const float* sx; const float* sy;  // stats of input / target [B][8]
    const float* zx; const float* zy;  // bark sums [(2B)][bands]
    int B, bands; float w[5];
    float* losses;   // [5]
};
__device__ __forceinline__ void afl_features(const float* s, float* rms, float* cf, float* sw, float* si) {
    for (int c = 0; c < 2; ++c) {
        rms[c] = sqrtf(fmaxf(s[c], 1e-8f));
        // __fmul_rn: no FMA contraction with the caller's subtraction (equal inputs => exactly 0)
        cf[c] = __fmul_rn(20.0f, log10f(fmaxf(s[4 + c] / fmaxf(rms[c], 1e-8f), 1e-8f)));
    }
    *sw = s[3] / fmaxf(s[2], 1e-8f);
    *si = (s[1] - s[0]) / fmaxf(s[1] + s[0], 1e-8f);
}
__global__ void afl_final_kernel(AflFinalArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double l_rms = 0, l_cf = 0, l_sw = 0, l_si = 0, l_bark = 0;
    for (int b = 0; b < a.B; ++b) {
        float rx[2], ry[2], cx[2], cy[2], swx, swy, six, siy;
        afl_features(a.sx + b * kAflStats, rx, cx, &swx, &six);
        afl_features(a.sy + b * kAflStats, ry, cy, &swy, &siy);
        for (int c = 0; c < 2; ++c) {
            l_rms += (double)(rx[c] - ry[c]) * (rx[c] - ry[c]);
            l_cf += (double)(cx[c] - cy[c]) * (cx[c] - cy[c]);
        }
        l_sw += (double)(swx - swy) * (swx - swy);
        l_si += (double)(six - siy) * (six - siy);
        for (int r = 0; r < 2; ++r)
            for (int k = 0; k < a.bands; ++k) {
                const float d = logf(a.zx[(2 * b + r) * a.bands + k] + 1e-8f) - logf(a.zy[(2 * b + r) * a.bands + k] + 1e-8f);
                l_bark += (double)d * d;
            }
    }
    a.losses[0] = (float)(a.w[0] * l_rms / (2.0 * a.B));
    a.losses[1] = (float)(a.w[1] * l_cf / (2.0 * a.B));
    a.losses[2] = (float)(a.w[2] * l_sw / a.B);
    a.losses[3] = (float)(a.w[3] * l_si / a.B);
    a.losses[4] = (float)(a.w[4] * l_bark / (2.0 * a.B * a.bands));
}

// ---- backward ----
// d(bark loss)/d m[row][bin] = sum_band fb[band][bin] * dz[row][band],
// dz = gw4 * w4 * 2 (lbx - lby) / (2 B bands) / (zx + 1e-8); then dX = dm/frames * X/|X|.
struct BarkGradArgs {
    float2* X;  // in spectrum of input (2B rows); out Z
    const float* zx; const float* zy; const float* fb; const float* gw;  // gw: upstream [5] or null
    int B, frames, bins, bands; float w4;
};
__global__ void bark_grad_kernel(BarkGradArgs a) {  // grid (ceil(bins/256), 2B)
    const int row = blockIdx.y, bin = blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= a.bins) return;
    const float up = a.gw ? a.gw[4] : 1.0f;
    float dm = 0.0f;
    for (int k = 0; k < a.bands; ++k) {
        const float zx = a.zx[row * a.bands + k], zy = a.zy[row * a.bands + k];
        const float d = logf(zx + 1e-8f) - logf(zy + 1e-8f);
        const float dz = up * a.w4 * d / ((float)a.B * a.bands) / (zx + 1e-8f);
        dm = fmaf(__ldg(a.fb + (long long)k * a.bins + bin), dz, dm);
    }
    dm /= a.frames;
    for (int f = 0; f < a.frames; ++f) {
        const long long idx = ((long long)row * a.frames + f) * a.bins + bin;
        const float2 x = a.X[idx];
        const float mag = sqrtf(fmaf(x.x, x.x, x.y * x.y));
        float2 z = make_float2(0.f, 0.f);
        if (mag > 0.0f) { const float s = dm / mag; z.x = s * x.x; z.y = s * x.y; }
        if (bin == 0 || bin == a.bins - 1) z.y = 0.0f; else { z.x *= 0.5f; z.y *= 0.5f; }
        a.X[idx] = z;
    }
}

// time-domain gradient: g_c[t] = a_c x_c[t] + b_c x_other[t] + peak term + (g_mid +/- g_side)
struct AflGradArgs {
    const float* x; long long batch_stride, ch_stride;
    const float* sx; const float* sy; const float* gw; float w[5];
    const float* gms;   // (2B) x T  gradient w.r.t. mid/side signals (bark path), may be null
    int B, T;
    float* gx;          // (B,2,T)
};
__global__ void afl_grad_kernel(AflGradArgs a) {  // grid (ceil(T/256), B)
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.T) return;
    const float* s = a.sx + b * kAflStats;
    float rx[2], ry[2], cx[2], cy[2], swx, swy, six, siy;
    afl_features(s, rx, cx, &swx, &six);
    afl_features(a.sy + b * kAflStats, ry, cy, &swy, &siy);
    const float u0 = (a.gw ? a.gw[0] : 1.f) * a.w[0], u1 = (a.gw ? a.gw[1] : 1.f) * a.w[1];
    const float u2 = (a.gw ? a.gw[2] : 1.f) * a.w[2], u3 = (a.gw ? a.gw[3] : 1.f) * a.w[3];
    const float* L = a.x + (long long)b * a.batch_stride;
    const float* R = L + a.ch_stride;
    const float xv[2] = {__ldg(L + t), __ldg(R + t)};
    const float invT = 1.0f / a.T;
    const float k20 = 8.685889638065035f;  // 20/ln(10)
    float g[2] = {0.f, 0.f};
    for (int c = 0; c < 2; ++c) {
        const bool rms_live = s[c] > 1e-8f;       // clamp(min) inside the sqrt
        const float drms_dx = rms_live ? xv[c] * invT / rx[c] : 0.0f;
        // rms loss: w0 * mean_{b,c} (r - r')^2
        g[c] += u0 * (rx[c] - ry[c]) / a.B * drms_dx;
        // crest factor: cf = k20 * ln(peak / max(rms, 1e-8)), clamp on the ratio at 1e-8
        const float ratio = s[4 + c] / fmaxf(rx[c], 1e-8f);
        if (ratio > 1e-8f) {
            const float dcf = u1 * (cx[c] - cy[c]) / a.B;
            if (rx[c] > 1e-8f) g[c] += dcf * (-k20 / rx[c]) * drms_dx;
            if (t == __float_as_int(s[6 + c]) && s[4 + c] > 0.0f)
                g[c] += dcf * (k20 / s[4 + c]) * (xv[c] > 0.f ? 1.f : (xv[c] < 0.f ? -1.f : 0.f));
        }
    }
    const float sum = xv[0] + xv[1], dif = xv[0] - xv[1];
    {   // stereo width D/S
        const float dsw = u2 * 2.0f * (swx - swy) / a.B;
        const float S = fmaxf(s[2], 1e-8f);
        const float dD = dsw / S, dS = (s[2] > 1e-8f) ? -dsw * s[3] / (S * S) : 0.0f;
        g[0] += dD * 2.f * dif * invT + dS * 2.f * sum * invT;
        g[1] += -dD * 2.f * dif * invT + dS * 2.f * sum * invT;
    }
    {   // stereo imbalance (ER - EL)/Q
        const float dsi = u3 * 2.0f * (six - siy) / a.B;
        const float Q = fmaxf(s[0] + s[1], 1e-8f);
        const bool live = (s[0] + s[1]) > 1e-8f;
        const float dEL = live ? dsi * (-2.0f * s[1]) / (Q * Q) : -dsi / Q;
        const float dER = live ? dsi * (2.0f * s[0]) / (Q * Q) : dsi / Q;
        g[0] += dEL * 2.f * xv[0] * invT;
        g[1] += dER * 2.f * xv[1] * invT;
    }
    if (a.gms) {
        const float gm = a.gms[(long long)(2 * b) * a.T + t], gs = a.gms[(long long)(2 * b + 1) * a.T + t];
        g[0] += gm + gs;
        g[1] += gm - gs;
    }
    a.gx[((long long)b * 2 + 0) * a.T + t] = g[0];
    a.gx[((long long)b * 2 + 1) * a.T + t] = g[1];
}

#ifndef DMST_EMULATE
struct AflWs {
    float *partial, *stats_x, *stats_y, *frames, *m, *zx, *zy, *gms;
    float2 *spec_x, *spec_y;
    void* fft_work;
    int chunks, frames_n;
    size_t total;
};
inline int afl_carve(void* base, int B, int T, int n, int bands, AflWs* w) {
    if (B <= 0 || T <= 0 || n <= 0 || (n & 1) || n / 2 >= T) return DMST_EINVAL;
    const int hop = n / 4;
    const int frames = 1 + T / hop, bins = n / 2 + 1, rows = 2 * B;
    unsigned char* b = reinterpret_cast<unsigned char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 255) & ~size_t(255); void* p = b ? b + off : nullptr; off += bytes; return p; };
    w->chunks = (T + kAflChunk - 1) / kAflChunk;
    w->frames_n = frames;
    w->partial = (float*)take((size_t)B * w->chunks * kAflStats * 4);
    w->stats_x = (float*)take((size_t)B * kAflStats * 4);
    w->stats_y = (float*)take((size_t)B * kAflStats * 4);
    w->frames = (float*)take((size_t)rows * frames * n * 4);
    w->spec_x = (float2*)take((size_t)rows * frames * bins * 8);
    w->spec_y = (float2*)take((size_t)rows * frames * bins * 8);
    w->m = (float*)take((size_t)rows * bins * 4);
    w->zx = (float*)take((size_t)rows * bands * 4);
    w->zy = (float*)take((size_t)rows * bands * 4);
    w->gms = (float*)take((size_t)rows * T * 4);
    const size_t wk = plan_work_bytes(n, rows * frames);
    if (wk == (size_t)-1) return 1002;
    w->fft_work = take(wk);
    w->total = (off + 255) & ~size_t(255);
    return 0;
}

inline int afl_forward(const float* input, const float* target, long long bs, long long cs, const float* fb,
                       const float* window, const float* weights, int B, int T, int n, int bands,
                       float* losses, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!input || !target || !fb || !window || !weights || !losses || !ws) return DMST_EINVAL;
    AflWs w;
    int e = afl_carve(ws, B, T, n, bands, &w);
    if (e) return e;
    if (ws_bytes < w.total) return DMST_EINVAL;
    const int hop = n / 4, frames = w.frames_n, bins = n / 2 + 1, rows = 2 * B;
    for (int pass = 0; pass < 2; ++pass) {
        const float* sig = pass ? target : input;
        float* stats = pass ? w.stats_y : w.stats_x;
        float2* spec = pass ? w.spec_y : w.spec_x;
        float* z = pass ? w.zy : w.zx;
        AflStatsArgs sa{sig, bs, cs, B, T, w.chunks, w.partial};
        afl_stats_kernel<<<dim3(w.chunks, B), 256, 0, stream>>>(sa);
        AflReduceArgs ra{w.partial, B, T, w.chunks, stats};
        afl_reduce_kernel<<<(B + 63) / 64, 64, 0, stream>>>(ra);
        MsFrameArgs fa{sig, bs, cs, B, T, n, hop, frames, window, w.frames};
        ms_frame_kernel<<<dim3(frames, rows), 256, 0, stream>>>(fa);
        e = exec_r2c(n, rows * frames, w.frames, spec, w.fft_work, stream);
        if (e) return e;
        MagMeanArgs ma{spec, rows, frames, bins, w.m};
        mag_mean_kernel<<<dim3((bins + 255) / 256, rows), 256, 0, stream>>>(ma);
        BarkArgs ba{w.m, fb, rows, bins, bands, z};
        bark_kernel<<<dim3(bands, rows), 256, 0, stream>>>(ba);
    }
    AflFinalArgs fa{w.stats_x, w.stats_y, w.zx, w.zy, B, bands, {weights[0], weights[1], weights[2], weights[3], weights[4]}, losses};
    afl_final_kernel<<<1, 32, 0, stream>>>(fa);
    return (int)cudaGetLastError();
}

inline int afl_backward(const float* input, long long bs, long long cs, const float* fb, const float* window,
                        const float* weights, const float* gw, int B, int T, int n, int bands, float* grad_input,
                        void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!input || !fb || !window || !weights || !grad_input || !ws) return DMST_EINVAL;
    AflWs w;
    int e = afl_carve(ws, B, T, n, bands, &w);
    if (e) return e;
    if (ws_bytes < w.total) return DMST_EINVAL;
    const int hop = n / 4, frames = w.frames_n, bins = n / 2 + 1, rows = 2 * B;
    BarkGradArgs bg{w.spec_x, w.zx, w.zy, fb, gw, B, frames, bins, bands, weights[4]};
    bark_grad_kernel<<<dim3((bins + 255) / 256, rows), 256, 0, stream>>>(bg);
    e = exec_c2r(n, rows * frames, w.spec_x, w.frames, w.fft_work, stream);
    if (e) return e;
    OlaArgs oa{w.frames, rows, T, n, hop, n, frames, window, w.gms, 0, 1.0f};
    ola_kernel<<<dim3((T + 255) / 256, rows), 256, 0, stream>>>(oa);
    AflGradArgs ga{input, bs, cs, w.stats_x, w.stats_y, gw, {weights[0], weights[1], weights[2], weights[3], weights[4]},
                   w.gms, B, T, grad_input};
    afl_grad_kernel<<<dim3((T + 255) / 256, B), 256, 0, stream>>>(ga);
    return (int)cudaGetLastError();
}
#endif

}  // namespace dmst
