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

namespace dmst {

constexpr int kMrBlock = 256;
constexpr int kMrItemsPerBlock = 4096;  // spectrum elements per block

struct MrLossArgs {
    const float2* X;  // rows x frames x bins
    const float2* Y;
    int rows, per_row;   // per_row = frames * bins
    int blocks_per_row;
    float eps;
    float* partial;      // [rows][blocks_per_row][4]: sum (|Y|-|X|)^2, sum |Y|^2, sum |log|, sum |lin|
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.0f;
    if (warp == 0) {
        r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;  // valid in warp 0
}

// grid: (blocks_per_row, rows)
__global__ void mr_loss_kernel(MrLossArgs a) {
    DMST_SHARED_ARRAY(float, sh, 32);
    const int row = blockIdx.y;
    const long long base = (long long)row * a.per_row;
    const int begin = blockIdx.x * kMrItemsPerBlock;
    const int end = min(begin + kMrItemsPerBlock, a.per_row);
    float s_d2 = 0.f, s_y2 = 0.f, s_log = 0.f, s_lin = 0.f;
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const float2 x = a.X[base + i], y = a.Y[base + i];
        const float px = fmaxf(fmaf(x.x, x.x, x.y * x.y), a.eps), py = fmaxf(fmaf(y.x, y.x, y.y * y.y), a.eps);
        const float mx = sqrtf(px), my = sqrtf(py);
        const float d = my - mx;
        s_d2 = fmaf(d, d, s_d2);
        s_y2 += py;
        s_log += fabsf(0.5f * (logf(px) - logf(py)));
        s_lin += fabsf(d);
    }
    float* out = a.partial + ((long long)row * a.blocks_per_row + blockIdx.x) * 4;
    float r;
    r = block_sum(s_d2, sh); if (threadIdx.x == 0) out[0] = r;
    r = block_sum(s_y2, sh); if (threadIdx.x == 0) out[1] = r;
    r = block_sum(s_log, sh); if (threadIdx.x == 0) out[2] = r;
    r = block_sum(s_lin, sh); if (threadIdx.x == 0) out[3] = r;
}

// Second stage of the reductions: one block per row sums that row's block partials in float64
// (fixed assignment of partials to threads + fixed-shape tree => deterministic).
struct MrRowSumArgs { const float* partial; int blocks_per_row; double* rowsum; /* [rows][4] */ };
__global__ void mr_rowsum_kernel(MrRowSumArgs a) {
    DMST_SHARED_ARRAY(double, sh, 4 * 128);
    const int row = blockIdx.x, tid = threadIdx.x;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = tid; b < a.blocks_per_row; b += blockDim.x) {
        const float4 p = *reinterpret_cast<const float4*>(a.partial + ((long long)row * a.blocks_per_row + b) * 4);
        s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += p.w;
    }
    for (int j = 0; j < 4; ++j) sh[j * 128 + tid] = s[j];
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o)
            for (int j = 0; j < 4; ++j) sh[j * 128 + tid] += sh[j * 128 + tid + o];
        __syncthreads();
    }
    if (tid < 4) a.rowsum[row * 4 + tid] = sh[tid * 128];
}

struct MrFinalArgs {
    const double* rowsum;
    int rows, per_row;
    float w_sc, w_log, w_lin;
    int n_res, res_index;
    float* loss;        // [0] total (accumulated over resolutions), [1 + 3*r ...] sc, log, lin
    float* row_coef;    // [rows]: d(total)/d|X| coefficient of (|X|-|Y|) for the SC term
    float* scal;        // [2]: coefficient of sign(log) / |X| and of sign(lin)
};

__global__ void mr_final_kernel(MrFinalArgs a) {
    if (threadIdx.x != 0) return;
    double sc = 0.0, slog = 0.0, slin = 0.0;
    for (int row = 0; row < a.rows; ++row) {
        const double d2 = a.rowsum[row * 4 + 0], y2 = a.rowsum[row * 4 + 1];
        slog += a.rowsum[row * 4 + 2]; slin += a.rowsum[row * 4 + 3];
        const double num = sqrt(d2), den = sqrt(y2);
        sc += num / den;
        // d/d|X| of w_sc * (1/rows) * ||Y|-|X||_F / ||Y||_F = w_sc/(rows) * (|X|-|Y|) / (num*den)
        a.row_coef[row] = (num > 0.0) ? (float)(a.w_sc / (a.rows * (double)a.n_res * num * den)) : 0.0f;
    }
    const double cnt = (double)a.rows * (double)a.per_row;
    const double l_sc = sc / a.rows, l_log = slog / cnt, l_lin = slin / cnt;
    const double lr = a.w_sc * l_sc + a.w_log * l_log + a.w_lin * l_lin;
    a.loss[1 + 3 * a.res_index + 0] = (float)l_sc;
    a.loss[1 + 3 * a.res_index + 1] = (float)l_log;
    a.loss[1 + 3 * a.res_index + 2] = (float)l_lin;
    const float prev = (a.res_index == 0) ? 0.0f : a.loss[0];
    a.loss[0] = prev + (float)(lr / a.n_res);
    a.scal[0] = (float)(a.w_log / (a.n_res * cnt));
    a.scal[1] = (float)(a.w_lin / (a.n_res * cnt));
}

struct MrGradArgs {
    float2* X;        // in: spectrum of x; out: Z, the C2R-ready half-spectrum gradient
    const float2* Y;
    int rows, frames, bins;
    float eps;
    const float* row_coef;
    const float* scal;
    int use_log, use_lin;
};

// grid: (ceil(per_row/256), rows)
__global__ void mr_grad_kernel(MrGradArgs a) {
    const int row = blockIdx.y;
    const int per_row = a.frames * a.bins;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_row) return;
    const long long idx = (long long)row * per_row + i;
    const int bin = i % a.bins;
    const float2 x = a.X[idx], y = a.Y[idx];
    const float px_raw = fmaf(x.x, x.x, x.y * x.y);
    float2 z = make_float2(0.0f, 0.0f);
    if (px_raw >= a.eps) {  // clamp passes gradient only where it is inactive
        const float py = fmaxf(fmaf(y.x, y.x, y.y * y.y), a.eps);
        const float mx = sqrtf(px_raw), my = sqrtf(py);
        float g = a.row_coef[row] * (mx - my);
        if (a.use_log) {
            const float dl = logf(px_raw) - logf(py);
            g += a.scal[0] * ((dl > 0.0f) - (dl < 0.0f)) / mx;
        }
        if (a.use_lin) g += a.scal[1] * ((mx > my) - (mx < my));
        const float s = g / mx;
        z.x = s * x.x; z.y = s * x.y;
    }
    // adjoint of the onesided real FFT expressed through an unnormalised C2R transform:
    // DC and Nyquist keep their real part, interior bins are halved
    if (bin == 0 || bin == a.bins - 1) z.y = 0.0f;
    else { z.x *= 0.5f; z.y *= 0.5f; }
    a.X[idx] = z;
}

#ifndef DMST_EMULATE
struct MrWs {
    float* frames;    // 2*rows*frames*n
    float2* spec;     // 2*rows*frames*bins
    float* partial; double* rowsum; float* row_coef; float* scal; void* fft_work;
    size_t total;
};
inline int mr_max_dims(const dmst_mrstft_cfg* c, int rows, int T, size_t* fr, size_t* sp, size_t* part, size_t* work) {
    *fr = *sp = *part = *work = 0;
    for (int r = 0; r < c->n_res; ++r) {
        const int n = c->fft_size[r], hop = c->hop_size[r], win = c->win_length[r];
        if (n <= 0 || hop <= 0 || win <= 0 || win > n || (n & 1) || n / 2 >= T) return DMST_EINVAL;
        const size_t frames = 1 + T / hop, bins = n / 2 + 1;
        *fr = max(*fr, (size_t)2 * rows * frames * n);
        *sp = max(*sp, (size_t)2 * rows * frames * bins);
        const size_t bpr = (frames * bins + kMrItemsPerBlock - 1) / kMrItemsPerBlock;
        *part = max(*part, (size_t)rows * bpr * 4);
        const size_t w1 = plan_work_bytes(n, 2 * rows * (int)frames), w2 = plan_work_bytes(n, rows * (int)frames);
        if (w1 == (size_t)-1 || w2 == (size_t)-1) return 1002;
        *work = max(*work, max(w1, w2));
    }
    return 0;
}
inline int mr_carve(void* base, const dmst_mrstft_cfg* c, int rows, int T, MrWs* w) {
    size_t fr, sp, part, work;
    int e = mr_max_dims(c, rows, T, &fr, &sp, &part, &work);
    if (e) return e;
    unsigned char* b = reinterpret_cast<unsigned char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 255) & ~size_t(255); void* p = b ? b + off : nullptr; off += bytes; return p; };
    w->frames = (float*)take(fr * 4);
    w->spec = (float2*)take(sp * 8);
    w->partial = (float*)take(part * 4);
    w->rowsum = (double*)take((size_t)rows * 4 * 8);
    w->row_coef = (float*)take((size_t)rows * 4);
    w->scal = (float*)take(16);
    w->fft_work = take(work);
    w->total = (off + 255) & ~size_t(255);
    return 0;
}

inline int mrstft_run(const float* x, long long xs, const float* y, long long ys, const float* windows,
                      const dmst_mrstft_cfg* c, int rows, int T, float* loss, float* grad_x, void* ws,
                      size_t ws_bytes, cudaStream_t stream) {
    if (!x || !y || !windows || !c || !loss || !ws || rows <= 0 || T <= 0) return DMST_EINVAL;
    if (c->n_res <= 0 || c->n_res > DMST_MRSTFT_MAX_RES) return DMST_EINVAL;
    MrWs w;
    int e = mr_carve(ws, c, rows, T, &w);
    if (e) return e;
    if (ws_bytes < w.total) return DMST_EINVAL;
    const float* win = windows;
    for (int r = 0; r < c->n_res; ++r) {
        const int n = c->fft_size[r], hop = c->hop_size[r], wl = c->win_length[r];
        const int frames = 1 + T / hop, bins = n / 2 + 1;
        const int per_row = frames * bins;
        float* fx = w.frames;
        float* fy = w.frames + (size_t)rows * frames * n;
        FrameArgs fa{x, xs, rows, T, n, hop, wl, frames, win, fx, y, ys, fy};
        frame_kernel<<<dim3(frames, rows, 2), 256, 0, stream>>>(fa);
        e = exec_r2c(n, 2 * rows * frames, w.frames, w.spec, w.fft_work, stream);
        if (e) return e;
        float2* X = w.spec;
        float2* Y = w.spec + (size_t)rows * per_row;
        const int bpr = (per_row + kMrItemsPerBlock - 1) / kMrItemsPerBlock;
        MrLossArgs la{X, Y, rows, per_row, bpr, c->eps, w.partial};
        mr_loss_kernel<<<dim3(bpr, rows), kMrBlock, 0, stream>>>(la);
        MrRowSumArgs rs{w.partial, bpr, w.rowsum};
        mr_rowsum_kernel<<<rows, 128, 0, stream>>>(rs);
        MrFinalArgs fa2{w.rowsum, rows, per_row, c->w_sc, c->w_log_mag, c->w_lin_mag, c->n_res, r,
                        loss, w.row_coef, w.scal};
        mr_final_kernel<<<1, 32, 0, stream>>>(fa2);
        if (grad_x) {
            MrGradArgs ga{X, Y, rows, frames, bins, c->eps, w.row_coef, w.scal, c->w_log_mag != 0.0f,
                          c->w_lin_mag != 0.0f};
            mr_grad_kernel<<<dim3((per_row + 255) / 256, rows), 256, 0, stream>>>(ga);
            e = exec_c2r(n, rows * frames, X, fx, w.fft_work, stream);
            if (e) return e;
            OlaArgs oa{fx, rows, T, n, hop, wl, frames, win, grad_x, r > 0 ? 1 : 0, 1.0f};
            ola_kernel<<<dim3((T + 255) / 256, rows), 256, 0, stream>>>(oa);
        }
        win += wl;
    }
    return (int)cudaGetLastError();
}
#endif

}  // namespace dmst
