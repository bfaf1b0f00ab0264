// STFT building blocks shared by the MRSTFT and audio-feature losses: torch.stft semantics
// (center=True, reflect padding, window zero-padded to n_fft and centred, onesided,
// unnormalised; auraloss STFTLoss.stft and mst/loss.py:106-112 both call it this way).
//
//   frames[row][f][i] = xpad[row][f*hop + i] * wpad[i]     (frame_kernel)
//   X = cuFFT R2C over the frames                          (library FFT; its work area comes
//                                                           from the caller's workspace)
//   backward: dframes = cuFFT C2R(Z), Z the half-spectrum gradient prepared by the loss
//   kernels; dx[t] = sum over padded positions that read x[t] of sum_f dframes*w
//   (ola_kernel: gather form, deterministic, no atomics).
#pragma once
#include "common.cuh"

#ifndef DMST_EMULATE
#include <cufft.h>
#include <map>
#include <mutex>
#include <tuple>
#endif

namespace dmst {

__host__ __device__ __forceinline__ int reflect_index(int j, int T) {
    // index into x for padded position j - pad (torch 'reflect': no edge repeat)
    if (j < 0) j = -j;
    if (j >= T) j = 2 * T - 2 - j;
    return j;
}

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

struct FrameArgs {
    const float* x;        // rows x T (row stride)
    long long row_stride;
    int rows, T, n, hop, win, frames;
    const float* window;   // win
    float* out;            // rows x frames x n
    const float* x2;       // optional second signal (grid.z == 2)
    long long row_stride2;
    float* out2;
};

// grid: (frames, rows, 1 or 2), block: 256
__global__ void frame_kernel(FrameArgs a) {
    const int f = blockIdx.x, row = blockIdx.y;
    const bool second = blockIdx.z != 0;
    const float* x = second ? a.x2 + (long long)row * a.row_stride2 : a.x + (long long)row * a.row_stride;
    float* o = (second ? a.out2 : a.out) + ((long long)row * a.frames + f) * a.n;
    const int pad = a.n / 2, wl = (a.n - a.win) / 2;
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        const int wi = i - wl;
        const float w = (wi >= 0 && wi < a.win) ? __ldg(a.window + wi) : 0.0f;
        const int t = reflect_index(f * a.hop + i - pad, a.T);
        o[i] = w * __ldg(x + t);
    }
}

struct OlaArgs {
    const float* dframes;  // rows x frames x n
    int rows, T, n, hop, win, frames;
    const float* window;
    float* gx;             // rows x T (contiguous)
    int accumulate;        // 0: write, 1: add
    float scale;
};

__device__ __forceinline__ float ola_at(const OlaArgs& a, const float* df, int j) {
    // sum over frames covering padded position j of dframes[f][j - f*hop] * w[j - f*hop]
    const int wl = (a.n - a.win) / 2;
    int f_hi = j / a.hop;
    if (f_hi > a.frames - 1) f_hi = a.frames - 1;
    int f_lo = (j - a.n + a.hop) / a.hop;  // ceil((j - n + 1)/hop)
    if (j - a.n + 1 <= 0) f_lo = 0;
    float s = 0.0f;
    for (int f = f_lo; f <= f_hi; ++f) {
        const int i = j - f * a.hop;
        if (i < 0 || i >= a.n) continue;
        const int wi = i - wl;
        if (wi < 0 || wi >= a.win) continue;
        s = fmaf(__ldg(df + (long long)f * a.n + i), __ldg(a.window + wi), s);
    }
    return s;
}

// grid: (ceil(T/256), rows), block 256
__global__ void ola_kernel(OlaArgs a) {
    const int row = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.T) return;
    const float* df = a.dframes + (long long)row * a.frames * a.n;
    const int pad = a.n / 2;
    float s = ola_at(a, df, t + pad);
    if (t >= 1 && t <= pad) s += ola_at(a, df, pad - t);                         // left reflection
    const int jr = pad + 2 * a.T - 2 - t;                                         // right reflection
    if (t <= a.T - 2 && jr >= pad + a.T && jr < a.T + 2 * pad) s += ola_at(a, df, jr);
    float* g = a.gx + (long long)row * a.T + t;
    s *= a.scale;
    *g = a.accumulate ? (*g + s) : s;
}

#ifndef DMST_EMULATE
// Host-side cache of cuFFT plans (handles only; device work areas are supplied per call).
struct FftPlan {
    cufftHandle handle;
    size_t work_bytes;
};
inline int get_plan(int n, int batch, bool inverse, FftPlan* out) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int>, FftPlan> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(dev, n, batch, inverse ? 1 : 0);
    auto it = cache.find(key);
    if (it == cache.end()) {
        FftPlan p;
        if (cufftCreate(&p.handle) != CUFFT_SUCCESS) return 1000;
        if (cufftSetAutoAllocation(p.handle, 0) != CUFFT_SUCCESS) return 1001;
        int nn[1] = {n};
        cufftResult r = cufftMakePlanMany(p.handle, 1, nn, nullptr, 1, 0, nullptr, 1, 0,
                                          inverse ? CUFFT_C2R : CUFFT_R2C, batch, &p.work_bytes);
        if (r != CUFFT_SUCCESS) return 1002 + (int)r;
        it = cache.emplace(key, p).first;
    }
    *out = it->second;
    return 0;
}
inline size_t plan_work_bytes(int n, int batch) {
    FftPlan a, b;
    if (get_plan(n, batch, false, &a) != 0 || get_plan(n, batch, true, &b) != 0) return (size_t)-1;
    return a.work_bytes > b.work_bytes ? a.work_bytes : b.work_bytes;
}
inline int exec_r2c(int n, int batch, float* in, float2* out, void* work, cudaStream_t s) {
    FftPlan p;
    int e = get_plan(n, batch, false, &p);
    if (e) return e;
    if (cufftSetStream(p.handle, s) != CUFFT_SUCCESS) return 1100;
    if (cufftSetWorkArea(p.handle, work) != CUFFT_SUCCESS) return 1101;
    cufftResult r = cufftExecR2C(p.handle, in, reinterpret_cast<cufftComplex*>(out));
    return r == CUFFT_SUCCESS ? 0 : 1200 + (int)r;
}
inline int exec_c2r(int n, int batch, float2* in, float* out, void* work, cudaStream_t s) {
    FftPlan p;
    int e = get_plan(n, batch, true, &p);
    if (e) return e;
    if (cufftSetStream(p.handle, s) != CUFFT_SUCCESS) return 1100;
    if (cufftSetWorkArea(p.handle, work) != CUFFT_SUCCESS) return 1101;
    cufftResult r = cufftExecC2R(p.handle, reinterpret_cast<cufftComplex*>(in), out);
    return r == CUFFT_SUCCESS ? 0 : 1200 + (int)r;
}
#endif

}  // namespace dmst
