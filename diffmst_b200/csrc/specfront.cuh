// Spectrogram front-end of the Cnn14 encoder (mst/modules.py:771-806): waveform -> STFT (torch.stft
// semantics: centred, reflect-padded, Hann window of n_fft, onesided) -> (|X| + eps)^power, written
// straight into the zero-bordered NHWC tensor the first convolution consumes (bins = height,
// frames = width, input channel = channel).  Three launches: the 128-bit framing kernel of the MRSTFT
// loss, one batched cuFFT R2C, and spec_pow_kernel below (magnitude, compression and the
// (frame, bin) -> (bin, frame) transposition through a shared-memory tile); the reference runs
// pad + stft + abs + add + pow, and the convolution trunk would need a layout conversion after them.
#pragma once
#include "mrstft.cuh"

#ifndef DMST_EMULATE
namespace dmst {

struct SpecPowArgs {
    const float2* X;   // rows x frames x bins
    float* out;        // (B, bins + 2, frames + 2, C), border zeroed beforehand
    int rows, C, frames, bins;
    float eps, power;
};

// grid: (ceil(bins/32), ceil(frames/32), rows), block (32, 8)
__global__ void spec_pow_kernel(SpecPowArgs a) {
    __shared__ float tile[32][33];
    const int row = blockIdx.z, b = row / a.C, c = row - b * a.C;
    const int bin0 = blockIdx.x * 32, fr0 = blockIdx.y * 32;
    const float2* X = a.X + (long long)row * a.frames * a.bins;
#pragma unroll
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int f = fr0 + j, k = bin0 + threadIdx.x;
        float v = 0.0f;
        if (f < a.frames && k < a.bins) {
            const float2 z = __ldg(X + (long long)f * a.bins + k);
            // torch.abs of a complex tensor is hypot(re, im); torch.pow is powf
            v = powf(hypotf(z.x, z.y) + a.eps, a.power);
        }
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    const int Wp = a.frames + 2;
#pragma unroll
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int k = bin0 + j, f = fr0 + threadIdx.x;
        if (k < a.bins && f < a.frames)
            a.out[(((long long)b * (a.bins + 2) + k + 1) * Wp + f + 1) * a.C + c] = tile[threadIdx.x][j];
    }
}

struct SpecWs {
    float* frames;     // rows x frames x n
    float2* spec;      // rows x frames x bins
    void* fft_work;
    size_t total;
};
inline int spec_carve(void* base, int rows, int T, int n, int hop, SpecWs* w) {
    if (rows <= 0 || T <= 0 || n < 8 || (n & (n - 1)) || hop <= 0 || T <= n / 2) return DMST_EINVAL;
    const int frames = 1 + T / hop, bins = n / 2 + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    const size_t o_frames = take((size_t)rows * frames * n * sizeof(float));
    const size_t o_spec = take((size_t)rows * frames * bins * sizeof(float2));
    const size_t wb = plan_work_bytes(n, rows * frames);
    if (wb == (size_t)-1) return 1003;
    const size_t o_work = take(wb);
    char* p = reinterpret_cast<char*>(base);
    w->frames = reinterpret_cast<float*>(p + o_frames);
    w->spec = reinterpret_cast<float2*>(p + o_spec);
    w->fft_work = p + o_work;
    w->total = off;
    return 0;
}

inline int spectrogram_frontend(const float* x, long long row_stride, const float* window, int B, int C, int T, int n,
                                int hop, float eps, float power, float* out_padded, void* ws, size_t ws_bytes,
                                cudaStream_t s) {
    if (!x || !window || !out_padded || !ws || B <= 0 || C <= 0) return DMST_EINVAL;
    const int rows = B * C;
    SpecWs w;
    int e = spec_carve(ws, rows, T, n, hop, &w);
    if (e) return e;
    if (ws_bytes < w.total) return DMST_EINVAL;
    const int frames = 1 + T / hop, bins = n / 2 + 1;
    Frame4Args fa;
    fa.x[0] = x; fa.x[1] = x; fa.row_stride[0] = row_stride; fa.row_stride[1] = row_stride;
    fa.vec_ok[0] = fa.vec_ok[1] = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (row_stride % 4 == 0) && (hop % 4 == 0);
    fa.out[0] = w.frames; fa.out[1] = w.frames;
    fa.rows = rows; fa.T = T; fa.n = n; fa.hop = hop; fa.win = n; fa.frames = frames; fa.window = window;
    frame4_kernel<<<dim3((frames * (n / 4) + 255) / 256, rows, 1), 256, 0, s>>>(fa);
    e = exec_r2c(n, rows * frames, w.frames, w.spec, w.fft_work, s);
    if (e) return e;
    if (cudaMemsetAsync(out_padded, 0, (size_t)B * (bins + 2) * (frames + 2) * C * sizeof(float), s) != cudaSuccess)
        return (int)cudaGetLastError();
    SpecPowArgs pa{w.spec, out_padded, rows, C, frames, bins, eps, power};
    spec_pow_kernel<<<dim3((bins + 31) / 32, (frames + 31) / 32, rows), dim3(32, 8), 0, s>>>(pa);
    return (int)cudaGetLastError();
}

}  // namespace dmst
#endif
