// batch_stereo_peak_normalize (mst/utils.py:14-29): per batch item, divide both channels by
// max(|x|) over (2, T), clamped at 1e-8.  Two kernels: block maxima, then scale.
#pragma once
#include "../../include/diffmst_b200.h"
#include "common.cuh"

namespace dmst {

// grid (1, B): one block scans the item (T is at most a few hundred thousand samples)
__global__ void peak_scale_kernel(const float* x, long long bs, long long cs, float* y, int T) {
    DMST_SHARED_ARRAY(float, sh, 33);
    const int b = blockIdx.y;
    const float* L = x + (long long)b * bs;
    const float* R = L + cs;
    // NaN-propagating maximum, as torch.max / torch.clamp in the reference: one NaN sample makes the whole item NaN
    // (mst/system.py:251-253 relies on that to raise "Found nan in ref_mix"); fmaxf would drop it
    auto nmax = [](float m, float a) { return (a > m || a != a) ? a : m; };
    float m = 0.0f;
    for (int t = threadIdx.x; t < T; t += blockDim.x) m = nmax(nmax(m, fabsf(__ldg(L + t))), fabsf(__ldg(R + t)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = nmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = nmax(m, sh[w]);
        sh[32] = m < 1e-8f ? 1e-8f : m;
    }
    __syncthreads();
    const float g = sh[32];
    float* yl = y + (long long)b * 2 * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        yl[t] = __ldg(L + t) / g;
        yl[T + t] = __ldg(R + t) / g;
    }
}

inline int peak_normalize(const float* x, long long bs, long long cs, float* y, int B, int T, cudaStream_t stream) {
    if (!x || !y || B <= 0 || T <= 0) return DMST_EINVAL;
    DMST_LAUNCH(peak_scale_kernel, dim3(1, B), dim3(1024), 0, stream, x, bs, cs, y, T);
#ifdef DMST_EMULATE
    return 0;
#else
    return (int)cudaGetLastError();
#endif
}

}  // namespace dmst
