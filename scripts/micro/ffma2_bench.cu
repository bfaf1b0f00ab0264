// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 x0 = make_float2(threadIdx.x, 1.f), x1 = make_float2(2.f, 3.f), x2 = make_float2(4.f, 5.f), x3 = make_float2(6.f, 7.f);
    float2 x4 = make_float2(8.f, 1.f), x5 = make_float2(2.5f, 3.f), x6 = make_float2(4.5f, 5.f), x7 = make_float2(6.5f, 7.f);
    float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                x0.x = fmaf(x0.x, a, b); x0.y = fmaf(x0.y, a, b); x1.x = fmaf(x1.x, a, b); x1.y = fmaf(x1.y, a, b);
                x2.x = fmaf(x2.x, a, b); x2.y = fmaf(x2.y, a, b); x3.x = fmaf(x3.x, a, b); x3.y = fmaf(x3.y, a, b);
                x4.x = fmaf(x4.x, a, b); x4.y = fmaf(x4.y, a, b); x5.x = fmaf(x5.x, a, b); x5.y = fmaf(x5.y, a, b);
                x6.x = fmaf(x6.x, a, b); x6.y = fmaf(x6.y, a, b); x7.x = fmaf(x7.x, a, b); x7.y = fmaf(x7.y, a, b);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                x0 = __ffma2_rn(x0, aa, bb); x1 = __ffma2_rn(x1, aa, bb); x2 = __ffma2_rn(x2, aa, bb); x3 = __ffma2_rn(x3, aa, bb);
                x4 = __ffma2_rn(x4, aa, bb); x5 = __ffma2_rn(x5, aa, bb); x6 = __ffma2_rn(x6, aa, bb); x7 = __ffma2_rn(x7, aa, bb);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y + x4.x + x4.y + x5.x + x5.y + x6.x + x6.y + x7.x + x7.y;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * iters * 8 * 16;
            printf("mode %d (%s): %.3f ms  %.2f TFMA/s  (%.1f TFLOP/s)\n", mode, mode ? "FFMA2" : "FFMA", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        }
    }
    printf("err %d\n", (int)cudaGetLastError());
    return 0;
}
