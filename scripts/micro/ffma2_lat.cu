// Microbenchmark: dependent-chain latency of FFMA vs FFMA2 on sm_100a, and issue throughput of a
// 50/50 mix of FP32 work and shuffles with the FP32 half packed or not.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void lat(float* out, long long* cyc, int iters, float a, float b) {
    float2 x = make_float2(threadIdx.x, 1.f), aa = make_float2(a, a), bb = make_float2(b, b);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            if (MODE == 0) x.x = fmaf(x.x, a, b); else x = __ffma2_rn(x, aa, bb);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x.x + x.y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// biquad-like recurrence on 2 x 16 samples per thread: scalar on each half in turn vs packed halves
template <int MODE>
__global__ void biq(float* out, int iters, float b0, float b1, float b2, float na1, float na2) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            float z1 = 0.f, z2 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float x = v[i];
                const float y = fmaf(b0, x, z1);
                z1 = fmaf(b1, x, fmaf(na1, y, z2));
                z2 = fmaf(b2, x, na2 * y);
                v[i] = y;
            }
            v[0] += z1 + z2;
        } else {
            float2 z1 = make_float2(0.f, 0.f), z2 = z1;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 x = make_float2(v[i], v[i + 16]);
                const float2 y = __ffma2_rn(make_float2(b0, b0), x, z1);
                z1 = __ffma2_rn(make_float2(b1, b1), x, __ffma2_rn(make_float2(na1, na1), y, z2));
                z2 = __ffma2_rn(make_float2(b2, b2), x, __fmul2_rn(make_float2(na2, na2), y));
                v[i] = y.x; v[i + 16] = y.y;
            }
            v[0] += z1.x + z2.x + z1.y + z2.y;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    long long* cyc; cudaMalloc(&cyc, 8);
    long long h;
    lat<0><<<1, 32>>>(out, cyc, 1024, 0.999f, 0.001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FFMA  dependent latency %.2f cycles\n", (double)h / (1024 * 32));
    lat<1><<<1, 32>>>(out, cyc, 1024, 0.999f, 0.001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FFMA2 dependent latency %.2f cycles\n", (double)h / (1024 * 32));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int nt = 128; nt <= 512; nt *= 2)
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            const int iters = 2000;
            cudaEventRecord(e0);
            if (mode == 0) biq<0><<<148, nt>>>(out, iters, 0.9f, 0.1f, 0.05f, 0.3f, -0.2f);
            else biq<1><<<148, nt>>>(out, iters, 0.9f, 0.1f, 0.05f, 0.3f, -0.2f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("biquad %s %d thr/SM: %.3f ms  %.1f Gsample/s\n", mode ? "packed halves" : "scalar", nt, ms, 148.0 * nt * 32 * iters / ms / 1e6);
        }
    }
    printf("err %d\n", (int)cudaGetLastError());
    return 0;
}
