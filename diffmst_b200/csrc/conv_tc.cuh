// 3x3 convolution (stride 1, pad 1, no bias) + per-channel affine (folded BatchNorm) + ReLU on the
// 5th-generation tensor cores: the dense contraction of the Cnn14 ConvBlock (mst/panns.py:27-85).
//
// Layout: activations are NHWC float32 with a one-pixel zero border, flattened to a 2-D matrix
// [P = B*(H+2)*(W+2) pixels, C channels].  The convolution is nine shifted GEMMs accumulated
// into one TMEM tile:  out[p, n] = sum_{tap=(dy,dx)} sum_c in[p + dy*(W+2) + dx, c] * w[tap][n][c],
// evaluated for every padded pixel index p (border outputs are overwritten with zeros, which is
// exactly the zero border the next convolution needs).  So an A operand tile is a plain 2-D TMA
// box at a shifted row coordinate (out-of-range rows are zero-filled by TMA) and each B tile is
// a box of the repacked weights [9*Cout, Cin].  The three horizontal taps of one kernel row read
// pixel rows that differ by one: ONE 136-row A tile per (kernel row, k-chunk) serves all three, the
// MMA descriptors start 0, 1 and 2 rows (128 bytes) into it (the 128-byte swizzle is a function of the
// absolute shared-memory address bits, so a start address that is not 1024-byte aligned reads the
// rows TMA wrote - measured: the descriptor's base-offset field must stay 0).  That cuts the TMA traffic per tile by 1/3 - 2/5;
// the kernel was bound by L2 -> shared-memory bandwidth, not by the tensor pipe.
//
// Kernel: one CTA per (128 pixels x BN channels) tile, 192 threads:
//   warp 0     TMA producer   (cp.async.bulk.tensor.2d -> 128B-swizzled shared tiles, mbarrier tx)
//   warp 1     TMEM allocation + MMA issue (tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BN, K=8;
//              A and B from shared-memory descriptors, FP32 accumulator in TMEM;
//              tcgen05.commit releases pipeline stages and finally signals the epilogue)
//   warps 2-5  epilogue: tcgen05.ld 32x32b -> registers -> scale/shift (+ReLU) -> NHWC store
// TF32 operands (float32 bits, 10-bit mantissa used by the tensor core) with FP32 accumulation:
// the numerics class of the reference's cuDNN convolutions (torch enables TF32 convs by default).
#pragma once
#include "../../include/diffmst_b200.h"
#include "common.cuh"

#ifndef DMST_EMULATE
#include <cuda.h>

namespace dmst {

constexpr int kConvBM = 128;      // pixels per tile (UMMA M)
constexpr int kConvBK = 32;       // float32 elements per 128-byte swizzled row
constexpr int kConvARows = 136;   // A tile rows: 128 + 2 (horizontal taps), rounded up to the 8-row swizzle atom
constexpr int kConvThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4)            // D format: F32
           | (2u << 7)          // A format: TF32
           | (2u << 10)         // B format: TF32
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // K-major A and B
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Round-to-nearest conversion to TF32 (cvt.rna: the result is a float32 whose low 13 mantissa bits are zero).
// tcgen05.mma.kind::tf32 reads float32 bits from shared memory and IGNORES those bits, i.e. it truncates; with
// structured weights and non-negative activations the truncation errors add coherently over the twelve Cnn14
// layers (measured on the reference-golden case: first-layer gradient cosine 0.982 truncated, 0.9997 rounded =
// cuDNN's TF32 convolutions).  Nothing sits between TMA and the MMA, so every kernel that PRODUCES a tensor-core
// operand rounds it: weight repacks, the ReLU / pooling outputs that feed a convolution, the BatchNorm backward's
// dz, the layout conversions.
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ float4 tf32_rn4(float4 v) { return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w)); }
// operands of the tensor-core kernels have channel counts that are multiples of 32 (kConvBK); tensors with other
// channel counts go to the FP32 CUDA-core kernels and keep full precision
__host__ __device__ __forceinline__ bool tf32_operand_channels(int C) { return C % 32 == 0; }
constexpr int kReluBit = 1, kRoundTf32Bit = 2;   // the `relu` argument of the convolution entry points is this mask

struct ConvArgs {
    int P;            // padded pixels = B * Hp * Wp
    int Hp, Wp;       // H + 2, W + 2
    int Cin, Cout;
    const float* scale;   // [Cout] or null (= 1)
    const float* shift;   // [Cout] or null (= 0)
    int relu;
    float* out;       // [P][Cout], zero border written
    int ksplit;       // > 1: the (kernel row, k-chunk) stages of a tile are split over `ksplit` work items, each
    float* partial;   //      writing its raw accumulator to partial[split][P][Cout] (summed by conv_splitk_reduce_kernel)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent kernel: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (channel tile
// fastest, so CTAs running at the same time share A tiles in L2).  The shared-memory ring and
// its barriers run continuously across tiles; the accumulator is double-buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1.
__host__ __device__ constexpr int conv_stage_bytes(int BN, int MT) { return MT * kConvARows * kConvBK * 4 + 3 * BN * kConvBK * 4; }
// shared memory of the epilogue: per epilogue warp a 32 x 32 block staged in rows of 36 floats (16-byte aligned, and
// both the row-wise writes and the transposed reads are conflict-free with 128-bit accesses)
constexpr int kEpiRow = 36;
constexpr int kEpiBytes = 4 * 32 * kEpiRow * 4;
__host__ __device__ constexpr int conv_stages(int BN, int MT) {
    return (200 * 1024 / conv_stage_bytes(BN, MT)) > 4 ? 4 : (200 * 1024 / conv_stage_bytes(BN, MT));
}

// MT = 1 or 2 pixel sub-tiles of 128 per CTA tile: with MT = 2 the weight tiles of a stage feed two
// accumulators, which halves the weight traffic per flop (the big layers are bound by L2 -> shared
// memory bandwidth); small layers keep MT = 1 to have enough tiles for all SMs.
template <int BN, int MT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, ConvArgs a) {
    constexpr uint32_t kABytes = kConvARows * kConvBK * 4, kBBytes = BN * kConvBK * 4;
    constexpr uint32_t kStageBytes = MT * kABytes + 3 * kBBytes;   // MT A tiles + the weights of the 3 horizontal taps
    constexpr int kConvStages = conv_stages(BN, MT);
    constexpr int kTileM = MT * kConvBM;
    constexpr int kAcc = 2;  // TMEM accumulator stages
    extern __shared__ __align__(1024) unsigned char smem[];
    // 1024-byte aligned tiles (128B swizzle atoms)
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[kConvStages], empty_bar[kConvStages], tmem_full_bar[kAcc], tmem_empty_bar[kAcc];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = a.Cin / kConvBK, iters = 3 * kchunks;   // (kernel row, k-chunk) stages per tile
    const int tiles_n = a.Cout / BN;
    const int tiles_m = (a.P + kTileM - 1) / kTileM;
    const int ksplit = a.ksplit > 1 ? a.ksplit : 1;
    const int num_tiles = tiles_m * tiles_n * ksplit;   // work items: (tile, k-split), splits of a tile adjacent

    if (threadIdx.x == 0) {
        for (int s = 0; s < kConvStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAcc; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: kAcc x MT x BN FP32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(kAcc * MT * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;  // ring position, continuous across tiles
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int split = tile % ksplit, t2 = tile / ksplit;
                const int p0 = (t2 / tiles_n) * kTileM, n0 = (t2 % tiles_n) * BN;
                const int it0 = split * iters / ksplit, it1 = (split + 1) * iters / ksplit;
                for (int it = it0; it < it1; ++it, ++g) {
                    const uint32_t s = g % kConvStages, round = g / kConvStages;
                    mbar_wait(&empty_bar[s], (round & 1) ^ 1);   // passes immediately on the first round
                    const int ky = it / kchunks, kc = it - ky * kchunks;
                    unsigned char* sa = tiles + (size_t)s * kStageBytes;
                    mbar_expect_tx(&full_bar[s], kStageBytes);
                    // pixel rows p0 + (ky-1)*Wp - 1 ... + 135: horizontal tap kx reads rows kx ... kx+127 of this tile
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        tma_load_2d(sa + m * kABytes, &map_a, &full_bar[s], kc * kConvBK, p0 + m * kConvBM + (ky - 1) * a.Wp - 1);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        tma_load_2d(sa + MT * kABytes + kx * kBBytes, &map_b, &full_bar[s], kc * kConvBK, (ky * 3 + kx) * a.Cout + n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kConvBM, BN);
            uint32_t g = 0, t = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const uint32_t acc = t % kAcc;
                mbar_wait(&tmem_empty_bar[acc], ((t / kAcc) & 1) ^ 1);   // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * (MT * BN);
                const int split = tile % ksplit;
                const int it0 = split * iters / ksplit, it1 = (split + 1) * iters / ksplit;
                for (int it = it0; it < it1; ++it, ++g) {
                    const uint32_t s = g % kConvStages, round = g / kConvStages;
                    mbar_wait(&full_bar[s], round & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(tiles + (size_t)s * kStageBytes);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint64_t bdesc = umma_smem_desc(sa + MT * kABytes + kx * kBBytes);
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            // A: start kx rows (128 bytes each) into the sub-tile
                            const uint64_t adesc = umma_smem_desc(sa + m * kABytes + kx * 128);
#pragma unroll
                            for (int k = 0; k < kConvBK / 8; ++k)   // UMMA K = 8 tf32 = 32 bytes: advance the start address
                                umma_tf32(tmem_d + m * BN, adesc + (uint64_t)((k * 32) >> 4), bdesc + (uint64_t)((k * 32) >> 4),
                                          idesc, ((it - it0) | kx | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[s]);           // stage free once these MMAs have read it
                }
                umma_commit(&tmem_full_bar[acc]);          // accumulator complete
            }
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int quarter = warp & 3;
        // Each thread holds one pixel's 32 channels (128 contiguous bytes of a row that is Cout*4 bytes long): stored
        // directly, a warp instruction touches 32 different lines with 16 bytes each.  Staged through shared memory and
        // read back transposed, 8 lanes write one pixel's 128 bytes and a warp instruction writes four whole lines
        // (measured: 64 -> 128 at 512 x 128: 165 -> 134 us, 128 -> 128: 249 -> 240 us).
        float* stage = reinterpret_cast<float*>(tiles + (size_t)kConvStages * kStageBytes) + (warp - 2) * 32 * kEpiRow;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
            const int split = tile % ksplit, t2 = tile / ksplit;
            const int p0 = (t2 / tiles_n) * kTileM, n0 = (t2 % tiles_n) * BN;
            const uint32_t acc = t % kAcc;
            mbar_wait(&tmem_full_bar[acc], (t / kAcc) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int m = 0; m < MT; ++m) {
                const int p = p0 + m * kConvBM + quarter * 32 + lane;
                bool interior = false;
                if (p < a.P) {
                    const int rem = p % (a.Hp * a.Wp);
                    const int hp = rem / a.Wp, wp = rem - hp * a.Wp;
                    interior = hp >= 1 && hp <= a.Hp - 2 && wp >= 1 && wp <= a.Wp - 2;
                }
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (MT * BN) + m * BN + (uint32_t)c0;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (m == MT - 1 && c0 + 32 >= BN) {  // accumulator fully read: hand it back to the MMA warp before the stores
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                    }
                    if (ksplit > 1) {   // raw partial sums; the reduction kernel applies the epilogue
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(stage + lane * kEpiRow + j) =
                                make_float4(__int_as_float((int)r[j]), __int_as_float((int)r[j + 1]), __int_as_float((int)r[j + 2]),
                                            __int_as_float((int)r[j + 3]));
                    } else {
                        // affine + ReLU + optional TF32 rounding: 128-bit loads of the per-channel constants, ReLU as a
                        // maximum with 0 or -inf, flags tested once per 4 channels (the epilogue must keep pace with the MMAs)
                        const float lo = (a.relu & kReluBit) ? 0.0f : -__int_as_float(0x7f800000);
                        const bool rnd = (a.relu & kRoundTf32Bit) != 0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int n = n0 + c0 + j;
                            const float4 sc = a.scale ? __ldg(reinterpret_cast<const float4*>(a.scale + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
                            const float4 sh = a.shift ? __ldg(reinterpret_cast<const float4*>(a.shift + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            float4 v;
                            v.x = fmaxf(fmaf(__int_as_float((int)r[j]), sc.x, sh.x), lo);
                            v.y = fmaxf(fmaf(__int_as_float((int)r[j + 1]), sc.y, sh.y), lo);
                            v.z = fmaxf(fmaf(__int_as_float((int)r[j + 2]), sc.z, sh.z), lo);
                            v.w = fmaxf(fmaf(__int_as_float((int)r[j + 3]), sc.w, sh.w), lo);
                            if (rnd) v = tf32_rn4(v);
                            if (!interior) v = make_float4(0.f, 0.f, 0.f, 0.f);
                            *reinterpret_cast<float4*>(stage + lane * kEpiRow + j) = v;
                        }
                    }
                    __syncwarp();
                    {   // transposed read-back: lane -> (pixel lane / 8 + 4 it, channels 4 (lane % 8) ...)
                        const int pbase = p0 + m * kConvBM + quarter * 32;
                        float* obase = (ksplit > 1 ? a.partial + (size_t)split * a.P * a.Cout : a.out) + n0 + c0 + 4 * (lane & 7);
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int px = 4 * it + (lane >> 3);
                            if (pbase + px < a.P)
                                *reinterpret_cast<float4*>(obase + (size_t)(pbase + px) * a.Cout) =
                                    *reinterpret_cast<const float4*>(stage + px * kEpiRow + 4 * (lane & 7));
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kAcc * MT * BN));
}

// ---- small helper kernels around the tensor-core convolution ----

// NCHW (B, C, H, W) -> zero-bordered NHWC (B, H+2, W+2, C); grid covers B*Hp*Wp*C elements
__global__ void nchw_to_padded_nhwc_kernel(const float* x, float* y, int B, int C, int H, int W) {
    const int Hp = H + 2, Wp = W + 2;
    const long long total = (long long)B * Hp * Wp * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int wp = (int)(r % Wp); r /= Wp;
        const int hp = (int)(r % Hp);
        const int b = (int)(r / Hp);
        float v = 0.0f;
        if (hp >= 1 && hp <= H && wp >= 1 && wp <= W) v = __ldg(x + (((long long)b * C + c) * H + (hp - 1)) * W + (wp - 1));
        y[i] = tf32_operand_channels(C) ? tf32_rn(v) : v;
    }
}

// direct 3x3 convolution for the first layer (Cin is tiny: 1 spectrogram channel), padded NHWC in/out
__global__ void conv3x3_direct_kernel(const float* x, const float* w9 /*[9][Cout][Cin]*/, const float* scale,
                                      const float* shift, float* y, int P, int Hp, int Wp, int Cin, int Cout, int relu) {
    const long long total = (long long)P * Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % Cout);
        const long long p = i / Cout;
        const int rem = (int)(p % ((long long)Hp * Wp));
        const int hp = rem / Wp, wp = rem - hp * Wp;
        float acc = 0.0f;
        const bool interior = hp >= 1 && hp <= Hp - 2 && wp >= 1 && wp <= Wp - 2;
        if (interior) {
            for (int tap = 0; tap < 9; ++tap) {
                const long long q = p + (tap / 3 - 1) * Wp + (tap % 3 - 1);
                for (int c = 0; c < Cin; ++c) acc = fmaf(__ldg(x + q * Cin + c), __ldg(w9 + ((long long)tap * Cout + n) * Cin + c), acc);
            }
            if (scale) acc *= __ldg(scale + n);
            if (shift) acc += __ldg(shift + n);
            if (relu & kReluBit) acc = fmaxf(acc, 0.0f);
            if (relu & kRoundTf32Bit) acc = tf32_rn(acc);
        }
        y[i] = acc;
    }
}

// First layer (Cin = 1 spectrogram channel; any Cin <= 4 with Cout % 4 == 0): the output write is the whole
// cost, so this is a streaming kernel: weights and the affine in shared memory, each thread produces 4
// consecutive output channels of one pixel (a warp writes whole 128-byte lines), inputs are warp-broadcast.
constexpr int kSmallCinMax = 4;
__global__ void conv3x3_small_cin_kernel(const float* x, const float* w9 /*[9][Cout][Cin]*/, const float* scale,
                                         const float* shift, float* y, int P, int Hp, int Wp, int Cin, int Cout, int relu) {
    extern __shared__ float sm_w[];   // [9][Cin][Cout] (channel fastest), then scale[Cout], shift[Cout]
    float* sm_scale = sm_w + 9 * Cin * Cout;
    float* sm_shift = sm_scale + Cout;
    for (int i = threadIdx.x; i < 9 * Cin * Cout; i += blockDim.x) {
        const int n = i % Cout, c = (i / Cout) % Cin, tap = i / (Cout * Cin);
        sm_w[i] = __ldg(w9 + ((long long)tap * Cout + n) * Cin + c);
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) {
        sm_scale[i] = scale ? __ldg(scale + i) : 1.0f;
        sm_shift[i] = shift ? __ldg(shift + i) : 0.0f;
    }
    __syncthreads();
    const int groups = Cout >> 2;                          // float4 channel groups per pixel
    const int pix_per_block = blockDim.x / groups;
    const int g = threadIdx.x % groups, lp = threadIdx.x / groups;
    if (lp >= pix_per_block) return;
    const int plane = Hp * Wp;
    for (int p = blockIdx.x * pix_per_block + lp; p < P; p += gridDim.x * pix_per_block) {   // P < 2^31 (checked on the host)
        const int rem = p % plane;
        const int hp = rem / Wp, wp = rem - hp * Wp;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hp >= 1 && hp <= Hp - 2 && wp >= 1 && wp <= Wp - 2) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int q = p + (tap / 3 - 1) * Wp + (tap % 3 - 1);
                for (int c = 0; c < Cin; ++c) {
                    const float xv = __ldg(x + (long long)q * Cin + c);
                    const float4 w = *reinterpret_cast<const float4*>(sm_w + (tap * Cin + c) * Cout + 4 * g);
                    acc.x = fmaf(xv, w.x, acc.x); acc.y = fmaf(xv, w.y, acc.y);
                    acc.z = fmaf(xv, w.z, acc.z); acc.w = fmaf(xv, w.w, acc.w);
                }
            }
            const float4 sc = *reinterpret_cast<const float4*>(sm_scale + 4 * g);
            const float4 sh = *reinterpret_cast<const float4*>(sm_shift + 4 * g);
            acc.x = fmaf(acc.x, sc.x, sh.x); acc.y = fmaf(acc.y, sc.y, sh.y);
            acc.z = fmaf(acc.z, sc.z, sh.z); acc.w = fmaf(acc.w, sc.w, sh.w);
            if (relu & kReluBit) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
            if (relu & kRoundTf32Bit) acc = tf32_rn4(acc);
        }
        *reinterpret_cast<float4*>(y + (long long)p * Cout + 4 * g) = acc;
    }
}

// Cin == 1 specialisation of the streaming first-layer kernel: the 9 x 4 weights and the affine of a thread's
// channel group live in registers for its whole life; per pixel that leaves 9 loads, 36 FMAs and one 16-byte store.
__global__ void __launch_bounds__(256) conv3x3_cin1_kernel(const float* __restrict__ x, const float* __restrict__ w9 /*[9][Cout]*/,
                                                           const float* scale, const float* shift, float* __restrict__ y,
                                                           int P, int Hp, int Wp, int Cout, int relu) {
    const int groups = Cout >> 2;
    const int pix_per_block = blockDim.x / groups;
    const int g = threadIdx.x % groups, lp = threadIdx.x / groups;
    if (lp >= pix_per_block) return;
    float4 w[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) w[tap] = __ldg(reinterpret_cast<const float4*>(w9 + tap * Cout + 4 * g));
    const float4 sc = scale ? __ldg(reinterpret_cast<const float4*>(scale + 4 * g)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 sh = shift ? __ldg(reinterpret_cast<const float4*>(shift + 4 * g)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int plane = Hp * Wp;
    for (int p = blockIdx.x * pix_per_block + lp; p < P; p += gridDim.x * pix_per_block) {   // P < 2^31 (checked on the host)
        const int rem = p % plane;
        const int hp = rem / Wp, wp = rem - hp * Wp;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hp >= 1 && hp <= Hp - 2 && wp >= 1 && wp <= Wp - 2) {
            const float* xp = x + p;
            float xv[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) xv[tap] = __ldg(xp + (tap / 3 - 1) * Wp + (tap % 3 - 1));
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                acc.x = fmaf(xv[tap], w[tap].x, acc.x); acc.y = fmaf(xv[tap], w[tap].y, acc.y);
                acc.z = fmaf(xv[tap], w[tap].z, acc.z); acc.w = fmaf(xv[tap], w[tap].w, acc.w);
            }
            acc.x = fmaf(acc.x, sc.x, sh.x); acc.y = fmaf(acc.y, sc.y, sh.y);
            acc.z = fmaf(acc.z, sc.z, sh.z); acc.w = fmaf(acc.w, sc.w, sh.w);
            if (relu & kReluBit) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
            if (relu & kRoundTf32Bit) acc = tf32_rn4(acc);
        }
        *reinterpret_cast<float4*>(y + (long long)p * Cout + 4 * g) = acc;
    }
}

// average pool (kh x kw, stride = kernel, floor) of padded NHWC -> NCHW (B, C, H/kh, W/kw) or padded NHWC
__global__ void avgpool_kernel(const float* x, float* y, int B, int C, int H, int W, int kh, int kw, int out_nhwc_padded) {
    const int Ho = H / kh, Wo = W / kw, Hp = H + 2, Wp = W + 2;
    const float inv = 1.0f / (float)(kh * kw);
    if ((C & 3) == 0) {
        // one thread per (output pixel, 4 channels): 128-bit loads over the window, channel-fastest => coalesced
        const int C4 = C >> 2;
        const long long total = (long long)B * Ho * Wo * C4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const unsigned pix = (unsigned)(i / (unsigned)C4);          // output pixel index < 2^31
            const int c = (int)(i - (long long)pix * C4) << 2;
            const unsigned row = pix / (unsigned)Wo;
            const int wo = (int)(pix - row * Wo);
            const int b = (int)(row / (unsigned)Ho);
            const int ho = (int)(row - (unsigned)b * Ho);
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* base = x + (((long long)b * Hp + (ho * kh + 1)) * Wp + (wo * kw + 1)) * C + c;
            for (int dy = 0; dy < kh; ++dy)
                for (int dx = 0; dx < kw; ++dx) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(base + ((long long)dy * Wp + dx) * C));
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
            s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
            if (out_nhwc_padded) {   // feeds the next block's convolution
                if (tf32_operand_channels(C)) s = tf32_rn4(s);
                *reinterpret_cast<float4*>(y + (((long long)b * (Ho + 2) + ho + 1) * (Wo + 2) + wo + 1) * C + c) = s;
            } else {
                const long long o = (((long long)b * C + c) * Ho + ho) * Wo + wo, cs = (long long)Ho * Wo;
                y[o] = s.x; y[o + cs] = s.y; y[o + 2 * cs] = s.z; y[o + 3 * cs] = s.w;
            }
        }
        return;
    }
    const long long total = (long long)B * Ho * Wo * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int b = (int)(r / Ho);
        float s = 0.0f;
        for (int dy = 0; dy < kh; ++dy)
            for (int dx = 0; dx < kw; ++dx)
                s += __ldg(x + (((long long)b * Hp + (ho * kh + dy + 1)) * Wp + (wo * kw + dx + 1)) * C + c);
        s *= inv;
        if (out_nhwc_padded) y[(((long long)b * (Ho + 2) + ho + 1) * (Wo + 2) + wo + 1) * C + c] = s;
        else y[(((long long)b * C + c) * Ho + ho) * Wo + wo] = s;
    }
}

// dst = src rounded to nearest TF32, optionally with the one-pixel border of a [P][C] zero-bordered NHWC tensor
// cleared (Hp == 0: plain element-wise).  For tensors that reach a tensor-core convolution from outside this library
// (a user's input, an upstream gradient produced by PyTorch ops).
__global__ void round_tf32_kernel(const float* src, float* dst, long long n, int Hp, int Wp, int C) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = tf32_rn(__ldg(src + i));
        if (Hp > 0) {
            const long long p = i / C;
            const int rem = (int)(p % ((long long)Hp * Wp));
            const int hp = rem / Wp, wp = rem - hp * Wp;
            if (hp == 0 || hp == Hp - 1 || wp == 0 || wp == Wp - 1) v = 0.0f;
        }
        dst[i] = v;
    }
}

// [Cout][Cin][3][3] -> [9][Cout][Cin]
// (rounded to TF32 when the tensor-core kernel will read them: conv3x3_forward's dispatch rule)
__host__ __device__ __forceinline__ bool conv_uses_tensor_cores(int Cin, int Cout) { return Cin % kConvBK == 0 && Cout % 64 == 0; }
__global__ void repack_weights_kernel(const float* w, float* w9, int Cout, int Cin) {
    const int total = 9 * Cout * Cin;
    const bool rnd = conv_uses_tensor_cores(Cin, Cout);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % Cin, n = (i / Cin) % Cout, tap = i / (Cin * Cout);
        const float v = __ldg(w + ((long long)n * Cin + c) * 9 + tap);
        w9[i] = rnd ? tf32_rn(v) : v;
    }
}


// [Cout][Cin][3][3] -> [9][Cin][Cout] with the taps flipped: the weights of the input-gradient convolution (dgrad =
// the same shifted GEMM with the channel roles swapped).  One block per 32 x 32 (Cout, Cin) tile: 288 contiguous
// floats are read per output channel, 32 contiguous output channels written per (tap, input channel).
// grid: (ceil(Cin/32), ceil(Cout/32)), block 256
__global__ void repack_weights_dgrad_kernel(const float* w, float* w9t, int Cout, int Cin) {
    __shared__ float tile[32][289];
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    const int nci = min(32, Cin - ci0);
    for (int i = threadIdx.x; i < 32 * 288; i += blockDim.x) {
        const int r = i / 288, q = i - r * 288;
        if (co0 + r < Cout && q < nci * 9) tile[r][q] = __ldg(w + ((size_t)(co0 + r) * Cin + ci0) * 9 + q);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 32 * 32; i += blockDim.x) {
        const int co = i & 31, ci = (i >> 5) & 31, tap = i >> 10;
        if (co0 + co < Cout && ci < nci) {   // dgrad runs conv3x3_forward with the channel roles swapped
            const float v = tile[co][ci * 9 + tap];
            w9t[((size_t)(8 - tap) * Cin + ci0 + ci) * Cout + co0 + co] = conv_uses_tensor_cores(Cout, Cin) ? tf32_rn(v) : v;
        }
    }
}

// ---- training-mode BatchNorm support: per-channel batch statistics of a raw conv output ----
// The zero border contributes nothing, so sums over all P rows are sums over the B*H*W pixels.
// rows of a [P][C] tensor per reduction chunk (= per CTA): the grid covers the device about eight times (at most
// ~1200 chunks, so that the second-stage kernels - one warp per channel over the chunks - stay a few microseconds
// at the training step's 10 M-row activations: with a 512-row cap they were 75 us each, 7 % of the step), never
// fewer than 32 rows on the small deep layers
__host__ __device__ inline int stat_rows(long long P) {
    const long long r = (P + 8 * 148 - 1) / (8 * 148);
    return (int)(r < 32 ? 32 : r);
}
// Thread layout of the per-channel reductions over a chunk of stat_rows(P) rows of a [P][C] tensor: a block of 256
// threads covers min(C/4, 256) float4 columns x (256 / columns) rows at a time (all threads load 128 bits,
// consecutive threads consecutive addresses); the row lanes are then added in a fixed order through shared memory.
struct StatLayout {
    int cols, lanes, cq, rl;
    __device__ __forceinline__ StatLayout(int C) {
        const int C4 = C >> 2;
        cols = C4 < 256 ? C4 : 256;
        lanes = 256 / cols;
        cq = threadIdx.x % cols;
        rl = threadIdx.x / cols;
    }
};
// sums the per-lane float4 pairs of one column group over the row lanes and stores them: partial[chunk][0/1][c..c+3]
__device__ __forceinline__ void stat_lane_reduce(const StatLayout& L, float4 a, float4 b, float4* sh /*[2][256]*/,
                                                 float* partial, int chunk, int C, int c) {
    sh[threadIdx.x] = a;
    sh[256 + threadIdx.x] = b;
    __syncthreads();
    if (L.rl == 0 && c < C) {
        for (int l = 1; l < L.lanes; ++l) {
            const float4 u = sh[l * L.cols + L.cq], v = sh[256 + l * L.cols + L.cq];
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
        }
        *reinterpret_cast<float4*>(partial + ((size_t)chunk * 2 + 0) * C + c) = a;
        *reinterpret_cast<float4*>(partial + ((size_t)chunk * 2 + 1) * C + c) = b;
    }
    __syncthreads();
}
// C % 4 == 0 (every Cnn14 layer); blockDim.x == 256.  CTAs take the chunks last-first: the tensor was just written
// front to back by its producer, so its tail is what the L2 still holds (and the pass leaves the front in L2 for the
// consumer that follows and reads front to back).
__global__ void __launch_bounds__(256) channel_partial_kernel(const float* y, int P, int C, float* partial /*[chunks][2][C]*/) {
    __shared__ float4 sh[512];
    const StatLayout L(C);
    const int chunk = gridDim.x - 1 - blockIdx.x;
    const int rows = stat_rows(P);
    const int r0 = chunk * rows, r1 = min(r0 + rows, P);
    for (int g = 0; g * L.cols < (C >> 2); ++g) {
        const int c = (g * L.cols + L.cq) << 2;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s;
        if (c < C && L.rl < L.lanes)
#pragma unroll 8
            for (int r = r0 + L.rl; r < r1; r += L.lanes) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(y + (size_t)r * C + c));
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
            }
        stat_lane_reduce(L, s, s2, sh, partial, chunk, C, c);
    }
}
// one warp per channel: lanes stride over the chunks in float64, fixed-order butterfly at the end
__device__ __forceinline__ void stat_chunk_sums(const float* partial, int chunks, int C, int c, double& s, double& s2) {
    const int lane = threadIdx.x & 31;
    s = 0.0; s2 = 0.0;
    for (int k = lane; k < chunks; k += 32) { s += partial[((size_t)k * 2 + 0) * C + c]; s2 += partial[((size_t)k * 2 + 1) * C + c]; }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
}
// grid: ceil(C / 8) blocks of 256 threads
__global__ void channel_final_kernel(const float* partial, int chunks, int C, double count, float* mean, float* var_biased) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    double s, s2;
    stat_chunk_sums(partial, chunks, C, c, s, s2);
    if ((threadIdx.x & 31) == 0) {
        const double m = s / count;
        mean[c] = (float)m;
        var_biased[c] = (float)fmax(s2 / count - m * m, 0.0);
    }
}
// in-place y = relu(y * scale[c] + shift[c]) on interior pixels (border stays zero)
__global__ void affine_relu_kernel(float* y, int P, int Hp, int Wp, int C, const float* scale, const float* shift, int relu) {
    const long long total = (long long)P * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long p = i / C;
        const int rem = (int)(p % ((long long)Hp * Wp));
        const int hp = rem / Wp, wp = rem - hp * Wp;
        if (hp >= 1 && hp <= Hp - 2 && wp >= 1 && wp <= Wp - 2) {
            float v = fmaf(y[i], __ldg(scale + c), __ldg(shift + c));
            if (relu & kReluBit) v = fmaxf(v, 0.0f);
            y[i] = (relu & kRoundTf32Bit) ? tf32_rn(v) : v;
        }
    }
}
// out-of-place twin of affine_relu_kernel (the differentiable path keeps the raw convolution output):
// y = relu(z * scale[c] + shift[c]) on interior pixels, zero on the border
__global__ void affine_relu_to_kernel(const float* z, float* y, int P, int Hp, int Wp, int C, const float* scale,
                                      const float* shift) {
    const int C4 = C >> 2;
    const long long total = (long long)P * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / C4), c = (int)(i - (long long)p * C4) << 2;
        const int rem = p % (Hp * Wp);
        const int hp = rem / Wp, wp = rem - hp * Wp;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hp >= 1 && hp <= Hp - 2 && wp >= 1 && wp <= Wp - 2) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(z + (size_t)p * C + c));
            const float4 a = __ldg(reinterpret_cast<const float4*>(scale + c)), b = __ldg(reinterpret_cast<const float4*>(shift + c));
            o.x = fmaxf(fmaf(v.x, a.x, b.x), 0.f); o.y = fmaxf(fmaf(v.y, a.y, b.y), 0.f);
            o.z = fmaxf(fmaf(v.z, a.z, b.z), 0.f); o.w = fmaxf(fmaf(v.w, a.w, b.w), 0.f);
            if (tf32_operand_channels(C)) o = tf32_rn4(o);   // y feeds the block's second convolution
        }
        *reinterpret_cast<float4*>(y + (size_t)p * C + c) = o;
    }
}

// ---- backward of y = relu(BatchNorm(z)) (mst/panns.py:79-80 under autograd) on zero-bordered NHWC ----
// With zhat = (z - mean) * rstd, dyr = dy * [scale * z + shift > 0]:
//   dbeta = sum dyr, dgamma = sum dyr * zhat,
//   batch statistics:   dz = gamma * rstd * (dyr - dbeta / n - zhat * dgamma / n)
//   running statistics: dz = gamma * rstd * dyr
// Pass 1 accumulates the two sums per channel over chunks of rows (fixed order => deterministic), pass 2 adds the
// chunks in float64, pass 3 writes dz (zero on the border).
// Where the gradient w.r.t. y = relu(BatchNorm(z)) comes from: a full zero-bordered NHWC tensor (dpool == nullptr), or
// the gradient of the average-pooled y, gathered on the fly (the pooled unit never materialises y or its gradient):
// dy[b][h][w][c] = dpool[b][h >> sh][w >> sw][c] / (kh * kw) inside the pooled region, 0 outside.  Pool sizes are
// powers of two (every Cnn14 block).
struct PoolGrad {
    const float* dpool;   // (B, C, Ho, Wo) or zero-bordered NHWC (B, Ho+2, Wo+2, C)
    int sh, sw, Ho, Wo, padded;
    float inv;
};
__device__ __forceinline__ float4 load_dy4(const float* dy, const PoolGrad& pg, size_t r, int b, int hp, int wp, int C, int c) {
    if (pg.dpool == nullptr) return __ldg(reinterpret_cast<const float4*>(dy + r * C + c));
    const int ho = (hp - 1) >> pg.sh, wo = (wp - 1) >> pg.sw;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hp >= 1 && wp >= 1 && ho < pg.Ho && wo < pg.Wo) {
        if (pg.padded) {
            v = __ldg(reinterpret_cast<const float4*>(pg.dpool + (((size_t)b * (pg.Ho + 2) + ho + 1) * (pg.Wo + 2) + wo + 1) * C + c));
        } else {
            const size_t o = (((size_t)b * C + c) * pg.Ho + ho) * pg.Wo + wo, cs = (size_t)pg.Ho * pg.Wo;
            v = make_float4(__ldg(pg.dpool + o), __ldg(pg.dpool + o + cs), __ldg(pg.dpool + o + 2 * cs), __ldg(pg.dpool + o + 3 * cs));
        }
        v.x *= pg.inv; v.y *= pg.inv; v.z *= pg.inv; v.w *= pg.inv;
    }
    return v;
}
__global__ void __launch_bounds__(256) bn_relu_bwd_partial_kernel(const float* z, const float* dy, PoolGrad pg, int P, int Hp, int Wp, int C,
                                                                  const float* scale, const float* shift, const float* mean,
                                                                  const float* rstd, float* partial /*[chunks][2][C]*/) {
    __shared__ float4 sh[512];
    const StatLayout L(C);
    const int chunk = gridDim.x - 1 - blockIdx.x;   // last-first, as channel_partial_kernel: dy was just written front to back
    const int rows = stat_rows(P);
    const int r0 = chunk * rows, r1 = min(r0 + rows, P);
    for (int g = 0; g * L.cols < (C >> 2); ++g) {
        const int c = (g * L.cols + L.cq) << 2;
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
        if (c < C && L.rl < L.lanes) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(scale + c)), b = __ldg(reinterpret_cast<const float4*>(shift + c));
            const float4 m = __ldg(reinterpret_cast<const float4*>(mean + c)), rs = __ldg(reinterpret_cast<const float4*>(rstd + c));
            // (hp, wp) of the row, advanced incrementally: no division in the loop
            int img = (r0 + L.rl) / (Hp * Wp);
            int rem = (r0 + L.rl) - img * (Hp * Wp);
            int hp = rem / Wp, wp = rem - hp * Wp;
#pragma unroll 2
            for (int r = r0 + L.rl; r < r1; r += L.lanes) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(z + (size_t)r * C + c));
                const float4 d = load_dy4(dy, pg, (size_t)r, img, hp, wp, C, c);
                const bool in = hp >= 1 && hp <= Hp - 2 && wp >= 1 && wp <= Wp - 2;   // border rows are padding
                wp += L.lanes;
                while (wp >= Wp) { wp -= Wp; if (++hp == Hp) { hp = 0; ++img; } }
                const float gx = (in && fmaf(v.x, a.x, b.x) > 0.0f) ? d.x : 0.0f, gy = (in && fmaf(v.y, a.y, b.y) > 0.0f) ? d.y : 0.0f;
                const float gz = (in && fmaf(v.z, a.z, b.z) > 0.0f) ? d.z : 0.0f, gw = (in && fmaf(v.w, a.w, b.w) > 0.0f) ? d.w : 0.0f;
                s1.x += gx; s1.y += gy; s1.z += gz; s1.w += gw;
                s2.x = fmaf(gx, (v.x - m.x) * rs.x, s2.x); s2.y = fmaf(gy, (v.y - m.y) * rs.y, s2.y);
                s2.z = fmaf(gz, (v.z - m.z) * rs.z, s2.z); s2.w = fmaf(gw, (v.w - m.w) * rs.w, s2.w);
            }
        }
        stat_lane_reduce(L, s1, s2, sh, partial, chunk, C, c);
    }
}
// grid: ceil(C / 8) blocks of 256 threads
__global__ void bn_relu_bwd_final_kernel(const float* partial, int chunks, int C, float* dgamma, float* dbeta) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    double s1, s2;
    stat_chunk_sums(partial, chunks, C, c, s1, s2);
    if ((threadIdx.x & 31) == 0) { dbeta[c] = (float)s1; dgamma[c] = (float)s2; }
}
__global__ void bn_relu_bwd_apply_kernel(const float* z, const float* dy, PoolGrad pg, float* dz, int P, int Hp, int Wp, int C,
                                         const float* scale, const float* shift, const float* mean, const float* rstd,
                                         const float* dgamma, const float* dbeta, float inv_count /* 0: running statistics */) {
    const int C4 = C >> 2;
    const long long total = (long long)P * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / C4), c = (int)(i - (long long)p * C4) << 2;
        const int img = p / (Hp * Wp);
        const int rem = p - img * (Hp * Wp);
        const int hp = rem / Wp, wp = rem - hp * Wp;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hp >= 1 && hp <= Hp - 2 && wp >= 1 && wp <= Wp - 2) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(z + (size_t)p * C + c));
            const float4 g4 = load_dy4(dy, pg, (size_t)p, img, hp, wp, C, c);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(scale + c)), b4 = __ldg(reinterpret_cast<const float4*>(shift + c));
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean + c)), r4 = __ldg(reinterpret_cast<const float4*>(rstd + c));
            const float4 dg4 = __ldg(reinterpret_cast<const float4*>(dgamma + c)), db4 = __ldg(reinterpret_cast<const float4*>(dbeta + c));
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
            const float rs[4] = {r4.x, r4.y, r4.z, r4.w}, dg[4] = {dg4.x, dg4.y, dg4.z, dg4.w}, db[4] = {db4.x, db4.y, db4.z, db4.w};
            float r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float gr = fmaf(v[e], a[e], b[e]) > 0.0f ? g[e] : 0.0f;
                const float zh = (v[e] - m[e]) * rs[e];
                r[e] = a[e] * (gr - inv_count * (db[e] + zh * dg[e]));
            }
            o = make_float4(r[0], r[1], r[2], r[3]);
            if (tf32_operand_channels(C)) o = tf32_rn4(o);   // dz is the dgrad / wgrad operand
        }
        *reinterpret_cast<float4*>(dz + (size_t)p * C + c) = o;
    }
}

// y = avgpool(relu(z * scale + shift)): the second unit of a ConvBlock under autograd never materialises its
// BatchNorm+ReLU output (mst/panns.py:80-85); C % 4 == 0
__global__ void bn_relu_avgpool_kernel(const float* z, const float* scale, const float* shift, float* y, int B, int C, int H, int W,
                                       int kh, int kw, int out_nhwc_padded) {
    const int Ho = H / kh, Wo = W / kw, Hp = H + 2, Wp = W + 2;
    const float inv = 1.0f / (float)(kh * kw);
    const int C4 = C >> 2;
    const long long total = (long long)B * Ho * Wo * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const unsigned pix = (unsigned)(i / (unsigned)C4);
        const int c = (int)(i - (long long)pix * C4) << 2;
        const unsigned row = pix / (unsigned)Wo;
        const int wo = (int)(pix - row * Wo);
        const int b = (int)(row / (unsigned)Ho);
        const int ho = (int)(row - (unsigned)b * Ho);
        const float4 a = __ldg(reinterpret_cast<const float4*>(scale + c)), sft = __ldg(reinterpret_cast<const float4*>(shift + c));
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* base = z + (((long long)b * Hp + (ho * kh + 1)) * Wp + (wo * kw + 1)) * C + c;
        for (int dy = 0; dy < kh; ++dy)
            for (int dx = 0; dx < kw; ++dx) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(base + ((long long)dy * Wp + dx) * C));
                s.x += fmaxf(fmaf(v.x, a.x, sft.x), 0.f); s.y += fmaxf(fmaf(v.y, a.y, sft.y), 0.f);
                s.z += fmaxf(fmaf(v.z, a.z, sft.z), 0.f); s.w += fmaxf(fmaf(v.w, a.w, sft.w), 0.f);
            }
        s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
        if (out_nhwc_padded) {   // feeds the next block's convolution
            if (tf32_operand_channels(C)) s = tf32_rn4(s);
            *reinterpret_cast<float4*>(y + (((long long)b * (Ho + 2) + ho + 1) * (Wo + 2) + wo + 1) * C + c) = s;
        } else {
            const long long o = (((long long)b * C + c) * Ho + ho) * Wo + wo, cs = (long long)Ho * Wo;
            y[o] = s.x; y[o + cs] = s.y; y[o + 2 * cs] = s.z; y[o + 3 * cs] = s.w;
        }
    }
}

// backward of avgpool_kernel: dx (zero-bordered NHWC) from dy (NCHW (B, C, Ho, Wo) or zero-bordered NHWC)
__global__ void avgpool_bwd_kernel(const float* dy, float* dx, int B, int C, int H, int W, int kh, int kw, int dy_nhwc_padded) {
    const int Ho = H / kh, Wo = W / kw, Hp = H + 2, Wp = W + 2;
    const float inv = 1.0f / (float)(kh * kw);
    const int C4 = C >> 2;   // C % 4 == 0 checked by the caller
    const long long total = (long long)B * Hp * Wp * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const unsigned pix = (unsigned)(i / (unsigned)C4);
        const int c = (int)(i - (long long)pix * C4) << 2;
        const unsigned row = pix / (unsigned)Wp;
        const int wp = (int)(pix - row * Wp);
        const int b = (int)(row / (unsigned)Hp);
        const int hp = (int)(row - (unsigned)b * Hp);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int ho = (hp - 1) / kh, wo = (wp - 1) / kw;
        if (hp >= 1 && wp >= 1 && ho < Ho && wo < Wo) {
            if (dy_nhwc_padded) {
                v = __ldg(reinterpret_cast<const float4*>(dy + (((long long)b * (Ho + 2) + ho + 1) * (Wo + 2) + wo + 1) * C + c));
            } else {
                const long long o = (((long long)b * C + c) * Ho + ho) * Wo + wo, cs = (long long)Ho * Wo;
                v = make_float4(__ldg(dy + o), __ldg(dy + o + cs), __ldg(dy + o + 2 * cs), __ldg(dy + o + 3 * cs));
            }
            v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        }
        *reinterpret_cast<float4*>(dx + (size_t)pix * C + c) = v;
    }
}
inline int grid_for(long long total) { long long b = (total + 255) / 256; return (int)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b)); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// 2-D float32 row-major matrix [rows][cols], box = [box_rows][32 cols] with 128-byte swizzle
inline int make_map_2d(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn f = encode_tiled();
    if (!f) return 2000;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)kConvBK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = f(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2001 + (int)r;
}

// Sum of the split-K partial accumulators (fixed order => deterministic) + affine + ReLU + zero border
__global__ void conv_splitk_reduce_kernel(const float* partial, int ksplit, float* y, int P, int Hp, int Wp, int C,
                                          const float* scale, const float* shift, int relu) {
    const int C4 = C >> 2;
    const long long total = (long long)P * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / C4), c = (int)(i - (long long)p * C4) << 2;
        const int rem = p % (Hp * Wp);
        const int hp = rem / Wp, wp = rem - hp * Wp;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hp >= 1 && hp <= Hp - 2 && wp >= 1 && wp <= Wp - 2) {
            for (int k = 0; k < ksplit; ++k) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)k * P + p) * C + c));
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            if (scale) { const float4 q = __ldg(reinterpret_cast<const float4*>(scale + c)); s.x *= q.x; s.y *= q.y; s.z *= q.z; s.w *= q.w; }
            if (shift) { const float4 q = __ldg(reinterpret_cast<const float4*>(shift + c)); s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w; }
            if (relu & kReluBit) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
            if (relu & kRoundTf32Bit) s = tf32_rn4(s);
        }
        *reinterpret_cast<float4*>(y + (size_t)p * C + c) = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores.  dW[tap][co][ci] = sum_p dz[p][co] * x[p + off(tap)][ci] over all padded
// pixels p (dz is zero on the border, so products that would cross an image edge vanish): per tap a GEMM with
// M = Cout, N = Cin and K = pixels.  Both operands are stored with their M / N dimension contiguous
// ([pixels][channels]), so they are fed as MN-MAJOR UMMA operands: a TMA box of 32 pixel rows x 32 channels
// (128-byte rows, 32-byte-chunk swizzle) is one "slab", K = 8 pixel rows are 1024 bytes, and the tap shift is
// a shift of the box's ROW coordinate, which TMA does not constrain (with pixels as the innermost, K-major,
// coordinate it would have to be a multiple of 16 bytes - why this kernel does not transpose anything).
// One CTA = (128 output channels) x (64 or 32 input channels) x (a contiguous chunk of the pixels) x (a group of
// taps): the dz tile of a stage feeds the accumulators of all the CTA's taps (5 or 4 x 64, or 9 x 32 TMEM columns);
// raw sums go to partial[split][9][Cout][Cin], wgrad_reduce_kernel adds the splits in a fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int kWgRows = 32;                 // pixel rows (K) per pipeline stage
constexpr int kWgBM = 128;                  // output channels per tile (UMMA M)
constexpr uint32_t kWgSlabBytes = kWgRows * 128;                           // 32 rows x 32 channels
constexpr uint32_t kWgABytes = (kWgBM / 32) * kWgSlabBytes;                // 4 slabs of dz
// BN input channels per tile: 64 where Cin allows it (each dz tile read from shared memory then feeds twice the
// columns; the nine taps are split over two CTAs, 5 + 4, because 9 x 64 accumulator columns exceed TMEM), else 32.
__host__ __device__ constexpr int wg_groups(int BN) { return BN == 64 ? 2 : 1; }
__host__ __device__ constexpr int wg_max_taps(int BN) { return BN == 64 ? 5 : 9; }
__host__ __device__ constexpr uint32_t wg_stage_bytes(int BN) { return kWgABytes + wg_max_taps(BN) * (BN / 32) * kWgSlabBytes; }
__host__ __device__ constexpr int wg_stages(int BN) { return 3; (void)BN; }   // (4 stages of the 32-channel shape would leave no room for the epilogue staging)

// MN-major TF32 operand.  32-bit MN-major operands exist in one shared-memory layout only: 128-byte rows whose
// 32-byte chunks are XOR-ed with the row index mod 4 (UMMA layout type "128B swizzle, 32-byte base" = TMA swizzle
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  32-channel slabs `lbo` bytes apart, 4-row K groups 512 bytes apart.
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;     // leading byte offset: between slabs along M / N
    d |= (uint64_t)(512 >> 4) << 32;                // stride byte offset: between 4-row groups along K
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                         // SWIZZLE_128B_BASE32B
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int M, int N) {
    return umma_idesc_tf32(M, N) | (1u << 15) | (1u << 16);   // A and B MN-major
}

struct WgradArgs {
    int P, Wp, Cin, Cout;
    int rows_per_split;   // multiple of kWgRows
    int tiles_m, tiles_n, splits;
    float* partial;       // [splits][9][Cout][Cin]
};

// grid: tiles_m * tiles_n * splits * wg_groups(BN) CTAs of 192 threads (warp 0 TMA, warp 1 MMA, warps 2-5 epilogue)
template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x, WgradArgs a) {
    constexpr int kStages = wg_stages(BN), kGroups = wg_groups(BN);
    constexpr uint32_t kStageBytes = wg_stage_bytes(BN);
    constexpr uint32_t kBBytes = (BN / 32) * kWgSlabBytes;   // one tap's x tile: BN / 32 slabs
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_bar;
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x % kGroups, b1 = blockIdx.x / kGroups;
    // (pixel chunks last-first: dz was just written front to back by the BatchNorm backward, its tail is still in L2)
    const int split = a.splits - 1 - b1 % a.splits, t2 = b1 / a.splits;
    const int n0 = (t2 % a.tiles_n) * BN, m0 = (t2 / a.tiles_n) * kWgBM;
    const int tap0 = group == 0 ? 0 : wg_max_taps(BN), ntaps = kGroups == 1 ? 9 : (group == 0 ? 5 : 4);
    const int r0 = split * a.rows_per_split;
    const int r1 = min(r0 + a.rows_per_split, a.P);
    const int iters = r1 > r0 ? (r1 - r0 + kWgRows - 1) / kWgRows : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // up to 9 x 32 or 5 x 64 FP32 accumulator columns -> 512 (power of two)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            // (slabs beyond Cout, when Cout < 128, are not loaded: their accumulator rows are never stored)
            const int slabs = min(kWgBM / 32, (a.Cout - m0) / 32);
            for (int it = 0; it < iters; ++it) {
                const uint32_t s = it % kStages, round = it / kStages;
                mbar_wait(&empty_bar[s], (round & 1) ^ 1);
                unsigned char* sa = tiles + (size_t)s * kStageBytes;
                mbar_expect_tx(&full_bar[s], slabs * kWgSlabBytes + ntaps * kBBytes);
                const int p = r0 + it * kWgRows;
                for (int j = 0; j < slabs; ++j)
                    tma_load_2d(sa + j * kWgSlabBytes, &map_dz, &full_bar[s], m0 + j * 32, p);
                for (int t = 0; t < ntaps; ++t) {
                    const int tap = tap0 + t;
                    const int row = p + (tap / 3 - 1) * a.Wp + (tap % 3 - 1);
#pragma unroll
                    for (int j = 0; j < BN / 32; ++j)
                        tma_load_2d(sa + kWgABytes + t * kBBytes + j * kWgSlabBytes, &map_x, &full_bar[s], n0 + j * 32, row);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32_mn(kWgBM, BN);
            for (int it = 0; it < iters; ++it) {
                const uint32_t s = it % kStages, round = it / kStages;
                mbar_wait(&full_bar[s], round & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(tiles + (size_t)s * kStageBytes);
                const uint64_t adesc = umma_smem_desc_mn(sa, kWgSlabBytes);
                for (int t = 0; t < ntaps; ++t) {
                    const uint64_t bdesc = umma_smem_desc_mn(sa + kWgABytes + t * kBBytes, kWgSlabBytes);
#pragma unroll
                    for (int k = 0; k < kWgRows / 8; ++k)   // UMMA K = 8 pixel rows = 1024 bytes of every slab
                        umma_tf32(tmem_base + t * BN, adesc + (uint64_t)((k * 1024) >> 4), bdesc + (uint64_t)((k * 1024) >> 4),
                                  idesc, (it | k) != 0);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&acc_bar);
        }
    } else {
        const int quarter = warp & 3;
        const int co = m0 + quarter * 32 + lane;
        // (same coalescing as the forward kernel's epilogue: a 32 x 32 block per warp staged in shared memory and read
        // back transposed, so a warp store writes four whole 128-byte lines of the partial-sum tensor)
        float* stage = reinterpret_cast<float*>(tiles + (size_t)kStages * kStageBytes) + (warp - 2) * 32 * kEpiRow;
        if (iters > 0) {
            mbar_wait(&acc_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#pragma unroll 1
        for (int t = 0; t < ntaps * (BN / 32); ++t) {
            const int tl = t / (BN / 32), c0 = (t % (BN / 32)) * 32;
            uint32_t r[32];
            if (iters > 0) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tl * BN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(stage + lane * kEpiRow + j) =
                    make_float4(__int_as_float((int)r[j]), __int_as_float((int)r[j + 1]), __int_as_float((int)r[j + 2]),
                                __int_as_float((int)r[j + 3]));
            __syncwarp();
            {
                const int co_base = m0 + quarter * 32;
                float* obase = a.partial + ((size_t)split * 9 + tap0 + tl) * a.Cout * a.Cin + n0 + c0 + 4 * (lane & 7);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = 4 * it + (lane >> 3);
                    if (co_base + row < a.Cout)
                        *reinterpret_cast<float4*>(obase + (size_t)(co_base + row) * a.Cin) =
                            *reinterpret_cast<const float4*>(stage + row * kEpiRow + 4 * (lane & 7));
                }
            }
            __syncwarp();
        }
        (void)co;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// dW[co][ci][tap] = sum over splits of partial[split][tap][co][ci] (fixed order): the reduction also turns the
// kernel's [tap][Cout][Cin] layout into nn.Conv2d's (Cout, Cin, 3, 3).  One thread per (co, ci): its nine reads are
// coalesced across the warp, its nine outputs are contiguous.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* partial, int splits, int Cout, int Cin, float* dw) {
    // a CTA owns 64 consecutive (co, ci) pairs = 576 contiguous outputs; thread = (pair, tap group of 3 taps)
    __shared__ float stage[64 * 9];
    const long long n = (long long)Cout * Cin;   // multiple of 64 (Cin % 32 == 0, Cout % 32 == 0)
    const int j = threadIdx.x & 63, g = threadIdx.x >> 6;
    for (long long base = (long long)blockIdx.x * 64; base < n; base += (long long)gridDim.x * 64) {
        float acc[3] = {0.0f, 0.0f, 0.0f};
        for (int k = 0; k < splits; ++k)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int t = g + 4 * q;
                if (t < 9) acc[q] += __ldg(partial + ((size_t)k * 9 + t) * n + base + j);
            }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (g + 4 * q < 9) stage[j * 9 + g + 4 * q] = acc[q];
        __syncthreads();
        for (int o = threadIdx.x; o < 64 * 9; o += 256) dw[base * 9 + o] = stage[o];   // coalesced
        __syncthreads();
    }
}

// Weight gradient of the 1-channel first layer: dW[tap][co] = sum_p dz[p][co] * x[p + off(tap)], one streaming pass
// over dz for all nine taps (the layer is memory bound: 4 * Cout bytes per pixel).  Same thread layout and two-stage
// fixed-order reduction as the BatchNorm statistics; partial[chunk][9][C].
__global__ void __launch_bounds__(256) wgrad_cin1_partial_kernel(const float* dz, const float* x, int P, int Wp, int C, float* partial) {
    __shared__ float4 sh[256];
    const StatLayout L(C);
    const int chunk = blockIdx.x;
    const int rows = stat_rows(P);
    const int r0 = chunk * rows, r1 = min(r0 + rows, P);
    for (int g = 0; g * L.cols < (C >> 2); ++g) {
        const int c = (g * L.cols + L.cq) << 2;
        float4 acc[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C && L.rl < L.lanes)
            for (int r = r0 + L.rl; r < r1; r += L.lanes) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(dz + (size_t)r * C + c));
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int q = r + (t / 3 - 1) * Wp + (t % 3 - 1);
                    const float xv = (q >= 0 && q < P) ? __ldg(x + q) : 0.0f;
                    acc[t].x = fmaf(d.x, xv, acc[t].x); acc[t].y = fmaf(d.y, xv, acc[t].y);
                    acc[t].z = fmaf(d.z, xv, acc[t].z); acc[t].w = fmaf(d.w, xv, acc[t].w);
                }
            }
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
            sh[threadIdx.x] = acc[t];
            __syncthreads();
            if (L.rl == 0 && c < C) {
                float4 a = acc[t];
                for (int l = 1; l < L.lanes; ++l) {
                    const float4 u = sh[l * L.cols + L.cq];
                    a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
                }
                *reinterpret_cast<float4*>(partial + ((size_t)chunk * 9 + t) * C + c) = a;
            }
            __syncthreads();
        }
    }
}
// dW[co][0][tap] (Cin = 1): one warp per (tap, co), lanes stride over the chunks in float64
__global__ void wgrad_cin1_final_kernel(const float* partial, int chunks, int C, float* dw) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // tap * C + co
    if (i >= 9 * C) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int k = lane; k < chunks; k += 32) s += partial[(size_t)k * 9 * C + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dw[(i % C) * 9 + i / C] = (float)s;
}

struct WgradPlan { int bn, tiles_m, tiles_n, splits, rows_per_split; };
inline bool wgrad_supported(int Cin, int Cout) { return (Cin % 32 == 0 && Cout % 32 == 0) || (Cin == 1 && Cout % 4 == 0); }
inline WgradPlan wgrad_plan(long long P, int Cin, int Cout) {
    WgradPlan w;
    w.bn = (Cin % 64 == 0) ? 64 : 32;
    w.tiles_m = (Cout + kWgBM - 1) / kWgBM;
    w.tiles_n = Cin / w.bn;
    const int ctas_per_split = w.tiles_m * w.tiles_n * wg_groups(w.bn);
    long long s = (2LL * 148 + ctas_per_split - 1) / ctas_per_split;    // about two CTAs per SM in total
    const long long max_s = (P + 8 * kWgRows - 1) / (8 * kWgRows);      // at least 8 stages per CTA
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    long long rows = (P + s - 1) / s;
    rows = (rows + kWgRows - 1) / kWgRows * kWgRows;
    w.rows_per_split = (int)rows;
    w.splits = (int)((P + rows - 1) / rows);
    return w;
}
inline size_t conv3x3_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    if (!wgrad_supported(Cin, Cout)) return 0;
    if (Cin == 1) {
        const long long P = (long long)B * (H + 2) * (W + 2);
        return (size_t)((P + stat_rows(P) - 1) / stat_rows(P)) * 9 * Cout * sizeof(float);
    }
    const WgradPlan w = wgrad_plan((long long)B * (H + 2) * (W + 2), Cin, Cout);
    return (size_t)w.splits * 9 * Cout * Cin * sizeof(float);
}
// x_padded (B, H+2, W+2, Cin), dz_padded (B, H+2, W+2, Cout) with a zero border -> dw (Cout, Cin, 3, 3)
inline int conv3x3_wgrad(const float* x_padded, const float* dz_padded, float* dw, int B, int H, int W, int Cin, int Cout,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!x_padded || !dz_padded || !dw || !workspace || B <= 0 || H <= 0 || W <= 0) return DMST_EINVAL;
    if (!wgrad_supported(Cin, Cout)) return DMST_EINVAL;
    const long long P = (long long)B * (H + 2) * (W + 2);
    if (P > 0x7fffffffLL) return DMST_EINVAL;
    if (workspace_bytes < conv3x3_wgrad_workspace_bytes(B, H, W, Cin, Cout)) return DMST_EINVAL;
    if (Cin == 1) {
        const int chunks = (int)((P + stat_rows(P) - 1) / stat_rows(P));
        float* partial = reinterpret_cast<float*>(workspace);
        wgrad_cin1_partial_kernel<<<chunks, 256, 0, stream>>>(dz_padded, x_padded, (int)P, W + 2, Cout, partial);
        wgrad_cin1_final_kernel<<<(9 * Cout + 7) / 8, 256, 0, stream>>>(partial, chunks, Cout, dw);
        return (int)cudaGetLastError();
    }
    const WgradPlan w = wgrad_plan(P, Cin, Cout);
    // (a runtime call first: in a thread that has made none yet - an autograd worker - the driver call below would
    // find no current context)
    const size_t smem = (size_t)wg_stages(w.bn) * wg_stage_bytes(w.bn) + 1024 + kEpiBytes;
    int e = w.bn == 64
                ? (int)cudaFuncSetAttribute(conv3x3_wgrad_tf32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                : (int)cudaFuncSetAttribute(conv3x3_wgrad_tf32_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    CUtensorMap mdz, mx;
    e = make_map_2d(&mdz, dz_padded, (uint64_t)P, (uint64_t)Cout, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (e) return e;
    e = make_map_2d(&mx, x_padded, (uint64_t)P, (uint64_t)Cin, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (e) return e;
    WgradArgs a{(int)P, W + 2, Cin, Cout, w.rows_per_split, w.tiles_m, w.tiles_n, w.splits, reinterpret_cast<float*>(workspace)};
    const int grid = w.tiles_m * w.tiles_n * w.splits * wg_groups(w.bn);
    if (w.bn == 64) conv3x3_wgrad_tf32_kernel<64><<<grid, kConvThreads, smem, stream>>>(mdz, mx, a);
    else conv3x3_wgrad_tf32_kernel<32><<<grid, kConvThreads, smem, stream>>>(mdz, mx, a);
    wgrad_reduce_kernel<<<grid_for((long long)Cout * Cin * 4), 256, 0, stream>>>(a.partial, w.splits, Cout, Cin, dw);  // (Cout * Cin) / 64 CTAs at most
    return (int)cudaGetLastError();
}

// k-splits for a layer: only when one wave of tiles would leave SMs idle (the small deep layers)
inline int conv_ksplit(long long P, int Cin, int Cout, int sms) {
    if (Cin % kConvBK != 0 || Cout % 64 != 0) return 1;
    const int BN = (Cout % 128 == 0) ? 128 : 64;
    const long long tiles1 = ((P + kConvBM - 1) / kConvBM) * (Cout / BN);
    if (tiles1 > sms) return 1;
    const int iters = 3 * (Cin / kConvBK);
    long long s = (2LL * sms) / tiles1;
    if (s > 8) s = 8;
    if (s > iters / 2) s = iters / 2;
    return s < 1 ? 1 : (int)s;
}
inline size_t conv3x3_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    const long long P = (long long)B * (H + 2) * (W + 2);
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int s = conv_ksplit(P, Cin, Cout, sms);
    return s > 1 ? (size_t)s * P * Cout * sizeof(float) : 0;
}

inline int conv3x3_forward(const float* x_padded, const float* w9, const float* scale, const float* shift, float* y_padded,
                           int B, int H, int W, int Cin, int Cout, int relu, cudaStream_t stream,
                           void* workspace = nullptr, size_t workspace_bytes = 0) {
    if (!x_padded || !w9 || !y_padded || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return DMST_EINVAL;
    const int Hp = H + 2, Wp = W + 2;
    const long long P = (long long)B * Hp * Wp;
    if (P > 0x7fffffffLL) return DMST_EINVAL;
    if (Cin <= kSmallCinMax && Cout % 4 == 0 && Cout <= 256 && 256 % (Cout / 4) == 0) {   // first layer: streaming kernel
        const int pix_per_block = 256 / (Cout / 4);
        const long long blocks_needed = (P + pix_per_block - 1) / pix_per_block;
        const int blocks = (int)(blocks_needed > 148LL * 16 ? 148 * 16 : blocks_needed);
        if (Cin == 1) {
            conv3x3_cin1_kernel<<<blocks, 256, 0, stream>>>(x_padded, w9, scale, shift, y_padded, (int)P, Hp, Wp, Cout, relu);
            return (int)cudaGetLastError();
        }
        const size_t smem = (size_t)(9 * Cin * Cout + 2 * Cout) * sizeof(float);
        conv3x3_small_cin_kernel<<<blocks, 256, smem, stream>>>(x_padded, w9, scale, shift, y_padded, (int)P, Hp, Wp, Cin, Cout, relu);
        return (int)cudaGetLastError();
    }
    if (Cin % kConvBK != 0 || Cout % 64 != 0) {   // small / odd channel counts: CUDA-core path
        const long long total = P * Cout;
        const int blocks = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
        conv3x3_direct_kernel<<<blocks, 256, 0, stream>>>(x_padded, w9, scale, shift, y_padded, (int)P, Hp, Wp, Cin, Cout, relu);
        return (int)cudaGetLastError();
    }
    const int BN = (Cout % 128 == 0) ? 128 : 64;
    CUtensorMap ma, mb;
    int e = make_map_2d(&ma, x_padded, (uint64_t)P, (uint64_t)Cin, kConvARows);
    if (e) return e;
    e = make_map_2d(&mb, w9, (uint64_t)9 * Cout, (uint64_t)Cin, (uint32_t)BN);
    if (e) return e;
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int ksplit = conv_ksplit(P, Cin, Cout, sms);
    if (ksplit > 1 && (!workspace || workspace_bytes < (size_t)ksplit * P * Cout * sizeof(float))) ksplit = 1;  // no workspace: unsplit
    ConvArgs a{(int)P, Hp, Wp, Cin, Cout, scale, shift, relu, y_padded, ksplit, reinterpret_cast<float*>(workspace)};
    // two pixel sub-tiles per CTA tile when that still leaves at least two waves of tiles
    const long long tiles1 = ((P + kConvBM - 1) / kConvBM) * (Cout / BN);
    const int MT = (tiles1 >= 4LL * sms) ? 2 : 1;
    const long long num_tiles = ((P + MT * kConvBM - 1) / (MT * kConvBM)) * (Cout / BN) * ksplit;
    const dim3 grid((unsigned)(num_tiles < sms ? num_tiles : sms));  // persistent: one CTA per SM
    auto launch = [&](auto kern, int bn, int mt) -> int {
        const size_t smem = (size_t)conv_stages(bn, mt) * conv_stage_bytes(bn, mt) + 1024 + kEpiBytes;
        int err = (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err) return err;
        kern<<<grid, kConvThreads, smem, stream>>>(ma, mb, a);
        return 0;
    };
    if (BN == 128) e = (MT == 2) ? launch(conv3x3_tf32_kernel<128, 2>, 128, 2) : launch(conv3x3_tf32_kernel<128, 1>, 128, 1);
    else e = (MT == 2) ? launch(conv3x3_tf32_kernel<64, 2>, 64, 2) : launch(conv3x3_tf32_kernel<64, 1>, 64, 1);
    if (e) return e;
    if (ksplit > 1)
        conv_splitk_reduce_kernel<<<grid_for(P * (Cout / 4)), 256, 0, stream>>>(a.partial, ksplit, y_padded, (int)P, Hp, Wp, Cout,
                                                                             scale, shift, relu);
    return (int)cudaGetLastError();
}

}  // namespace dmst
#endif
