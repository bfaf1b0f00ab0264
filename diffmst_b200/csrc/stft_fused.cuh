// Fused STFT kernels of the multi-resolution STFT loss, with their own FFT (no library call), for power-of-two FFT
// sizes 64..8192 (auraloss.freq.STFTLoss.stft + the three loss terms and their gradient, SURVEY.md Appendix B;
// instantiated at configs/models/naive.yaml:54-68, called at mst/system.py:332):
//
//   stft_loss_kernel   framing (reflect padding, window), the real FFT of every frame of BOTH signals and the loss
//                      reductions; replaces frame4_kernel + cuFFT R2C + mr_loss_kernel.
//   istft_grad_kernel  spectrum of the prediction + max(|Y|^2, eps) -> half-spectrum gradient -> inverse real FFT of
//                      every frame; replaces mr_grad_kernel + cuFFT C2R.
//
// Why: the library path writes the frames (2x the signal), reads them back, writes both spectra and reads them again:
// 300 MB of DRAM traffic per resolution at the headline shape for 84 MB of necessary traffic (read the signals once
// through L2, write the spectrum of the prediction and |Y|^2 of the target for the gradient pass).  Here a frame never
// leaves the SM between the signal and the loss sums.
//
// Layout: a block of 256 threads owns 4096 complex points = FPB = 8192 / n frames of one row; a real FFT of
// length n is a complex FFT of length M = n/2 over z[m] = x[2m] + i x[2m+1] followed by the even/odd split.
// The complex FFT is a Cooley-Tukey decomposition M = 16 * R2 [* R3]: every thread holds 16 points in
// registers per stage (16 / R butterflies of radix R), stages exchange through shared memory IN PLACE (a
// butterfly reads and writes the same R addresses, so one barrier per stage), the index padding i + i/16 + i/256
// keeps all strides that occur free of bank conflicts.  Twiddles come from float64-built tables (w_M^j and w_n^k).
// The inverse transform is the same stage code run on the conjugated input, result conjugated.
#pragma once
#include "common.cuh"
#include "stft.cuh"

namespace dmst {

constexpr int kSfThreads = 256;
constexpr int kSfPoints = 4096;   // complex points per block and signal
// Index padding of the shared-memory exchange buffers: one float2 every 16 and every 256.  Every access pattern
// of the stages (strides 1, 4, 16, M/16, 256) then hits 16 different 8-byte banks per half-warp; the digit-reversed
// read-out costs 1.5 wavefronts instead of up to 15 (checked by enumeration for all M).
__host__ __device__ constexpr int sf_pad(int i) { return i + (i >> 4) + (i >> 8); }
constexpr int kSfSmemFloat2 = kSfPoints + (kSfPoints >> 4) + (kSfPoints >> 8);
// sf_pad(q L + n_low + n S) = sf_base<L, S>(q, n_low) + sf_off<L, S>(n) for sub-FFTs of length L = R S (n_low < S):
// the per-element part is a compile-time constant (enumerated for every (L, S) the plans use)
template <int L, int S>
__device__ __forceinline__ int sf_base(int q, int n_low) {
    return sf_pad(q * L) + n_low + (S >= 16 ? (n_low >> 4) : 0) + (S >= 256 ? (n_low >> 8) : 0);
}
template <int L, int S>
__host__ __device__ constexpr int sf_off(int n) {
    return n * S + (L < 16 ? 0 : (S >= 16 ? n * (S >> 4) : ((n * S) >> 4))) + (L < 256 ? 0 : (S >= 256 ? n * (S >> 8) : ((n * S) >> 8)));
}

__device__ __forceinline__ float2 sf_cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 sf_ldg2(const float2* p) {
#ifdef DMST_EMULATE
    return *p;
#else
    return __ldg(p);
#endif
}

// ---- register DFTs, forward sign exp(-2 pi i / R) -------------------------------------------------------
__device__ __forceinline__ void sf_dft2(float2& a, float2& b) {
    const float2 s = make_float2(a.x + b.x, a.y + b.y), d = make_float2(a.x - b.x, a.y - b.y);
    a = s; b = d;
}
// natural order in, natural order out
__device__ __forceinline__ void sf_dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y), t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
    const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y), t3 = make_float2(a1.x - a3.x, a1.y - a3.y);
    a0 = make_float2(t0.x + t2.x, t0.y + t2.y);
    a2 = make_float2(t0.x - t2.x, t0.y - t2.y);
    a1 = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i t3
    a3 = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i t3
}
constexpr float kSfC1 = 0.92387953251128674f, kSfS1 = 0.38268343236508977f, kSfR = 0.70710678118654752f;
// a * w16^J for the J that occur
template <int J>
__device__ __forceinline__ float2 sf_mul_w16(float2 a) {
    if constexpr (J == 0) return a;
    else if constexpr (J == 1) return sf_cmul(a, make_float2(kSfC1, -kSfS1));
    else if constexpr (J == 2) return make_float2((a.x + a.y) * kSfR, (a.y - a.x) * kSfR);
    else if constexpr (J == 3) return sf_cmul(a, make_float2(kSfS1, -kSfC1));
    else if constexpr (J == 4) return make_float2(a.y, -a.x);
    else if constexpr (J == 6) return make_float2((a.y - a.x) * kSfR, -(a.x + a.y) * kSfR);
    else { static_assert(J == 9, "twiddle"); return sf_cmul(a, make_float2(-kSfC1, kSfS1)); }
}
// In-place DFT of R points held as v[0..R).  Result X[k] sits at v[DftPos<R>::of(k)].
template <int R> struct SfDft;
template <> struct SfDft<2> {
    __device__ static __forceinline__ void run(float2* v) { sf_dft2(v[0], v[1]); }
    __host__ __device__ static constexpr int pos(int k) { return k; }
};
template <> struct SfDft<4> {
    __device__ static __forceinline__ void run(float2* v) { sf_dft4(v[0], v[1], v[2], v[3]); }
    __host__ __device__ static constexpr int pos(int k) { return k; }
};
template <> struct SfDft<8> {   // n = 4 a + b: DFT2 over a, twiddle w8^(b k1), DFT4 over b; X[k1 + 2 k2] at 4 k1 + k2
    __device__ static __forceinline__ void run(float2* v) {
        sf_dft2(v[0], v[4]); sf_dft2(v[1], v[5]); sf_dft2(v[2], v[6]); sf_dft2(v[3], v[7]);
        v[5] = sf_mul_w16<2>(v[5]); v[6] = sf_mul_w16<4>(v[6]); v[7] = sf_mul_w16<6>(v[7]);
        sf_dft4(v[0], v[1], v[2], v[3]);
        sf_dft4(v[4], v[5], v[6], v[7]);
    }
    __host__ __device__ static constexpr int pos(int k) { return 4 * (k & 1) + (k >> 1); }
};
template <> struct SfDft<16> {  // n = 4 a + b: DFT4 over a, twiddle w16^(b k1), DFT4 over b; X[k1 + 4 k2] at 4 k1 + k2
    __device__ static __forceinline__ void run(float2* v) {
        sf_dft4(v[0], v[4], v[8], v[12]); sf_dft4(v[1], v[5], v[9], v[13]);
        sf_dft4(v[2], v[6], v[10], v[14]); sf_dft4(v[3], v[7], v[11], v[15]);
        v[5] = sf_mul_w16<1>(v[5]);   v[6] = sf_mul_w16<2>(v[6]);   v[7] = sf_mul_w16<3>(v[7]);
        v[9] = sf_mul_w16<2>(v[9]);   v[10] = sf_mul_w16<4>(v[10]); v[11] = sf_mul_w16<6>(v[11]);
        v[13] = sf_mul_w16<3>(v[13]); v[14] = sf_mul_w16<6>(v[14]); v[15] = sf_mul_w16<9>(v[15]);
        sf_dft4(v[0], v[1], v[2], v[3]);     sf_dft4(v[4], v[5], v[6], v[7]);
        sf_dft4(v[8], v[9], v[10], v[11]);   sf_dft4(v[12], v[13], v[14], v[15]);
    }
    __host__ __device__ static constexpr int pos(int k) { return 4 * (k & 3) + (k >> 2); }
};

// Radix plan of the complex length M = 16 * R2 [* R3]
template <int M> struct SfPlan {
    static_assert(M >= 32 && M <= 4096 && (M & (M - 1)) == 0, "complex FFT length");
    static constexpr int L2 = M / 16;                        // sub-FFT length after the first stage
    static constexpr int R2 = L2 >= 16 ? 16 : L2;
    static constexpr int L3 = L2 / R2;                       // 1: two stages
    static constexpr int R3 = L3;
    static constexpr int FPB = kSfPoints / M;                // frames per block
    // position (inside the frame's M points) of frequency index k after the last stage
    __host__ __device__ static constexpr int pos(int k) {
        return (k & 15) * L2 + ((k >> 4) & (R2 - 1)) * L3 + (k >> 4) / R2;
    }
};

// One in-place shared-memory stage over both signals (sm0: prediction, sm1: target): sub-FFTs of length L = R * S;
// every thread does 16 / R butterflies per signal.  Output k is multiplied by w_L^(n_low k) = tw[(M / L) n_low k]
// (loaded once, ahead of the butterflies, and used for both signals).
template <int M, int R, int L, int NSIG = 2>
__device__ __forceinline__ void sf_stage(float2* sm0, float2* sm1, const float2* tw, int tid) {
    constexpr int S = L / R, G = 16 / R;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const int g = tid + kSfThreads * j;
        const int q = g / S, n_low = g - q * S;
        const int base = sf_base<L, S>(q, n_low);
        float2 w[R];
        if (S > 1) {
#pragma unroll
            for (int k = 1; k < R; ++k) w[k] = sf_ldg2(tw + (M / L) * n_low * k);
        }
#pragma unroll
        for (int sig = 0; sig < NSIG; ++sig) {
            float2* sm = sig ? sm1 : sm0;
            float2 v[R];
#pragma unroll
            for (int n = 0; n < R; ++n) v[n] = sm[base + sf_off<L, S>(n)];
            SfDft<R>::run(v);
#pragma unroll
            for (int k = 0; k < R; ++k) {
                float2 o = v[SfDft<R>::pos(k)];
                if (S > 1 && k > 0) o = sf_cmul(o, w[k]);
                sm[base + sf_off<L, S>(k)] = o;
            }
        }
    }
}

// Second stage of the loss reductions of one resolution (mrstft.cuh: mr_final, run by the block that finishes last)
struct MrFinalArgs {
    const float* partial;
    int blocks_per_row;
    double* rowsum;     // [rows][4] scratch
    int rows, per_row;
    float w_sc, w_log, w_lin;
    int n_res;
    float* res_loss;    // [4]: this resolution's contribution to the total, then its sc, log, lin terms
    float* row_coef;    // [rows]: d(total)/d|X| coefficient of (|X|-|Y|) for the SC term
    float* scal;        // [2]: coefficient of sign(log) / |X| and of sign(lin)
};
#ifndef DMST_EMULATE
__device__ void mr_final(const MrFinalArgs& a);
#endif

struct SfArgs {
    const float* x[2];          // prediction, target: rows x T with row strides
    long long row_stride[2];
    int vec_ok[2];              // 8-byte aligned rows (base and stride) for the float2 interior path
    int rows, T, n, hop, win, frames;
    const float* window;        // win floats
    int win_vec_ok;             // win == n and 8-byte aligned
    const float2* tw_m;         // [M]      exp(-2 pi i j / M)
    const float2* tw_n;         // [M/2+1]  exp(-2 pi i k / n)
    float2* X;                  // rows x frames x (M+1) spectrum of the prediction, or null (loss only)
    float* PY;                  // rows x frames x (M+1) max(|Y|^2, eps), or null
    float eps;
    float* partial;             // [rows][gridDim.x][4]: sum (|Y|-|X|)^2, sum |Y|^2, sum |log|X| - log|Y||, sum ||Y|-|X||
    unsigned* done;             // blocks finished (zero at launch; the last block resets it), or null: no second stage
    MrFinalArgs fin;
};

__device__ __forceinline__ int sf_reflect(int t, int T) {
    if (t < 0) t = -t;
    if (t >= T) t = 2 * T - 2 - t;
    return t;
}

// Loads the block's FPB frames of both signals and runs the complex FFTs; on return (after a barrier) sm0 / sm1
// hold Z of every frame of the prediction / the target, frequency k of frame p at sf_pad(p * M + SfPlan<M>::pos(k)).
template <int M>
__device__ __forceinline__ void sf_fft_frames(const SfArgs& a, int row, int frame0, float2* sm0, float2* sm1, int tid) {
    typedef SfPlan<M> P;
    constexpr int S = M / 16;
    const int p = tid / S, n_low = tid - p * S;
    const int f = frame0 + p;
    const int sbase = sf_base<M, S>(p, n_low);
    float2 tw[16];
#pragma unroll
    for (int k = 1; k < 16; ++k) tw[k] = sf_ldg2(a.tw_m + n_low * k);
    if (f < a.frames) {
        const int base = f * a.hop - (a.n >> 1);   // signal index of the frame's sample 0
        const int wl = (a.n - a.win) >> 1;
        const bool inside = base >= 0 && base + a.n <= a.T && (base & 1) == 0;
        // the window of this thread's 16 points serves both signals
        float2 w[16];
        if (a.win_vec_ok) {
#pragma unroll
            for (int n = 0; n < 16; ++n) w[n] = sf_ldg2(reinterpret_cast<const float2*>(a.window + 2 * (n * S + n_low)));
        } else {
#pragma unroll
            for (int n = 0; n < 16; ++n) {
                const int i0 = 2 * (n * S + n_low) - wl, i1 = i0 + 1;
                w[n] = make_float2((i0 >= 0 && i0 < a.win) ? __ldg(a.window + i0) : 0.0f,
                                   (i1 >= 0 && i1 < a.win) ? __ldg(a.window + i1) : 0.0f);
            }
        }
#pragma unroll
        for (int sig = 0; sig < 2; ++sig) {
            const float* xrow = (sig ? a.x[1] : a.x[0]) + (long long)row * (sig ? a.row_stride[1] : a.row_stride[0]);
            const bool vec = (sig ? a.vec_ok[1] : a.vec_ok[0]) != 0;
            float2 v[16];
            if (inside && vec) {
#pragma unroll
                for (int n = 0; n < 16; ++n) v[n] = sf_ldg2(reinterpret_cast<const float2*>(xrow + base + 2 * (n * S + n_low)));
            } else {
#pragma unroll
                for (int n = 0; n < 16; ++n) {
                    const int i = base + 2 * (n * S + n_low);
                    v[n] = make_float2(__ldg(xrow + sf_reflect(i, a.T)), __ldg(xrow + sf_reflect(i + 1, a.T)));
                }
            }
#pragma unroll
            for (int n = 0; n < 16; ++n) { v[n].x *= w[n].x; v[n].y *= w[n].y; }
            SfDft<16>::run(v);
            float2* sm = sig ? sm1 : sm0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float2 o = v[SfDft<16>::pos(k)];
                if (k > 0) o = sf_cmul(o, tw[k]);
                sm[sbase + sf_off<M, S>(k)] = o;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            sm0[sbase + sf_off<M, S>(k)] = make_float2(0.0f, 0.0f);
            sm1[sbase + sf_off<M, S>(k)] = make_float2(0.0f, 0.0f);
        }
    }
    __syncthreads();
    sf_stage<M, P::R2, P::L2>(sm0, sm1, a.tw_m, tid);
    __syncthreads();
    if (P::L3 > 1) {
        sf_stage<M, (P::R3 > 1 ? P::R3 : 2), (P::L3 > 1 ? P::L3 : 2)>(sm0, sm1, a.tw_m, tid);
        __syncthreads();
    }
}

// Spectrum bins of slot k (0 <= k < M/2) of one frame from its Z (zf = the frame's first point in shared memory),
// given w = w_n^k: k >= 1 gives the pair (k, M - k); k = 0 gives bins 0 and M (out[0], out[1]) and bin M/2 (out[2]).
template <int M>
__device__ __forceinline__ void sf_split(const float2* zf, float2 w, int k, float2 (&out)[3]) {
    typedef SfPlan<M> P;
    if (k == 0) {
        const float2 z0 = zf[sf_pad(P::pos(0))], zh = zf[sf_pad(P::pos(M / 2))];
        out[0] = make_float2(z0.x + z0.y, 0.0f);
        out[1] = make_float2(z0.x - z0.y, 0.0f);
        out[2] = make_float2(zh.x, -zh.y);   // Xe + (-i) Xo with Xe = re(z), Xo = im(z)
        return;
    }
    const float2 za = zf[sf_pad(P::pos(k))], zb = zf[sf_pad(P::pos(M - k))];
    // Xe = (za + conj(zb)) / 2, Xo = (za - conj(zb)) / (2 i)
    const float2 xe = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
    const float2 xo = make_float2(0.5f * (za.y + zb.y), -0.5f * (za.x - zb.x));
    const float2 pw = sf_cmul(xo, w);
    out[0] = make_float2(xe.x + pw.x, xe.y + pw.y);
    out[1] = make_float2(xe.x - pw.x, -(xe.y - pw.y));
    out[2] = make_float2(0.0f, 0.0f);
}

#ifndef DMST_EMULATE
__device__ __forceinline__ float sf_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sf_log2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
__device__ __forceinline__ float sf_rsqrt(float x) { return 1.0f / std::sqrt(x); }
__device__ __forceinline__ float sf_log2(float x) { return std::log2(x); }
#endif

struct SfSums { float d2, y2, lg, lin; };
__device__ __forceinline__ void sf_accumulate(SfSums& s, float px, float py) {
    // (__fmul_rn: no contraction into an FMA, so that |Y| - |X| is exactly 0 for identical spectra)
    const float d = __fmul_rn(py, sf_rsqrt(py)) - __fmul_rn(px, sf_rsqrt(px));
    s.d2 = fmaf(d, d, s.d2);
    s.y2 += py;
    s.lg += fabsf(sf_log2(px) - sf_log2(py));
    s.lin += fabsf(d);
}

// Block-level part of the kernel: FFTs of both signals, spectrum / |Y|^2 stores, the four sums of the block
// (valid in warp 0 on return).
template <int M>
__device__ __forceinline__ SfSums sf_block(const SfArgs& a, float2* sm0, float2* sm1, float* red) {
    typedef SfPlan<M> P;
    constexpr int FPB = P::FPB, HALF = M / 2, SLOTS = FPB * HALF, J = SLOTS / kSfThreads, BINS = M + 1;
    static_assert(SLOTS % kSfThreads == 0, "slots per thread");
    const int tid = threadIdx.x, row = blockIdx.y;
    const int frame0 = blockIdx.x * FPB;
    sf_fft_frames<M>(a, row, frame0, sm0, sm1, tid);
    SfSums s{0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int slot = tid + kSfThreads * j;
        const int p = slot / HALF, k = slot - p * HALF;
        const int f = frame0 + p;
        if (f < a.frames) {
            const float2 w = sf_ldg2(a.tw_n + k);
            float2 ox[3], oy[3];
            const int fb = sf_pad(p * M);
            sf_split<M>(sm0 + fb, w, k, ox);
            sf_split<M>(sm1 + fb, w, k, oy);
            const int b0 = k, b1 = M - k;   // (k = 0: bins 0 and M, and M/2 as the third)
            const float py0 = fmaxf(fmaf(oy[0].x, oy[0].x, oy[0].y * oy[0].y), a.eps);
            const float py1 = fmaxf(fmaf(oy[1].x, oy[1].x, oy[1].y * oy[1].y), a.eps);
            sf_accumulate(s, fmaxf(fmaf(ox[0].x, ox[0].x, ox[0].y * ox[0].y), a.eps), py0);
            sf_accumulate(s, fmaxf(fmaf(ox[1].x, ox[1].x, ox[1].y * ox[1].y), a.eps), py1);
            float py2 = a.eps;
            if (k == 0) {
                py2 = fmaxf(fmaf(oy[2].x, oy[2].x, oy[2].y * oy[2].y), a.eps);
                sf_accumulate(s, fmaxf(fmaf(ox[2].x, ox[2].x, ox[2].y * ox[2].y), a.eps), py2);
            }
            if (a.X) {
                const long long o = ((long long)row * a.frames + f) * BINS;
                a.X[o + b0] = ox[0]; a.X[o + b1] = ox[1];
                a.PY[o + b0] = py0; a.PY[o + b1] = py1;
                if (k == 0) { a.X[o + HALF] = ox[2]; a.PY[o + HALF] = py2; }
            }
        }
    }
    s.lg *= 0.5f * 0.6931471805599453f;   // log|X| - log|Y| = ln2 / 2 * (log2 px - log2 py)
    SfSums r;
    r.d2 = block_sum(s.d2, red); r.y2 = block_sum(s.y2, red); r.lg = block_sum(s.lg, red); r.lin = block_sum(s.lin, red);
    return r;
}

// grid: (ceil(frames / FPB), rows); block 256
template <int M>
__global__ void __launch_bounds__(kSfThreads, 2) stft_loss_kernel(SfArgs a) {
    DMST_DYN_SMEM(smem_raw);
    float2* sm0 = reinterpret_cast<float2*>(smem_raw);
    float2* sm1 = sm0 + kSfSmemFloat2;
    DMST_SHARED_ARRAY(float, red, 33);
    const SfSums r = sf_block<M>(a, sm0, sm1, red);
    if (threadIdx.x == 0) {
        float* out = a.partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 4;
        out[0] = r.d2; out[1] = r.y2; out[2] = r.lg; out[3] = r.lin;
    }
#ifndef DMST_EMULATE
    if (!a.done) return;
    // the block that finishes last reduces the partials (threadFenceReduction pattern)
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned prev = atomicAdd(a.done, 1u);
        red[32] = (prev == total - 1) ? 1.0f : 0.0f;
        if (prev == total - 1) { *a.done = 0u; __threadfence(); }   // ready for the next launch
    }
    __syncthreads();
    if (red[32] != 0.0f) mr_final(a.fin);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// Gradient side: spectrum of the prediction + max(|Y|^2, eps) -> half-spectrum gradient -> inverse real FFT of every
// frame, in one kernel (replaces mr_grad_kernel + cuFFT C2R for the same FFT sizes).  The loss's derivative w.r.t. the
// frame is r[n] = Re sum_{k=0..n/2} g[k] e^{+2 pi i k n / N}, g = dL/dRe X + i dL/dIm X; with H = g/2 on interior bins
// (real part of g at DC and Nyquist) r is the unnormalised inverse DFT of H's Hermitian extension, which a complex
// inverse FFT of length M = N/2 delivers as z'[m] = r[2m] + i r[2m+1] from
//     Z'[k] = (H[k] + conj(H[M-k])) + i conj(w_N^k) (H[k] - conj(H[M-k])),      k = 0 .. M-1   (H[M]: Nyquist)
// and the inverse FFT is the forward machinery on conj(Z') with the result conjugated.
// ---------------------------------------------------------------------------------------------------------------
struct SfGradArgs {
    const float2* X;            // rows x frames x (M+1) spectrum of the prediction
    const float* PY;            // rows x frames x (M+1) max(|Y|^2, eps)
    float* dframes;             // rows x frames x n: d loss / d frame (before the window; the overlap-add applies it)
    int rows, n, frames;
    const float2* tw_m;
    const float2* tw_n;
    float eps;
    const float* row_coef;      // [rows]
    const float* scal;          // [2]
    int use_log, use_lin;
};

// g/2 (interior) of one bin: x = X[bin], py = max(|Y|^2, eps)
__device__ __forceinline__ float2 sf_bin_gradient(float2 x, float py, float rc, float c_log, float c_lin, float eps) {
    const float px = fmaf(x.x, x.x, x.y * x.y);
    if (!(px >= eps)) return make_float2(0.0f, 0.0f);   // the clamp passes gradient only where it is inactive
    const float rx = sf_rsqrt(px);
    const float d = __fmul_rn(px, rx) - __fmul_rn(py, sf_rsqrt(py));   // |X| - |Y|
    const float sg_log = (px > py) ? 1.0f : ((px < py) ? -1.0f : 0.0f);
    const float sg_lin = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
    const float s = fmaf(rc, d, fmaf(c_log * sg_log, rx, c_lin * sg_lin)) * rx;
    return make_float2(s * x.x, s * x.y);
}

// grid: (ceil(frames / FPB), rows); block 256; one exchange buffer
template <int M>
__global__ void __launch_bounds__(kSfThreads, 4) istft_grad_kernel(SfGradArgs a) {
    typedef SfPlan<M> P;
    constexpr int FPB = P::FPB, HALF = M / 2, SLOTS = FPB * HALF, J = SLOTS / kSfThreads, BINS = M + 1, S = M / 16;
    DMST_DYN_SMEM(smem_raw);
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int tid = threadIdx.x, row = blockIdx.y;
    const int frame0 = blockIdx.x * FPB;
    const float rc = __ldg(a.row_coef + row), c_log = a.use_log ? __ldg(a.scal) : 0.0f, c_lin = a.use_lin ? __ldg(a.scal + 1) : 0.0f;
    // ---- conj(Z') of every frame into shared memory, natural order ----
    // (all global loads of the thread's slots are issued before the first use: one memory latency, not one per slot)
    float2 xa[J], xb[J];
    float pa[J], pb[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int slot = tid + kSfThreads * j;
        const int p = slot / HALF, k = slot - p * HALF;
        const int f = frame0 + p;
        xa[j] = make_float2(0.0f, 0.0f); xb[j] = xa[j]; pa[j] = 1.0f; pb[j] = 1.0f;
        if (f < a.frames) {   // (k = 0: bins 0 and M, the same addresses as the pair (k, M - k))
            const long long o = ((long long)row * a.frames + f) * BINS;
            xa[j] = sf_ldg2(a.X + o + k); xb[j] = sf_ldg2(a.X + o + M - k);
            pa[j] = __ldg(a.PY + o + k); pb[j] = __ldg(a.PY + o + M - k);
        }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int slot = tid + kSfThreads * j;
        const int p = slot / HALF, k = slot - p * HALF;
        const int f = frame0 + p;
        float2* zf = sm + sf_pad(p * M);
        // (frames beyond the end: x = 0 gives zero gradients, so zeros are stored)
        const float2 ga = sf_bin_gradient(xa[j], pa[j], rc, c_log, c_lin, a.eps);
        const float2 gb = sf_bin_gradient(xb[j], pb[j], rc, c_log, c_lin, a.eps);
        if (k == 0) {
            // DC and Nyquist keep the real part of g; bin M/2 is an interior bin paired with itself
            float2 gh = make_float2(0.0f, 0.0f);
            if (f < a.frames) {
                const long long o = ((long long)row * a.frames + f) * BINS;
                gh = sf_bin_gradient(sf_ldg2(a.X + o + HALF), __ldg(a.PY + o + HALF), rc, c_log, c_lin, a.eps);
            }
            zf[sf_pad(0)] = make_float2(ga.x + gb.x, -(ga.x - gb.x));   // conj((h0 + hm) + i (h0 - hm))
            zf[sf_pad(HALF)] = gh;                                       // conj(2 conj(H)) with H = g / 2
        } else {
            const float2 A = make_float2(0.5f * ga.x, 0.5f * ga.y), B = make_float2(0.5f * gb.x, -0.5f * gb.y);   // H[k], conj(H[M-k])
            const float2 sum = make_float2(A.x + B.x, A.y + B.y), dif = make_float2(A.x - B.x, A.y - B.y);
            const float2 w = sf_ldg2(a.tw_n + k);                       // w_N^k
            // i conj(w) dif  and  i w conj(dif)
            const float2 cw = sf_cmul(make_float2(w.x, -w.y), dif), wc = sf_cmul(w, make_float2(dif.x, -dif.y));
            const float2 zk = make_float2(sum.x - cw.y, sum.y + cw.x);         // Z'[k]
            const float2 zm = make_float2(sum.x - wc.y, -sum.y + wc.x);        // Z'[M-k] = conj(sum) + i w conj(dif)
            zf[sf_pad(k)] = make_float2(zk.x, -zk.y);
            zf[sf_pad(M - k)] = make_float2(zm.x, -zm.y);
        }
    }
    __syncthreads();
    // ---- forward FFT machinery on conj(Z'), in place ----
    sf_stage<M, 16, M, 1>(sm, sm, a.tw_m, tid);
    __syncthreads();
    sf_stage<M, P::R2, P::L2, 1>(sm, sm, a.tw_m, tid);
    __syncthreads();
    if (P::L3 > 1) {
        sf_stage<M, (P::R3 > 1 ? P::R3 : 2), (P::L3 > 1 ? P::L3 : 2), 1>(sm, sm, a.tw_m, tid);
        __syncthreads();
    }
    // ---- r[2m] + i r[2m+1] = conj(out[m]): coalesced float2 stores ----
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int idx = tid + kSfThreads * j;
        const int p = idx / M, m = idx - p * M;
        const int f = frame0 + p;
        if (f < a.frames) {
            const float2 o = sm[sf_pad(p * M) + sf_pad(P::pos(m))];
            *reinterpret_cast<float2*>(a.dframes + ((long long)row * a.frames + f) * a.n + 2 * m) = make_float2(o.x, -o.y);
        }
    }
    (void)S;
}

constexpr size_t kSfSmemBytes = 2 * sizeof(float2) * kSfSmemFloat2;
#ifdef DMST_EMULATE
#define DMST_SF_SET_SMEM(kernel) 0
#else
#define DMST_SF_SET_SMEM(kernel) \
    (int)cudaFuncSetAttribute((kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSfSmemBytes)
#endif
// Host dispatch over the complex length M = n / 2; false when n is not served by the fused kernel
inline bool sf_supported(int n) { return n >= 64 && n <= 8192 && (n & (n - 1)) == 0; }
inline bool sf_launch(const SfArgs& a, cudaStream_t stream) {
    const dim3 block(kSfThreads);
#define DMST_SF_CASE(MM)                                                                                   \
    case MM: {                                                                                             \
        const dim3 grid((a.frames + SfPlan<MM>::FPB - 1) / SfPlan<MM>::FPB, a.rows);                       \
        if (DMST_SF_SET_SMEM(stft_loss_kernel<MM>) != 0) return false;                                     \
        DMST_LAUNCH(stft_loss_kernel<MM>, grid, block, kSfSmemBytes, stream, a);                           \
        return true;                                                                                       \
    }
    switch (a.n / 2) {
        DMST_SF_CASE(32) DMST_SF_CASE(64) DMST_SF_CASE(128) DMST_SF_CASE(256) DMST_SF_CASE(512)
        DMST_SF_CASE(1024) DMST_SF_CASE(2048) DMST_SF_CASE(4096)
        default: return false;
    }
#undef DMST_SF_CASE
}
inline bool sf_grad_launch(const SfGradArgs& a, cudaStream_t stream) {
    const dim3 block(kSfThreads);
    constexpr size_t smem = sizeof(float2) * kSfSmemFloat2;
#define DMST_SF_CASE(MM)                                                                                   \
    case MM: {                                                                                             \
        const dim3 grid((a.frames + SfPlan<MM>::FPB - 1) / SfPlan<MM>::FPB, a.rows);                       \
        DMST_LAUNCH(istft_grad_kernel<MM>, grid, block, smem, stream, a);                                  \
        return true;                                                                                       \
    }
    switch (a.n / 2) {
        DMST_SF_CASE(32) DMST_SF_CASE(64) DMST_SF_CASE(128) DMST_SF_CASE(256) DMST_SF_CASE(512)
        DMST_SF_CASE(1024) DMST_SF_CASE(2048) DMST_SF_CASE(4096)
        default: return false;
    }
#undef DMST_SF_CASE
}
inline int sf_blocks_per_row(int n, int frames) { const int fpb = kSfPoints / (n / 2); return (frames + fpb - 1) / fpb; }

}  // namespace dmst
