// Parameter handling for the mix console, in float64 on the device:
//   * prepare: normalised (0..1) controller outputs -> denormalised values
//     (mst/modules.py:71-97) -> RBJ biquad coefficients, compressor constants, pan gains
//      -> the scan tables of common.cuh (RowTab);
//     out-of-range parameters are reported through `status` (ValueError of modules.py:86-89).
//   * grad epilogue: sums the per-tile gradient partials of the backward kernels and
//     chains them through the Jacobian of the parameter design, using forward-mode
//     dual numbers so the derivative code cannot drift from the design code.
#pragma once
#include "common.cuh"

namespace dmst {

constexpr double kPi = 3.14159265358979323846;

// ---- tiny forward-mode dual number: value + 3 partials (gain_db, cutoff, q) ----
struct Dual {
    double v, d[3];
};
__device__ __forceinline__ Dual dconst(double c) { return {c, {0, 0, 0}}; }
__device__ __forceinline__ Dual dvar(double c, int i) { Dual r = {c, {0, 0, 0}}; r.d[i] = 1; return r; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2]}}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]}}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, {-a.d[0], -a.d[1], -a.d[2]}}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) {
    return {a.v * b.v, {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2]}};
}
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
    double inv = 1.0 / b.v, q = a.v * inv;
    return {q, {(a.d[0] - q * b.d[0]) * inv, (a.d[1] - q * b.d[1]) * inv, (a.d[2] - q * b.d[2]) * inv}};
}
__device__ __forceinline__ Dual operator*(double s, Dual a) { return {s * a.v, {s * a.d[0], s * a.d[1], s * a.d[2]}}; }
__device__ __forceinline__ Dual operator+(double s, Dual a) { return {s + a.v, {a.d[0], a.d[1], a.d[2]}}; }
__device__ __forceinline__ Dual dsin(Dual a) { double c = cos(a.v); return {sin(a.v), {c * a.d[0], c * a.d[1], c * a.d[2]}}; }
__device__ __forceinline__ Dual dcos(Dual a) { double s = -sin(a.v); return {cos(a.v), {s * a.d[0], s * a.d[1], s * a.d[2]}}; }
__device__ __forceinline__ Dual dsqrt(Dual a) { double r = sqrt(a.v), k = 0.5 / r; return {r, {k * a.d[0], k * a.d[1], k * a.d[2]}}; }
__device__ __forceinline__ Dual dpow10(Dual a) { double r = pow(10.0, a.v), k = r * 2.302585092994046; return {r, {k * a.d[0], k * a.d[1], k * a.d[2]}}; }

// RBJ cookbook section, normalised by a0 (dasp_pytorch.signal.biquad, Appendix A).
// kind: 0 low_shelf, 1 peaking, 2 high_shelf.  out = {b0, b1, b2, a1, a2}.
__device__ inline void rbj_design(double gain_db, double freq, double q, double sr, int kind, Dual out[5]) {
    Dual G = dvar(gain_db, 0), F = dvar(freq, 1), Q = dvar(q, 2);
    Dual A = dpow10((1.0 / 40.0) * G);
    Dual w0 = (2.0 * kPi / sr) * F;
    Dual al = dsin(w0) / (2.0 * Q);
    Dual c = dcos(w0);
    Dual b0, b1, b2, a0, a1, a2;
    if (kind == 1) {
        b0 = 1.0 + al * A; b1 = -2.0 * c; b2 = 1.0 + (-(al * A));
        a0 = 1.0 + al / A; a1 = -2.0 * c; a2 = 1.0 + (-(al / A));
    } else {
        Dual sA2al = 2.0 * (dsqrt(A) * al);
        Dual Ap1 = 1.0 + A, Am1 = -1.0 + A;
        if (kind == 0) {
            b0 = A * (Ap1 - Am1 * c + sA2al);
            b1 = 2.0 * (A * (Am1 - Ap1 * c));
            b2 = A * (Ap1 - Am1 * c - sA2al);
            a0 = Ap1 + Am1 * c + sA2al;
            a1 = -2.0 * (Am1 + Ap1 * c);
            a2 = Ap1 + Am1 * c - sA2al;
        } else {
            b0 = A * (Ap1 + Am1 * c + sA2al);
            b1 = -2.0 * (A * (Am1 + Ap1 * c));
            b2 = A * (Ap1 + Am1 * c - sA2al);
            a0 = Ap1 - Am1 * c + sA2al;
            a1 = 2.0 * (Am1 - Ap1 * c);
            a2 = Ap1 - Am1 * c - sA2al;
        }
    }
    out[0] = b0 / a0; out[1] = b1 / a0; out[2] = b2 / a0; out[3] = a1 / a0; out[4] = a2 / a0;
}

__device__ __forceinline__ int section_kind(int k) { return k == 0 ? 0 : (k == 5 ? 2 : 1); }

struct M2 { double a, b, c, d; };
__device__ __forceinline__ M2 mmul(M2 x, M2 y) {
    return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
__device__ __forceinline__ void mstore(float* o, M2 m) { o[0] = (float)m.a; o[1] = (float)m.b; o[2] = (float)m.c; o[3] = (float)m.d; }

struct PrepareArgs {
    const float* params;   // tracks: (rows, np) ; master: (rows, 26)
    int rows;
    int np;                // parameters per row (27, 2 or 26)
    int kind;              // 0 advanced track, 1 basic track, 2 master, 3 neutral (no params)
    float lo[32], hi[32];
    double sr;
    int L[2];              // thread chunk lengths the tables are built for (powers of two)
    RowTab* tab[2];        // [rows] each; tab[1] may be null
    int* status;
    int status_base;       // added to the flat index reported in status[0]
    // chain workspace to clear before the chain kernels run (tickets, flags, mailboxes): 16-byte units, or null
    int4* zero[2];
    long long zero_n16[2];
};

// Cooperative clear of up to two workspace regions by a whole grid (replaces cudaMemsetAsync nodes in front of the
// chain kernels: in a replayed CUDA graph each memset node costs 3-4 us of latency on the critical path)
__device__ __forceinline__ void grid_zero(int4* const (&ptr)[2], const long long (&n16)[2]) {
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int4* p = r ? ptr[1] : ptr[0];
        const long long n = r ? n16[1] : n16[0];
        if (!p) continue;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = make_int4(0, 0, 0, 0);
    }
}

// index maps into the normalised parameter vector (mst/modules.py:353-460)
struct ParamMap { int gain_in, eq0, comp0, pan, gain_out; };
__device__ __forceinline__ ParamMap param_map(int kind) {
    if (kind == 0) return {0, 1, 19, 25, -1};
    if (kind == 1) return {0, -1, -1, 1, -1};
    if (kind == 3) return {-1, -1, -1, -1, -1};  // no parameters: neutral row
    return {25, 0, 18, -1, 24};
}

template <class Args>
__device__ __forceinline__ double denorm(const Args& a, const float* p, int i) {
    return (double)p[i] * ((double)a.hi[i] - (double)a.lo[i]) + (double)a.lo[i];
}

__device__ __forceinline__ M2 mpow_from_squares(const M2* sq, int e) {
    // product of sq[j] (= base^(2^j)) over the set bits of e
    M2 r = {1, 0, 0, 1};
    for (int j = 0; j < 9; ++j)
        if ((e >> j) & 1) r = mmul(r, sq[j]);
    return r;
}

// One warp per (row, job): jobs 0..5 = EQ sections, job 6 = gains / compressor / pan, job 7 =
// range check.  Lanes split the table entries (lane l builds P^l), everything in float64.
// grid: rows blocks of 256 threads.
__device__ __forceinline__ void prepare_row(const PrepareArgs& a, const int row) {
    const int job = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* p = a.params + (long long)row * a.np;
    const ParamMap pm = param_map(a.kind);

    if (job == 7) {
        int bad = 0x7fffffff;
        for (int i = lane; i < a.np; i += 32)
            if (p[i] < 0.0f || p[i] > 1.0f) bad = min(bad, a.status_base + 1 + i);
        if (bad != 0x7fffffff) atomicMin(a.status, bad);
        return;
    }
    if (job < 6) {
        double cf[5] = {1, 0, 0, 0, 0};
        if (pm.eq0 >= 0) {
            Dual o[5];
            rbj_design(denorm(a, p, pm.eq0 + 3 * job), denorm(a, p, pm.eq0 + 3 * job + 1),
                       denorm(a, p, pm.eq0 + 3 * job + 2), a.sr, section_kind(job), o);
            for (int j = 0; j < 5; ++j) cf[j] = o[j].v;
        }
        const float fb0 = (float)cf[0], fb1 = (float)cf[1], fb2 = (float)cf[2], fa1 = (float)cf[3], fa2 = (float)cf[4];
        // powers of the state matrix of the recursion the kernels actually run
        // (float32-rounded a1, a2), evaluated in float64
        const M2 A = {-(double)fa1, 1.0, -(double)fa2, 0.0};
        for (int v = 0; v < 2; ++v) {
            if (!a.tab[v]) continue;
            SectionTab& st = a.tab[v][row].sec[job];
            M2 P = A;
            for (int s = 1; s < a.L[v]; s <<= 1) P = mmul(P, P);  // A^L
            M2 sq[9];
            sq[0] = P;
            for (int j = 1; j < 9; ++j) sq[j] = mmul(sq[j - 1], sq[j - 1]);
            mstore(st.Ppow[lane], mpow_from_squares(sq, lane));
            if (lane < 9) mstore(st.P2[lane], sq[lane]);
            if (lane == 9) {
                st.b0 = fb0; st.b1 = fb1; st.b2 = fb2; st.a1 = fa1; st.a2 = fa2;
                st.inv_b0 = (float)(1.0 / (double)fb0);
                st.pad0 = st.pad1 = 0.0f;
                st.pad2[0] = st.pad2[1] = st.pad2[2] = st.pad2[3] = 0.0f;
            }
        }
        return;
    }
    // job 6: gains, compressor, pan
    double thr = 0, ratio = 1, attack = 1, knee = 1, makeup = 0;
    if (pm.comp0 >= 0) {
        thr = denorm(a, p, pm.comp0 + 0); ratio = denorm(a, p, pm.comp0 + 1);
        attack = denorm(a, p, pm.comp0 + 2); knee = denorm(a, p, pm.comp0 + 4);
        makeup = denorm(a, p, pm.comp0 + 5);  // comp0 + 3 = release_ms: unused upstream
    }
    const float alpha = (float)exp(-log(9.0) / (a.sr * (attack / 1e3)));
    for (int v = 0; v < 2; ++v) {
        if (!a.tab[v]) continue;
        RowTab& tb = a.tab[v][row];
        const double aL = pow((double)alpha, (double)a.L[v]);
        tb.a_lane[lane] = (float)pow(aL, (double)lane);
        tb.a_i[lane] = (float)pow((double)alpha, (double)(lane + 1));
        if (lane < 9) tb.a2pow[lane] = (float)pow(aL, (double)(1 << lane));
        if (lane == 9) {
            tb.g_in = (pm.gain_in >= 0) ? (float)pow(10.0, denorm(a, p, pm.gain_in) / 20.0) : 1.0f;
            tb.g_out = (pm.gain_out >= 0) ? (float)pow(10.0, denorm(a, p, pm.gain_out) / 20.0) : 1.0f;
            if (pm.pan >= 0) {
                const double th = denorm(a, p, pm.pan) * (kPi / 2);
                tb.gL = (float)sqrt(((kPi / 2) - th) * (2 / kPi) * cos(th));
                tb.gR = (float)sqrt(th * (2 / kPi) * sin(th));
            } else {
                tb.gL = tb.gR = 1.0f;
            }
            tb.alpha = alpha;
            tb.beta = (float)(1.0 - (double)alpha);
            tb.thr_lo = (float)(thr - knee / 2);
            tb.knee = (float)knee; tb.inv_knee = (float)(1.0 / knee); tb.inv_2knee = (float)(0.5 / knee);
            tb.slope = (float)(1.0 / ratio - 1.0);
            tb.makeup = (float)makeup;
            tb.inv_ratio2 = (float)(1.0 / (ratio * ratio));
            tb.pad[0] = tb.pad[1] = tb.pad[2] = 0.0f; tb.pad2[0] = tb.pad2[1] = tb.pad2[2] = 0.0f;
        }
    }
}
__global__ void prepare_kernel(PrepareArgs a) {
    grid_zero(a.zero, a.zero_n16);
    prepare_row(a, blockIdx.x);
}
// track rows and master-bus rows in one launch (blocks [0, a.rows) serve a, the rest b; a's clear is the grid's)
__global__ void prepare2_kernel(PrepareArgs a, PrepareArgs b) {
    grid_zero(a.zero, a.zero_n16);
    if ((int)blockIdx.x < a.rows) prepare_row(a, blockIdx.x);
    else prepare_row(b, blockIdx.x - a.rows);
}

// Tables of the track backward kernel (console_bwd2.cuh): per (row, section) the two delta-form recursions
// g = A^-T u and h = (B/b0)^-T u (common.cuh, RecTab), designed in float64 from the denormalised parameters.
struct PrepareBwdArgs {
    const float* params;   // (rows, np) normalised track parameters
    int rows, np;
    float lo[32], hi[32];
    double sr;
    EqBwdTab* tab;         // [rows]
    int4* zero[2];         // backward chain workspace to clear (see grid_zero), or null
    long long zero_n16[2];
};
// grid: rows blocks of kNumRec warps; warp r builds recursion r = 2 * section + type, lane l builds P^l
__global__ void prepare_bwd_kernel(PrepareBwdArgs a) {
    grid_zero(a.zero, a.zero_n16);
    const int row = blockIdx.x;
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (r >= kNumRec) return;
    const int sec = r >> 1, type = r & 1;
    const float* p = a.params + (long long)row * a.np;
    const int i0 = 1 + 3 * sec;   // advanced track layout (mst/modules.py:353-380)
    Dual o[5];
    rbj_design(denorm(a, p, i0), denorm(a, p, i0 + 1), denorm(a, p, i0 + 2), a.sr, section_kind(sec), o);
    double c1, c2, scale;
    if (type == 0) { c1 = o[3].v; c2 = o[4].v; scale = 1.0; }
    else { c1 = o[1].v / o[0].v; c2 = o[2].v / o[0].v; scale = 1.0 / o[0].v; }
    const float nc0 = (float)(-(1.0 + c1 + c2)), nd2 = (float)(-(1.0 - c2)), k2 = (float)c2;
    // one-sample transition of (s, v) of the recursion the kernel actually runs (its float32 constants), in float64
    const double c0e = -(double)nc0, a2e = (double)k2;
    M2 P = {1.0 - c0e, a2e, -c0e, a2e};
    for (int s = 1; s < kBwd2Chunk; s <<= 1) P = mmul(P, P);   // M^32: one chunk
    M2 sq[6];
    sq[0] = P;
    for (int j = 1; j < 6; ++j) sq[j] = mmul(sq[j - 1], sq[j - 1]);
    M2 pw = {1, 0, 0, 1};
    for (int j = 0; j < 5; ++j)
        if ((lane >> j) & 1) pw = mmul(pw, sq[j]);
    PairTab& pt = a.tab[row].sec[sec];
    auto put = [&](float2* m4, M2 m) {   // this recursion's half of a pair of row-major 2x2 matrices
        float* f = reinterpret_cast<float*>(m4) + type;
        f[0] = (float)m.a; f[2] = (float)m.b; f[4] = (float)m.c; f[6] = (float)m.d;
    };
    put(pt.Ppow[lane], pw);
    if (lane < 6) put(pt.P2[lane], sq[lane]);
    if (lane == 6) {
        (&pt.nc0.x)[type] = nc0; (&pt.nd2.x)[type] = nd2; (&pt.scale.x)[type] = (float)scale; (&pt.k2.x)[type] = k2;
    }
}

struct EpilogueArgs {
    const float* params;
    int rows, np, kind;
    float lo[32], hi[32];
    double sr;
    const float* partial;  // [rows][ntiles][kGradCount]
    int ntiles;
    unsigned flags;        // kChain* actually enabled (disabled stages contribute zero)
    float* grad;           // [rows][np], gradient w.r.t. the NORMALISED parameters
};

// one 64-thread block per row: threads 0..kGradCount-1 reduce the tile partials (coalesced,
// fixed order => deterministic), then threads 0..5 chain one EQ section each and thread 6
// the gains / compressor / pan through the design Jacobian.
__device__ __forceinline__ void grad_epilogue_row(const EpilogueArgs& a, const int row) {
    const int tid = threadIdx.x;
    DMST_SHARED_ARRAY(double, acc, kGradCount);
    if (tid < kGradCount) {
        double s = 0.0;
        const float* q = a.partial + (long long)row * a.ntiles * kGradCount + tid;
        for (int t = 0; t < a.ntiles; ++t) s += (double)q[(long long)t * kGradCount];
        acc[tid] = s;
    }
    __syncthreads();
    const float* p = a.params + (long long)row * a.np;
    float* g = a.grad + (long long)row * a.np;
    const ParamMap pm = param_map(a.kind);
    auto scale = [&](int i) { return (double)a.hi[i] - (double)a.lo[i]; };
    const double ln10_20 = 0.11512925464970229;
    if (tid < kNumSections) {
        const int k = tid;
        if (pm.eq0 >= 0) {
            const int i0 = pm.eq0 + 3 * k;
            if (a.flags & kChainEq) {
                Dual o[5];
                rbj_design(denorm(a, p, i0), denorm(a, p, i0 + 1), denorm(a, p, i0 + 2), a.sr, section_kind(k), o);
                // the backward kernel accumulates in the basis {b0+b1+b2, b1+2 b2, b2, a1+a2, a2}
                Dual basis[5] = {o[0] + o[1] + o[2], o[1] + 2.0 * o[2], o[2], o[3] + o[4], o[4]};
                for (int d = 0; d < 3; ++d) {
                    double s = 0.0;
                    for (int j = 0; j < 5; ++j) s += acc[kGradEq + 5 * k + j] * basis[j].d[d];
                    g[i0 + d] = (float)(s * scale(i0 + d));
                }
            } else {
                g[i0] = g[i0 + 1] = g[i0 + 2] = 0.0f;
            }
        }
    } else if (tid == 6) {
        if (pm.gain_in >= 0)
            g[pm.gain_in] = (a.flags & kChainGain) ? (float)(acc[kGradGin] * ln10_20 * scale(pm.gain_in)) : 0.0f;
        if (pm.gain_out >= 0)
            g[pm.gain_out] = (a.flags & kChainOutGain) ? (float)(acc[kGradGout] * ln10_20 * scale(pm.gain_out)) : 0.0f;
        if (pm.comp0 >= 0) {
            if (a.flags & kChainComp) {
                const double attack = denorm(a, p, pm.comp0 + 2);
                const double alpha = exp(-log(9.0) / (a.sr * (attack / 1e3)));
                const double dalpha_dattack = alpha * log(9.0) * 1e3 / (a.sr * attack * attack);
                g[pm.comp0 + 0] = (float)(acc[kGradThr] * scale(pm.comp0 + 0));
                g[pm.comp0 + 1] = (float)(acc[kGradRatio] * scale(pm.comp0 + 1));
                g[pm.comp0 + 2] = (float)(acc[kGradAlpha] * dalpha_dattack * scale(pm.comp0 + 2));
                g[pm.comp0 + 3] = 0.0f;  // release_ms has no effect upstream
                g[pm.comp0 + 4] = (float)(acc[kGradKnee] * scale(pm.comp0 + 4));
                g[pm.comp0 + 5] = (float)(acc[kGradMakeup] * scale(pm.comp0 + 5));
            } else {
                for (int j = 0; j < 6; ++j) g[pm.comp0 + j] = 0.0f;
            }
        }
        if (pm.pan >= 0) {
            const double th = denorm(a, p, pm.pan) * (kPi / 2);
            const double uL = ((kPi / 2) - th) * (2 / kPi) * cos(th), uR = th * (2 / kPi) * sin(th);
            const double duL = (2 / kPi) * (-cos(th) - ((kPi / 2) - th) * sin(th));
            const double duR = (2 / kPi) * (sin(th) + th * cos(th));
            const double dgL = uL > 0 ? 0.5 * duL / sqrt(uL) : 0.0, dgR = uR > 0 ? 0.5 * duR / sqrt(uR) : 0.0;
            g[pm.pan] = (float)((acc[kGradGL] * dgL + acc[kGradGR] * dgR) * (kPi / 2) * scale(pm.pan));
        }
        if (a.kind == 0) g[26] = 0.0f;  // fx send: the fx bus is off, no path to the mix
    }
}
__global__ void grad_epilogue_kernel(EpilogueArgs a) { grad_epilogue_row(a, blockIdx.x); }
// master-bus rows and track rows in one launch (blocks [0, a.rows) serve a, the rest b)
__global__ void grad_epilogue2_kernel(EpilogueArgs a, EpilogueArgs b) {
    if ((int)blockIdx.x < a.rows) grad_epilogue_row(a, blockIdx.x);
    else grad_epilogue_row(b, blockIdx.x - a.rows);
}

}  // namespace dmst
