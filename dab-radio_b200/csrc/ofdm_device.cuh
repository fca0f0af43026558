// Device building blocks of the OFDM demodulator kernels (sm_100a):
//   * register radix-16 / radix-R3 DFTs and the three-pass block FFT (N = 16 * 16 * R3, N/16 threads, 16 points each)
//     with two shared-memory exchanges laid out bank-conflict free for 8-byte accesses,
//   * the PLL rotation with the reference's float phase arithmetic (apply_pll.cpp:82-116) and MUFU sin/cos,
//   * warp/group reductions.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace dabb200 {

// Complex arithmetic on the packed FP32x2 pipe of sm_100 (FADD2 / FMUL2 / FFMA2): one instruction per complex add, two per
// complex multiply.  ptxas folds the operand shapes used below into the instruction's own operand modifiers -- scalar
// broadcast (R.F32), half swap (R.F32x2.LO_HI) and per-half negation (.NP / .PN) -- so multiplying by +-j or by a conjugate
// costs nothing (probe: tools/f32x2_probe.cu; the packed forms issue at half the rate of the scalar ones, so the FP pipe
// time is the same and the issue slots halve).  Rounding is identical to the scalar fmaf forms these replace.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return __ffma2_rn(a, make_float2(b.x, b.x), __fmul2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y)));
}
// a * conj(b)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
    return __ffma2_rn(a, make_float2(b.x, b.x), __fmul2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y)));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -j
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }

// forward DFT of 4 points (W = exp(-2 pi j / 4) = -j)
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
    const float2 s0 = cadd(a, c), d0 = csub(a, c), s1 = cadd(b, d), d1 = mul_mj(csub(b, d));
    a = cadd(s0, s1);
    b = cadd(d0, d1);
    c = csub(s0, s1);
    d = csub(d0, d1);
}

__device__ __forceinline__ void dft2(float2& a, float2& b) {
    const float2 s = cadd(a, b), d = csub(a, b);
    a = s;
    b = d;
}

constexpr float kC8 = 0.70710678118654752440f;   // cos(pi/4)
constexpr float kC16 = 0.92387953251128675613f;  // cos(pi/8)
constexpr float kS16 = 0.38268343236508977173f;  // sin(pi/8)

// multiply by W_8^1 = (1 - j)/sqrt(2) and W_8^3 = (-1 - j)/sqrt(2)
__device__ __forceinline__ float2 mul_w8_1(float2 a) { return __fmul2_rn(__fadd2_rn(a, make_float2(a.y, -a.x)), make_float2(kC8, kC8)); }
__device__ __forceinline__ float2 mul_w8_3(float2 a) { return __fmul2_rn(__fadd2_rn(a, make_float2(-a.y, a.x)), make_float2(-kC8, -kC8)); }

// forward DFT of 8 points, natural order in and out
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    // n = 4 n1 + n2 (n1 < 2, n2 < 4), k = k1 + 2 k2
    float2 a[4], b[4];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
        a[n2] = cadd(v[n2], v[n2 + 4]);  // k1 = 0
        b[n2] = csub(v[n2], v[n2 + 4]);  // k1 = 1, twiddle W_8^{n2}
    }
    b[1] = mul_w8_1(b[1]);
    b[2] = mul_mj(b[2]);
    b[3] = mul_w8_3(b[3]);
    dft4(a[0], a[1], a[2], a[3]);
    dft4(b[0], b[1], b[2], b[3]);
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) {
        v[2 * k2] = a[k2];
        v[2 * k2 + 1] = b[k2];
    }
}

// forward DFT of 16 points, natural order in and out: n = 4 n1 + n2, k = k1 + 4 k2
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);  // v[4 k1 + n2] = A[n2][k1]
    // twiddles W_16^{n2 k1}
    const float2 w1 = make_float2(kC16, -kS16), w2 = make_float2(kC8, -kC8), w3 = make_float2(kS16, -kC16);
    v[4 * 1 + 1] = cmul(v[4 * 1 + 1], w1);
    v[4 * 2 + 1] = mul_w8_1(v[4 * 2 + 1]);
    v[4 * 3 + 1] = cmul(v[4 * 3 + 1], w3);
    v[4 * 1 + 2] = mul_w8_1(v[4 * 1 + 2]);
    v[4 * 2 + 2] = mul_mj(v[4 * 2 + 2]);
    v[4 * 3 + 2] = mul_w8_3(v[4 * 3 + 2]);
    v[4 * 1 + 3] = cmul(v[4 * 1 + 3], w3);
    v[4 * 2 + 3] = mul_w8_3(v[4 * 2 + 3]);
    v[4 * 3 + 3] = cmul(v[4 * 3 + 3], make_float2(-kC16, kS16));  // W_16^9
    (void)w2;
    // DFT4 over n2 for each k1; result X[k1 + 4 k2] lands in v[4 k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // transpose the 4x4 register tile so that v[k] = X[k]
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
        for (int k2 = k1 + 1; k2 < 4; k2++) {
            const float2 tmp = v[4 * k1 + k2];
            v[4 * k1 + k2] = v[4 * k2 + k1];
            v[4 * k2 + k1] = tmp;
        }
}

// Geometry of the block FFT for N = 16 * 16 * R3
template <int NFFT>
struct FftGeom {
    static_assert(NFFT == 256 || NFFT == 512 || NFFT == 1024 || NFFT == 2048, "supported FFT sizes");
    static constexpr int R3 = NFFT / 256;       // radix of the last pass
    static constexpr int T = NFFT / 16;         // threads per transform
    static constexpr int PER3 = 16 / R3;        // pass-3 sub-transforms per thread (R3 >= 2)
    static constexpr int E1_STRIDE = T + R3;    // row stride (float2) of the first exchange: conflict free on both sides
    static constexpr int E1_SIZE = 16 * E1_STRIDE;
    static constexpr int E2_SIZE = NFFT;        // second exchange, rotated columns instead of padding
    static constexpr int TW1_SIZE = 16 * T;     // W_N^{t k1}, [k1][t]
    static constexpr int TW2_SIZE = 16 * R3;    // W_T^{n3 k2}, [k2][n3]
    // exchange 2 address of element (k1, k2, n3)
    __device__ static __forceinline__ int e2(int k1, int k2, int n3) { return (k2 * R3 + n3) * 16 + ((k1 + PER3 * n3) & 15); }
};

// Twiddle tables: computed once per handle into global memory (fft_twiddle_init_kernel), then copied into shared memory
// by every CTA that runs transforms.  Layout: tw1[k1][t] = W_N^{t k1} (TW1_SIZE), followed by tw2[k2][n3] = W_T^{n3 k2} (TW2_SIZE).
template <int NFFT>
__global__ void fft_twiddle_init_kernel(float2* table) {
    using G = FftGeom<NFFT>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G::TW1_SIZE) {
        const int k1 = i / G::T, t = i % G::T;
        const int e = (k1 * t) % NFFT;
        float s, c;
        sincospif(-2.0f * float(e) / float(NFFT), &s, &c);
        table[i] = make_float2(c, s);
    } else if (i < G::TW1_SIZE + G::TW2_SIZE) {
        const int j = i - G::TW1_SIZE;
        const int k2 = j / G::R3, n3 = j % G::R3;
        const int e = (k2 * n3) % G::T;
        float s, c;
        sincospif(-2.0f * float(e) / float(G::T), &s, &c);
        table[i] = make_float2(c, s);
    }
}

template <int NFFT>
__device__ __forceinline__ void fft_load_twiddles(float2* tw_smem, const float2* __restrict__ table, int tid, int nthreads) {
    using G = FftGeom<NFFT>;
    // 16-byte copies: both table sizes are even
    const float4* src = reinterpret_cast<const float4*>(table);
    float4* dst = reinterpret_cast<float4*>(tw_smem);
    for (int i = tid; i < (G::TW1_SIZE + G::TW2_SIZE) / 2; i += nthreads) dst[i] = __ldg(src + i);
}

// Three-pass forward FFT, split at its two barriers so callers can overlap other work with them.
//   fft_pass1: on entry v[n1] = x[n1 * T + t] for thread t of the transform; writes exchange 1
//   -- barrier --
//   fft_pass2: reads exchange 1, writes exchange 2 (for N = 256 the transform ends here: v[k2] = X[t + 16 k2])
//   -- barrier --
//   fft_pass3: v[m * R3 + k3] = X[k1 + 16 k2 + 256 k3] with p = t + T m, k1 = p & 15, k2 = p >> 4
template <int NFFT>
__device__ __forceinline__ void fft_pass1(float2 (&v)[16], int t, float2* e1, const float2* tw1) {
    using G = FftGeom<NFFT>;
    dft16(v);  // DFT16 over n1, then twiddle W_N^{t k1}, scatter A[k1][t]
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) {
        const float2 a = (k1 == 0) ? v[0] : cmul(v[k1], tw1[k1 * G::T + t]);
        e1[k1 * G::E1_STRIDE + t] = a;
    }
}

template <int NFFT>
__device__ __forceinline__ void fft_pass2(float2 (&v)[16], int t, const float2* e1, float2* e2, const float2* tw2) {
    using G = FftGeom<NFFT>;
    // thread u = (k1, n3): DFT16 over n2 of A[k1][n2 * R3 + n3], twiddle W_T^{n3 k2}
    const int k1p = t / G::R3, n3p = t % G::R3;
#pragma unroll
    for (int n2 = 0; n2 < 16; n2++) v[n2] = e1[k1p * G::E1_STRIDE + n2 * G::R3 + n3p];
    dft16(v);
    if (G::R3 > 1) {
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const float2 b = (k2 == 0) ? v[0] : cmul(v[k2], tw2[k2 * G::R3 + n3p]);
            e2[G::e2(k1p, k2, n3p)] = b;
        }
    }
}

// Exchange 1 without padding (ofdm_frame_v3.cuh): element (k1, col) lives at k1 * T + (col ^ e1_swizzle(k1)).  The pass-1 store of a
// row is a permutation of consecutive columns; the pass-2 load of 16 lanes covers 16 / R3 consecutive rows at the columns
// n2 R3 + n3: the row-dependent xor above the n3 bits sends them to 16 different 8-byte banks.
template <int NFFT>
__device__ __forceinline__ int e1_swizzle(int k1) {
    using G = FftGeom<NFFT>;
    return (k1 & (16 / G::R3 - 1)) * G::R3;
}

// fft_pass2 over the swizzled exchange 1, with the twiddle loads (global memory, L1-resident: 16 R3 values per CTA) issued one batch
// of four ahead of the exchange-2 stores (see ofdm_frame_v3.cuh, pass 1)
template <int NFFT>
__device__ __forceinline__ void fft_pass2_pipelined(float2 (&v)[16], int t, const float2* e1, float2* e2, const float2* __restrict__ tw2) {
    using G = FftGeom<NFFT>;
    const int k1p = t / G::R3, n3p = t % G::R3;
    const int sw = e1_swizzle<NFFT>(k1p);
#pragma unroll
    for (int n2 = 0; n2 < 16; n2++) v[n2] = e1[k1p * G::T + ((n2 * G::R3 + n3p) ^ sw)];
    if (G::R3 == 1) {
        dft16(v);
        return;
    }
    float2 wn[4];
#pragma unroll
    for (int q = 0; q < 4; q++) wn[q] = __ldg(tw2 + q * G::R3 + n3p);
    dft16(v);
#pragma unroll
    for (int b = 0; b < 4; b++) {
        float2 wc[4];
#pragma unroll
        for (int q = 0; q < 4; q++) wc[q] = wn[q];
        if (b < 3) {
#pragma unroll
            for (int q = 0; q < 4; q++) wn[q] = __ldg(tw2 + (4 * (b + 1) + q) * G::R3 + n3p);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k2 = 4 * b + q;
            e2[G::e2(k1p, k2, n3p)] = (k2 == 0) ? v[0] : cmul(v[k2], wc[q]);
        }
    }
}

template <int NFFT>
__device__ __forceinline__ void fft_pass3(float2 (&v)[16], int t, const float2* e2) {
    using G = FftGeom<NFFT>;
    if (G::R3 == 1) return;
#pragma unroll
    for (int m = 0; m < G::PER3; m++) {
        const int p = t + G::T * m, k1 = p & 15, k2 = p >> 4;
        if (G::R3 == 8) {
            float2 c[8];
#pragma unroll
            for (int n3 = 0; n3 < 8; n3++) c[n3] = e2[G::e2(k1, k2, n3)];
            dft8(c);
#pragma unroll
            for (int k3 = 0; k3 < 8; k3++) v[m * 8 + k3] = c[k3];
        } else if (G::R3 == 4) {
            float2 c0 = e2[G::e2(k1, k2, 0)], c1 = e2[G::e2(k1, k2, 1)], c2 = e2[G::e2(k1, k2, 2)], c3 = e2[G::e2(k1, k2, 3)];
            dft4(c0, c1, c2, c3);
            v[m * 4 + 0] = c0; v[m * 4 + 1] = c1; v[m * 4 + 2] = c2; v[m * 4 + 3] = c3;
        } else {
            float2 c0 = e2[G::e2(k1, k2, 0)], c1 = e2[G::e2(k1, k2, 1)];
            dft2(c0, c1);
            v[m * 2 + 0] = c0; v[m * 2 + 1] = c1;
        }
    }
}

// FFT bin held in register slot r of thread t after block_fft
template <int NFFT>
__device__ __forceinline__ int fft_out_bin(int t, int r) {
    using G = FftGeom<NFFT>;
    if (G::R3 == 1) return t + 16 * r;
    const int m = r / G::R3, k3 = r % G::R3;
    const int p = t + G::T * m;
    return (p & 15) + 16 * (p >> 4) + 256 * k3;
}

// round to nearest even for |x| < 2^22 without the conversion pipe
__device__ __forceinline__ float rint_magic(float x) {
    const float magic = 12582912.0f;  // 1.5 * 2^23
    return __fsub_rn(__fadd_rn(x, magic), magic);
}

// sin / cos of 2*pi*turns for turns in [-0.5, 0.5] on the MUFU pipe (abs error ~5e-7; the reference's polynomial
// chebyshev_sine.h:13-41 has 3.6e-8 -- both far below the 1/127 soft-bit step)
__device__ __forceinline__ float2 sincos_turns(float t_cos, float t_sin) {
    const float two_pi = 6.283185307179586f;
    return make_float2(__sinf(t_cos * two_pi), __sinf(t_sin * two_pi));
}

// Per-thread PLL state for one symbol: the reference evaluates, for sample i of a symbol (apply_pll.cpp:94-107, AVX):
//   base = fma(float(i & ~3), f, dt0);  t_sin = base + fl((i&3) f);  t_cos = base + fl(fl((i&3) f) + 0.25)
//   t -= roundeven(t);  y = x * (sin(2 pi t_cos) + j sin(2 pi t_sin))
// and, for the n % 4 tail samples (apply_pll.cpp:12-30), t_sin = fma(float(i - n_vec), f, fma(float(n_vec), f, dt0)), t_cos = t_sin + 0.25.
struct PllSymbol {
    float f, dt0, dt_tail;
    int n_vec;
};

__device__ __forceinline__ PllSymbol pll_symbol(float f, int sample_offset, int n_samples) {
    PllSymbol p;
    p.f = f;
    p.dt0 = float(sample_offset) * f;
    p.n_vec = n_samples & ~3;
    p.dt_tail = fmaf(float(p.n_vec), f, p.dt0);
    return p;
}

__device__ __forceinline__ float2 pll_rotate(const PllSymbol& p, float2 x, int i) {
    float ts, tc;
    if (i < p.n_vec) {
        const int k = i & 3;
        const float base = fmaf(float(i - k), p.f, p.dt0);
        const float pk = float(k) * p.f;  // exact for k = 0, 1, 2; fl(3 f) for k = 3, as the reference's packed constants
        ts = base + pk;
        tc = base + (pk + 0.25f);
    } else {
        ts = fmaf(float(i - p.n_vec), p.f, p.dt_tail);
        tc = ts + 0.25f;
    }
    ts -= rint_magic(ts);
    tc -= rint_magic(tc);
    const float2 cs = sincos_turns(tc, ts);  // (cos, sin)
    // c32_mul_avx (x86/c32_mul.h:10-40): re = fma(cos, x.re, -(sin x.im)), im = fma(cos, x.im, sin x.re)
    return make_float2(fmaf(cs.x, x.x, -(cs.y * x.y)), fmaf(cs.x, x.y, cs.y * x.x));
}

// sum over the `width` consecutive lanes of a group (width power of two <= 32)
template <int WIDTH>
__device__ __forceinline__ float2 group_reduce_sum(float2 v) {
#pragma unroll
    for (int d = WIDTH / 2; d >= 1; d >>= 1) {
        v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, d, WIDTH);
        v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, d, WIDTH);
    }
    return v;
}

}  // namespace dabb200
