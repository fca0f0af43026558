// TEST INFRASTRUCTURE ONLY (oracle/). Not shipped, not on the product path.
//
// Minimal stand-in for <fftw3.h> so that the reference's own sources
// (/root/reference/src/ofdm/ofdm_demodulator.cpp:11,113-114,893,898 and
//  /root/reference/src/ofdm/ofdm_modulator.cpp:18,46,162-165) compile in this image,
// where FFTW3 (vcpkg.json:19-21, fftw3 >= 3.3.10) is not installed.
//
// Semantics reproduced (the only ones the reference uses):
//   - fftwf_plan_dft_1d(n, nullptr, nullptr, sign, FFTW_ESTIMATE): power-of-two c2c plan
//   - fftwf_execute_dft(plan, in, out): unnormalised DFT, FORWARD = e^{-j..}, BACKWARD = e^{+j..},
//     in-place (in == out) allowed
//   - fftwf_destroy_plan
// The transform is evaluated in double precision and rounded once to float, i.e. it is a
// *tighter* DFT than FFTW's single precision codelets.  FFT rounding at this boundary is
// "parity unpinned" (no reference test pins it); north_star's +-1 LSB tolerance absorbs it.
//
// Two variants, chosen at compile time:
//   default            double-precision radix-2, rounded once: the accuracy reference used by every parity test
//   -DDAB_SHIM_FAST    single-precision Stockham radix-4 (AVX2 + FMA intrinsics, needs -march=x86-64-v3) with per-stage twiddle
//                      tables and thread-local work buffers.  Used ONLY for CPU-baseline timing (bench.py), so that the
//                      reference is not handicapped by a slow stand-in FFT; its time per transform is reported in every
//                      bench line (cpu_baseline.fft) so that the reader can bound what real FFTW codelets would change.
#pragma once
#include <cmath>
#include <complex>
#include <cstddef>
#include <vector>

typedef float fftwf_complex[2];

#ifdef DAB_SHIM_FAST
#include <immintrin.h>
struct fftwf_plan_s {
    int n, sign;
    int n4;                        // radix-4 stages; one radix-2 stage follows when log2(n) is odd
    bool tail2;
    std::vector<float> tw;         // per radix-4 stage: w1r, w1i, w2r, w2i, w3r, w3i, each [n / 4], indexed by j
    std::vector<float> tw2r, tw2i; // radix-2 tail: [n / 2]
};
typedef fftwf_plan_s* fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex*, fftwf_complex*, int sign, unsigned) {
    auto* p = new fftwf_plan_s();
    p->n = n;
    p->sign = sign;
    int log2n = 0;
    while ((1 << log2n) < n) log2n++;
    p->n4 = log2n / 2;
    p->tail2 = (log2n & 1) != 0;
    const int q = n / 4;
    const double two_pi = 6.283185307179586476925286766559;
    p->tw.resize(size_t(p->n4) * 6 * size_t(q));
    int Ns = 1;
    for (int s = 0; s < p->n4; s++, Ns *= 4) {
        float* t = &p->tw[size_t(s) * 6 * size_t(q)];
        for (int j = 0; j < q; j++) {
            const double a = double(sign) * two_pi * double(j % Ns) / double(4 * Ns);
            for (int m = 1; m <= 3; m++) {
                t[size_t(2 * (m - 1)) * q + j] = float(std::cos(double(m) * a));
                t[size_t(2 * (m - 1) + 1) * q + j] = float(std::sin(double(m) * a));
            }
        }
    }
    if (p->tail2) {
        const int half = n / 2;
        p->tw2r.resize(size_t(half));
        p->tw2i.resize(size_t(half));
        for (int j = 0; j < half; j++) {   // Ns = n / 2: j % Ns = j
            const double a = double(sign) * two_pi * double(j) / double(n);
            p->tw2r[size_t(j)] = float(std::cos(a));
            p->tw2i[size_t(j)] = float(std::sin(a));
        }
    }
    return p;
}

// Stockham autosort, radix 4 (+ one radix-2 stage for odd log2 n), split real / imaginary work arrays, AVX2 + FMA intrinsics
// (8 butterflies per step; the first two stages transpose their outputs in registers).  n >= 32.
struct dab_shim_bfly4 { __m256 y0r, y0i, y1r, y1i, y2r, y2i, y3r, y3i; };
static inline dab_shim_bfly4 dab_shim_radix4(const float* xr, const float* xi, int q, const float* w1r, const float* w1i, const float* w2r,
                                             const float* w2i, const float* w3r, const float* w3i, bool twiddle, __m256 sg) {
    const __m256 x0r = _mm256_loadu_ps(xr), x0i = _mm256_loadu_ps(xi);
    __m256 x1r = _mm256_loadu_ps(xr + q), x1i = _mm256_loadu_ps(xi + q);
    __m256 x2r = _mm256_loadu_ps(xr + 2 * q), x2i = _mm256_loadu_ps(xi + 2 * q);
    __m256 x3r = _mm256_loadu_ps(xr + 3 * q), x3i = _mm256_loadu_ps(xi + 3 * q);
    if (twiddle) {
        __m256 wr = _mm256_loadu_ps(w1r), wi = _mm256_loadu_ps(w1i), tr = x1r;
        x1r = _mm256_fmsub_ps(tr, wr, _mm256_mul_ps(x1i, wi));
        x1i = _mm256_fmadd_ps(tr, wi, _mm256_mul_ps(x1i, wr));
        wr = _mm256_loadu_ps(w2r); wi = _mm256_loadu_ps(w2i); tr = x2r;
        x2r = _mm256_fmsub_ps(tr, wr, _mm256_mul_ps(x2i, wi));
        x2i = _mm256_fmadd_ps(tr, wi, _mm256_mul_ps(x2i, wr));
        wr = _mm256_loadu_ps(w3r); wi = _mm256_loadu_ps(w3i); tr = x3r;
        x3r = _mm256_fmsub_ps(tr, wr, _mm256_mul_ps(x3i, wi));
        x3i = _mm256_fmadd_ps(tr, wi, _mm256_mul_ps(x3i, wr));
    }
    const __m256 s02r = _mm256_add_ps(x0r, x2r), s02i = _mm256_add_ps(x0i, x2i), d02r = _mm256_sub_ps(x0r, x2r), d02i = _mm256_sub_ps(x0i, x2i);
    const __m256 s13r = _mm256_add_ps(x1r, x3r), s13i = _mm256_add_ps(x1i, x3i);
    const __m256 e13i = _mm256_mul_ps(sg, _mm256_sub_ps(x1r, x3r));          // sign * i * (x1 - x3)
    const __m256 e13r = _mm256_mul_ps(sg, _mm256_sub_ps(x3i, x1i));
    dab_shim_bfly4 b;
    b.y0r = _mm256_add_ps(s02r, s13r); b.y0i = _mm256_add_ps(s02i, s13i);
    b.y1r = _mm256_add_ps(d02r, e13r); b.y1i = _mm256_add_ps(d02i, e13i);
    b.y2r = _mm256_sub_ps(s02r, s13r); b.y2i = _mm256_sub_ps(s02i, s13i);
    b.y3r = _mm256_sub_ps(d02r, e13r); b.y3i = _mm256_sub_ps(d02i, e13i);
    return b;
}
// y[4 j + m] = ym[j] for the 8 butterflies of one step
static inline void dab_shim_store_interleaved4(float* y, __m256 y0, __m256 y1, __m256 y2, __m256 y3) {
    const __m256 t0 = _mm256_unpacklo_ps(y0, y1), t1 = _mm256_unpackhi_ps(y0, y1), t2 = _mm256_unpacklo_ps(y2, y3), t3 = _mm256_unpackhi_ps(y2, y3);
    const __m256 u0 = _mm256_shuffle_ps(t0, t2, 0x44), u1 = _mm256_shuffle_ps(t0, t2, 0xEE), u2 = _mm256_shuffle_ps(t1, t3, 0x44), u3 = _mm256_shuffle_ps(t1, t3, 0xEE);
    _mm256_storeu_ps(y, _mm256_permute2f128_ps(u0, u1, 0x20));
    _mm256_storeu_ps(y + 8, _mm256_permute2f128_ps(u2, u3, 0x20));
    _mm256_storeu_ps(y + 16, _mm256_permute2f128_ps(u0, u1, 0x31));
    _mm256_storeu_ps(y + 24, _mm256_permute2f128_ps(u2, u3, 0x31));
}

static inline void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out) {
    const int n = p->n, q = n / 4;
    static thread_local std::vector<float> work;
    if (work.size() < size_t(4 * n)) work.resize(size_t(4 * n));
    float* ar = work.data();
    float* ai = ar + n;
    float* br = ai + n;
    float* bi = br + n;
    for (int i = 0; i < n; i++) { ar[i] = in[i][0]; ai[i] = in[i][1]; }
    const __m256 sg = _mm256_set1_ps((p->sign < 0) ? -1.0f : 1.0f);
    int Ns = 1;
    for (int s = 0; s < p->n4; s++, Ns *= 4) {
        const float* t = &p->tw[size_t(s) * 6 * size_t(q)];
        const float *w1r = t, *w1i = t + q, *w2r = t + 2 * q, *w2i = t + 3 * q, *w3r = t + 4 * q, *w3i = t + 5 * q;
        if (Ns == 1) {          // no twiddles; butterfly j writes y[4j .. 4j + 3]
            for (int j = 0; j < q; j += 8) {
                const dab_shim_bfly4 b = dab_shim_radix4(ar + j, ai + j, q, w1r, w1i, w2r, w2i, w3r, w3i, false, sg);
                dab_shim_store_interleaved4(br + 4 * j, b.y0r, b.y1r, b.y2r, b.y3r);
                dab_shim_store_interleaved4(bi + 4 * j, b.y0i, b.y1i, b.y2i, b.y3i);
            }
        } else if (Ns == 4) {   // butterfly j = 4a + k writes y[16a + 4m + k]: two groups a per step
            for (int j = 0; j < q; j += 8) {
                const dab_shim_bfly4 b = dab_shim_radix4(ar + j, ai + j, q, w1r + j, w1i + j, w2r + j, w2i + j, w3r + j, w3i + j, true, sg);
                float* yr = br + 4 * j;
                float* yi = bi + 4 * j;
                const __m256 vr[4] = {b.y0r, b.y1r, b.y2r, b.y3r}, vi[4] = {b.y0i, b.y1i, b.y2i, b.y3i};
                for (int m = 0; m < 4; m++) {
                    _mm_storeu_ps(yr + 4 * m, _mm256_castps256_ps128(vr[m]));
                    _mm_storeu_ps(yr + 16 + 4 * m, _mm256_extractf128_ps(vr[m], 1));
                    _mm_storeu_ps(yi + 4 * m, _mm256_castps256_ps128(vi[m]));
                    _mm_storeu_ps(yi + 16 + 4 * m, _mm256_extractf128_ps(vi[m], 1));
                }
            }
        } else {                // runs of Ns >= 16 consecutive butterflies, outputs y[4 j0 + k + m Ns]
            for (int j0 = 0; j0 < q; j0 += Ns)
                for (int k = 0; k < Ns; k += 8) {
                    const int j = j0 + k;
                    const dab_shim_bfly4 b = dab_shim_radix4(ar + j, ai + j, q, w1r + j, w1i + j, w2r + j, w2i + j, w3r + j, w3i + j, true, sg);
                    float* yr = br + 4 * j0 + k;
                    float* yi = bi + 4 * j0 + k;
                    _mm256_storeu_ps(yr, b.y0r); _mm256_storeu_ps(yi, b.y0i);
                    _mm256_storeu_ps(yr + Ns, b.y1r); _mm256_storeu_ps(yi + Ns, b.y1i);
                    _mm256_storeu_ps(yr + 2 * Ns, b.y2r); _mm256_storeu_ps(yi + 2 * Ns, b.y2i);
                    _mm256_storeu_ps(yr + 3 * Ns, b.y3r); _mm256_storeu_ps(yi + 3 * Ns, b.y3i);
                }
        }
        std::swap(ar, br);
        std::swap(ai, bi);
    }
    if (p->tail2) {
        const int half = n / 2;
        const float* wr = p->tw2r.data();
        const float* wi = p->tw2i.data();
        for (int j = 0; j < half; j += 8) {   // Ns = half: butterfly j writes y[j], y[j + half]
            const __m256 xr0 = _mm256_loadu_ps(ar + j), xi0 = _mm256_loadu_ps(ai + j), xr1 = _mm256_loadu_ps(ar + j + half), xi1 = _mm256_loadu_ps(ai + j + half);
            const __m256 cr = _mm256_loadu_ps(wr + j), ci = _mm256_loadu_ps(wi + j);
            const __m256 tr = _mm256_fmsub_ps(xr1, cr, _mm256_mul_ps(xi1, ci)), ti = _mm256_fmadd_ps(xr1, ci, _mm256_mul_ps(xi1, cr));
            _mm256_storeu_ps(br + j, _mm256_add_ps(xr0, tr)); _mm256_storeu_ps(bi + j, _mm256_add_ps(xi0, ti));
            _mm256_storeu_ps(br + j + half, _mm256_sub_ps(xr0, tr)); _mm256_storeu_ps(bi + j + half, _mm256_sub_ps(xi0, ti));
        }
        std::swap(ar, br);
        std::swap(ai, bi);
    }
    for (int i = 0; i < n; i++) { out[i][0] = ar[i]; out[i][1] = ai[i]; }
}

static inline void fftwf_destroy_plan(fftwf_plan p) { delete p; }
#else

struct fftwf_plan_s {
    int n;
    int sign;
    std::vector<std::complex<double>> twiddle;  // e^{sign*2*pi*j*k/n}, k < n/2
    std::vector<int> bitrev;
};
typedef fftwf_plan_s* fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex*, fftwf_complex*, int sign, unsigned) {
    auto* p = new fftwf_plan_s();
    p->n = n;
    p->sign = sign;
    p->twiddle.resize(size_t(n) / 2);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < n / 2; k++) {
        const double a = double(sign) * two_pi * double(k) / double(n);
        p->twiddle[size_t(k)] = std::complex<double>(std::cos(a), std::sin(a));
    }
    int log2n = 0;
    while ((1 << log2n) < n) log2n++;
    p->bitrev.resize(size_t(n));
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < log2n; b++) r |= ((i >> b) & 1) << (log2n - 1 - b);
        p->bitrev[size_t(i)] = r;
    }
    return p;
}

static inline void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out) {
    const int n = p->n;
    std::vector<std::complex<double>> x(static_cast<size_t>(n));
    for (int i = 0; i < n; i++) {
        x[size_t(p->bitrev[size_t(i)])] = std::complex<double>(double(in[i][0]), double(in[i][1]));
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len / 2;
        const int step = n / len;
        for (int base = 0; base < n; base += len) {
            for (int k = 0; k < half; k++) {
                const std::complex<double> w = p->twiddle[size_t(k * step)];
                const std::complex<double> a = x[size_t(base + k)];
                const std::complex<double> b = x[size_t(base + k + half)];
                // explicit arithmetic so -ffast-math cannot reassociate across butterflies
                const double tr = b.real() * w.real() - b.imag() * w.imag();
                const double ti = b.real() * w.imag() + b.imag() * w.real();
                x[size_t(base + k)] = std::complex<double>(a.real() + tr, a.imag() + ti);
                x[size_t(base + k + half)] = std::complex<double>(a.real() - tr, a.imag() - ti);
            }
        }
    }
    for (int i = 0; i < n; i++) {
        out[i][0] = float(x[size_t(i)].real());
        out[i][1] = float(x[size_t(i)].imag());
    }
}

static inline void fftwf_destroy_plan(fftwf_plan p) { delete p; }
#endif  // DAB_SHIM_FAST
