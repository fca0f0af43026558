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
//   -DDAB_SHIM_FAST    single-precision Stockham radix-2 with per-stage twiddle tables and thread-local work buffers, written
//                      so that gcc -O3 vectorises it (AVX2).  Used ONLY for CPU-baseline timing (bench.py), so that the
//                      reference is not handicapped by a slow stand-in FFT; ~2-3x slower than real FFTW codelets, stated in
//                      every report.
#pragma once
#include <cmath>
#include <complex>
#include <cstddef>
#include <vector>

typedef float fftwf_complex[2];

#ifdef DAB_SHIM_FAST
struct fftwf_plan_s {
    int n, log2n, sign;
    std::vector<float> twr, twi;  // [stage][j], j < n/2: exp(sign*2*pi*i*(j % Ns)/(2*Ns)), Ns = 1 << stage
};
typedef fftwf_plan_s* fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex*, fftwf_complex*, int sign, unsigned) {
    auto* p = new fftwf_plan_s();
    p->n = n;
    p->sign = sign;
    p->log2n = 0;
    while ((1 << p->log2n) < n) p->log2n++;
    const int half = n / 2;
    p->twr.resize(size_t(p->log2n) * half);
    p->twi.resize(size_t(p->log2n) * half);
    const double two_pi = 6.283185307179586476925286766559;
    for (int s = 0; s < p->log2n; s++) {
        const int Ns = 1 << s;
        for (int j = 0; j < half; j++) {
            const double a = double(sign) * two_pi * double(j % Ns) / double(2 * Ns);
            p->twr[size_t(s) * half + j] = float(std::cos(a));
            p->twi[size_t(s) * half + j] = float(std::sin(a));
        }
    }
    return p;
}

static inline void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out) {
    const int n = p->n, half = n / 2;
    static thread_local std::vector<float> work;
    if (work.size() < size_t(4 * n)) work.resize(size_t(4 * n));
    float* ar = work.data();
    float* ai = ar + n;
    float* br = ai + n;
    float* bi = br + n;
    for (int i = 0; i < n; i++) { ar[i] = in[i][0]; ai[i] = in[i][1]; }
    for (int s = 0; s < p->log2n; s++) {
        const int Ns = 1 << s;
        const float* __restrict__ wr = &p->twr[size_t(s) * half];
        const float* __restrict__ wi = &p->twi[size_t(s) * half];
        const float* __restrict__ xr = ar; const float* __restrict__ xi = ai;
        float* __restrict__ yr = br; float* __restrict__ yi = bi;
        if (Ns >= 8) {
            for (int j0 = 0; j0 < half; j0 += Ns) {
                const int o = 2 * j0;
                for (int k = 0; k < Ns; k++) {
                    const int j = j0 + k;
                    const float tr = xr[j + half] * wr[j] - xi[j + half] * wi[j];
                    const float ti = xr[j + half] * wi[j] + xi[j + half] * wr[j];
                    yr[o + k] = xr[j] + tr; yi[o + k] = xi[j] + ti;
                    yr[o + k + Ns] = xr[j] - tr; yi[o + k + Ns] = xi[j] - ti;
                }
            }
        } else {
            for (int j = 0; j < half; j++) {
                const int k = j & (Ns - 1);
                const int o = (j - k) * 2 + k;
                const float tr = xr[j + half] * wr[j] - xi[j + half] * wi[j];
                const float ti = xr[j + half] * wi[j] + xi[j + half] * wr[j];
                yr[o] = xr[j] + tr; yi[o] = xi[j] + ti;
                yr[o + Ns] = xr[j] - tr; yi[o + Ns] = xi[j] - ti;
            }
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
