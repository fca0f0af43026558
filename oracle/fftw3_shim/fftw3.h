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
#pragma once
#include <cmath>
#include <complex>
#include <cstddef>
#include <vector>

typedef float fftwf_complex[2];

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
