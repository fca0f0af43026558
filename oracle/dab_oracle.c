/* TEST INFRASTRUCTURE ONLY (oracle/) -- see dab_oracle.h.  Plain-C restatement of the reference hot path.
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Built by oracle/Makefile with -ffp-contract=off: fused multiply-adds appear only where written as fmaf(),
 * which is where the reference's AVX/FMA build (-march=native -ffast-math, CMakePresets.json:77-78) has them.
 */
#define _GNU_SOURCE
#include "dab_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------ tables */

/* src/ofdm/dab_ofdm_params_ref.cpp:10-57 */
int orc_get_params(int mode, orc_params* p) {
    switch (mode) {
    case 1: p->nb_frame_symbols = 76;  p->nb_symbol_period = 2552; p->nb_null_period = 2656; p->nb_fft = 2048; p->nb_data_carriers = 1536; break;
    case 2: p->nb_frame_symbols = 76;  p->nb_symbol_period = 638;  p->nb_null_period = 664;  p->nb_fft = 512;  p->nb_data_carriers = 384;  break;
    case 3: p->nb_frame_symbols = 153; p->nb_symbol_period = 319;  p->nb_null_period = 345;  p->nb_fft = 256;  p->nb_data_carriers = 192;  break;
    case 4: p->nb_frame_symbols = 76;  p->nb_symbol_period = 1276; p->nb_null_period = 1328; p->nb_fft = 1024; p->nb_data_carriers = 768;  break;
    default: return -1;
    }
    p->nb_cyclic_prefix = p->nb_symbol_period - p->nb_fft;
    return 0;
}

/* ETSI EN 300 401 clause 14.3.2, tables 23/24 as used by src/ofdm/dab_prs_ref.cpp:24-138.
 * One (i, n) pair per block of 32 carriers, blocks ordered k = -K/2 .. -1 then 1 .. K/2. */
static const char* const PRS_I[4] = {
    "012301230123012301230123" "032103210321032103210321",
    "012301" "210321",
    "012" "321",
    "012301230123" "032103210321",
};
static const char* const PRS_N[4] = {
    "120132232123123322211312" "311122102233021333303011",
    "232212" "022103",
    "230" "222",
    "011222033132" "010201222130",
};
static const char* const PRS_H[4] = {
    "0200001120002211", "0323013021232330", "0002021322022013", "0121033223212132",
};

/* src/ofdm/dab_prs_ref.cpp:140-194 */
int orc_get_prs(int mode, orc_c32* buf) {
    orc_params p;
    if (orc_get_params(mode, &p) != 0) return -1;
    const int nb_fft = p.nb_fft;
    const int half = p.nb_data_carriers / 2;
    for (int i = 0; i < nb_fft; i++) { buf[i].re = 0.0f; buf[i].im = 0.0f; }
    const char* ti = PRS_I[mode - 1];
    const char* tn = PRS_N[mode - 1];
    for (int c = 0; c < 2 * half; c++) {
        const int k = (c < half) ? (c - half) : (c - half + 1); /* carrier number, DC skipped */
        const int block = c / 32;
        const int h = PRS_H[ti[block] - '0'][(c % 32) % 16] - '0';
        const int n = tn[block] - '0';
        const float phi = (float)M_PI / 2.0f * (float)(h + n);
        orc_c32 v = { cosf(phi), sinf(phi) };
        buf[(k < 0) ? (nb_fft + k) : k] = v;
    }
    return 0;
}

/* src/ofdm/dab_mapper_ref.cpp:10-50 */
int orc_get_mapper(int mode, int* carrier_map) {
    orc_params p;
    if (orc_get_params(mode, &p) != 0) return -1;
    const int N = p.nb_fft, K = N / 4, nb_carriers = p.nb_data_carriers;
    const int dc = N / 2, start = dc - nb_carriers / 2, end = dc + nb_carriers / 2;
    int pi = 0, n_out = 0;
    for (int i = 0; i < N; i++) {
        if (i > 0) pi = (13 * pi + K - 1) % N;
        if (pi < start || pi > end || pi == dc) continue;
        carrier_map[n_out++] = (pi < dc) ? (pi - start) : (pi - start - 1);
    }
    return 0;
}

/* src/ofdm/ofdm_demodulator.h:24-45 */
void orc_default_config(orc_config* c) {
    c->signal_l1_update_beta = 0.95f;
    c->signal_l1_nb_samples = 100;
    c->signal_l1_nb_decimate = 5;
    c->thresh_null_start = 0.35f;
    c->thresh_null_end = 0.75f;
    c->fine_freq_update_beta = 0.9f;
    c->is_coarse_freq_correction = 1;
    c->max_coarse_freq_correction_norm = 0.5f;
    c->coarse_freq_slow_beta = 0.1f;
    c->impulse_peak_threshold_db = 20.0f;
    c->impulse_peak_distance_probability = 0.15f;
}

/* ------------------------------------------------------------------------------------------------ dsp */

/* src/ofdm/dsp/chebyshev_sine.h:13-41 with the FMA form of :78-98 */
static const float CHEB[6] = { -25.13274193f, 64.83583069f, -67.07687378f, 38.50016403f, -14.07150173f, 3.20396066f };

static float cheb_sine_fma(float x) {
    const float z = x * x;
    float b = CHEB[5];
    b = fmaf(b, z, CHEB[4]);
    b = fmaf(b, z, CHEB[3]);
    b = fmaf(b, z, CHEB[2]);
    b = fmaf(b, z, CHEB[1]);
    b = fmaf(b, z, CHEB[0]);
    return (b * x) * (z - 0.25f); /* association chosen by the reference build (gcc -ffast-math): (g*x)*(z-0.25) */
}

/* scalar tail of apply_pll.cpp:12-30 as that build evaluates it: ((z-0.25)*x)*g */
static float cheb_sine_fma_tail(float x) {
    const float z = x * x;
    float b = CHEB[5];
    b = fmaf(b, z, CHEB[4]);
    b = fmaf(b, z, CHEB[3]);
    b = fmaf(b, z, CHEB[2]);
    b = fmaf(b, z, CHEB[1]);
    b = fmaf(b, z, CHEB[0]);
    return ((z - 0.25f) * x) * b;
}

/* std::round under -ffast-math: trunc(t + copysign(0.49999997f, t)) */
static float fast_round(float t) { return truncf(t + copysignf(0.49999997f, t)); }

/* src/ofdm/dsp/apply_pll.cpp:82-116: 4 samples per AVX vector, phase = (dt_norm + float(i)*f) + (k*f [+0.25]),
 * wrapped by t - roundeven(t); rotation by c32_mul_avx with fmaddsub (x86/c32_mul.h:10-40).
 * The scalar tail (apply_pll.cpp:12-30) handles n % 4 samples. */
void orc_apply_pll(const orc_c32* x, orc_c32* y, size_t n, float freq_norm, float dt_norm) {
    const size_t n_vec = (n / 4u) * 4u;
    float pack_cos[4], pack_sin[4];
    for (int k = 0; k < 4; k++) {
        const float dt = (float)k * freq_norm;
        pack_cos[k] = dt + 0.25f;
        pack_sin[k] = dt;
    }
    for (size_t i = 0; i < n_vec; i += 4) {
        const float base = fmaf((float)i, freq_norm, dt_norm); /* contracted by the reference's -ffast-math build */
        for (int k = 0; k < 4; k++) {
            float tc = base + pack_cos[k];
            float ts = base + pack_sin[k];
            tc = tc - rintf(tc);
            ts = ts - rintf(ts);
            const float c = cheb_sine_fma(tc);
            const float s = cheb_sine_fma(ts);
            const orc_c32 v = x[i + (size_t)k];
            orc_c32 o;
            o.re = fmaf(c, v.re, -(s * v.im));
            o.im = fmaf(c, v.im, (s * v.re));
            y[i + (size_t)k] = o;
        }
    }
    const float dt_scalar = fmaf((float)n_vec, freq_norm, dt_norm);
    for (size_t i = n_vec; i < n; i++) {
        float ts = fmaf((float)(i - n_vec), freq_norm, dt_scalar);
        float tc = ts + 0.25f;
        ts = ts - fast_round(ts);
        tc = tc - fast_round(tc);
        const float c = cheb_sine_fma_tail(tc), s = cheb_sine_fma_tail(ts);
        const orc_c32 v = x[i];
        orc_c32 o;
        o.re = fmaf(v.re, c, -(v.im * s));
        o.im = fmaf(v.re, s, v.im * c);
        y[i] = o;
    }
}

/* src/ofdm/dsp/complex_conj_mul_sum.cpp:65-100 + x86/c32_conj_mul.h:11-43: sum x0*conj(x1), four AVX lanes
 * accumulated separately then folded (lane0+lane2)+(lane1+lane3). */
orc_c32 orc_conj_mul_sum(const orc_c32* x0, const orc_c32* x1, size_t n) {
    float acc_re[4] = {0, 0, 0, 0}, acc_im[4] = {0, 0, 0, 0};
    const size_t n_vec = (n / 4u) * 4u;
    for (size_t i = 0; i < n_vec; i += 4) {
        for (int k = 0; k < 4; k++) {
            const float a = x0[i + (size_t)k].re, b = x0[i + (size_t)k].im;
            const float c = x1[i + (size_t)k].re, d = x1[i + (size_t)k].im;
            acc_re[k] += fmaf(b, d, a * c);
            acc_im[k] += fmaf(b, c, -(a * d));
        }
    }
    orc_c32 y;
    y.re = (acc_re[0] + acc_re[2]) + (acc_re[1] + acc_re[3]);
    y.im = (acc_im[0] + acc_im[2]) + (acc_im[1] + acc_im[3]);
    /* scalar tail (complex_conj_mul_sum.cpp:12-25) summed from zero, then added to the vector part */
    float tr = 0.0f, ti = 0.0f;
    for (size_t i = n_vec; i < n; i++) {
        const float a = x0[i].re, b = x0[i].im, c = x1[i].re, d = x1[i].im;
        tr += fmaf(c, a, d * b);
        ti += fmaf(-d, a, c * b);
    }
    y.re += tr;
    y.im += ti;
    return y;
}

/* FFTW3 c2c semantics (ofdm_demodulator.cpp:113-114,891-899): unnormalised, sign=-1 forward / +1 backward, in-place ok.
 * FFTW3 itself is absent (vcpkg.json:19-21); evaluated in double and rounded once, like oracle/fftw3_shim/fftw3.h. */
typedef struct { int n; double* wr; double* wi; int* rev; } fft_plan;
static fft_plan g_plans[8];
static int g_n_plans = 0;
static pthread_mutex_t g_plan_mtx = PTHREAD_MUTEX_INITIALIZER;

static const fft_plan* get_plan(int n) {
    pthread_mutex_lock(&g_plan_mtx);
    for (int i = 0; i < g_n_plans; i++)
        if (g_plans[i].n == n) { pthread_mutex_unlock(&g_plan_mtx); return &g_plans[i]; }
    fft_plan* p = &g_plans[g_n_plans];
    p->n = n;
    p->wr = (double*)malloc(sizeof(double) * (size_t)n);
    p->wi = p->wr + n / 2;
    p->rev = (int*)malloc(sizeof(int) * (size_t)n);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < n / 2; k++) { /* forward twiddles e^{-2 pi j k / n}; backward uses the conjugate */
        const double a = two_pi * (double)k / (double)n;
        p->wr[k] = cos(a); p->wi[k] = sin(a);
    }
    int log2n = 0;
    while ((1 << log2n) < n) log2n++;
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < log2n; b++) r |= ((i >> b) & 1) << (log2n - 1 - b);
        p->rev[i] = r;
    }
    g_n_plans++;
    pthread_mutex_unlock(&g_plan_mtx);
    return p;
}

void orc_fft(const orc_c32* in, orc_c32* out, int n, int sign) {
    const fft_plan* pl = get_plan(n);
    double* xr = (double*)malloc(sizeof(double) * 2u * (size_t)n);
    double* xi = xr + n;
    for (int i = 0; i < n; i++) { xr[pl->rev[i]] = (double)in[i].re; xi[pl->rev[i]] = (double)in[i].im; }
    const double s = (double)sign;
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len / 2, step = n / len;
        for (int base = 0; base < n; base += len) {
            for (int k = 0; k < half; k++) {
                const double wr = pl->wr[k * step], wi = s * pl->wi[k * step];
                const int p = base + k, q = p + half;
                const double tr = xr[q] * wr - xi[q] * wi;
                const double ti = xr[q] * wi + xi[q] * wr;
                xr[q] = xr[p] - tr; xi[q] = xi[p] - ti;
                xr[p] = xr[p] + tr; xi[p] = xi[p] + ti;
            }
        }
    }
    for (int i = 0; i < n; i++) { out[i].re = (float)xr[i]; out[i].im = (float)xi[i]; }
    free(xr);
}

static orc_c32 c_mul(orc_c32 a, orc_c32 b) { orc_c32 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static orc_c32 c_conj(orc_c32 a) { orc_c32 r = { a.re, -a.im }; return r; }

/* ------------------------------------------------------------------------------------------------ modulator */

/* src/ofdm/ofdm_modulator.cpp:10-165 */
int orc_modulate(int mode, const uint8_t* bytes, size_t nbytes, orc_c32* out, size_t nsamples) {
    orc_params p;
    if (orc_get_params(mode, &p) != 0) return -1;
    const size_t frame_size = (size_t)p.nb_null_period + (size_t)p.nb_symbol_period * (size_t)p.nb_frame_symbols;
    const size_t data_size = (size_t)(p.nb_frame_symbols - 1) * (size_t)p.nb_data_carriers * 2u / 8u;
    if (nbytes != data_size || nsamples != frame_size) return -2;
    const int nfft = p.nb_fft, cp = p.nb_cyclic_prefix, half = p.nb_data_carriers / 2;
    orc_c32* last = (orc_c32*)calloc((size_t)nfft * 3u, sizeof(orc_c32));
    orc_c32* curr = last + nfft;
    orc_c32* prs = curr + nfft;
    orc_get_prs(mode, prs);
    for (int i = 0; i < p.nb_null_period; i++) { out[i].re = 0; out[i].im = 0; }
    orc_c32* sym = out + p.nb_null_period;
    orc_fft(prs, sym + cp, nfft, +1);
    for (int i = 0; i < cp; i++) sym[i] = sym[nfft + i];
    memcpy(last, prs, sizeof(orc_c32) * (size_t)nfft);
    const float A = 1.0f / sqrtf(2.0f);
    const orc_c32 PHASE_MAP[4] = { { -A, -A }, { A, -A }, { A, A }, { -A, A } };
    const size_t sym_bytes = (size_t)p.nb_data_carriers * 2u / 8u;
    for (int s = 0; s < p.nb_frame_symbols - 1; s++) {
        const uint8_t* d = bytes + (size_t)s * sym_bytes;
        sym += p.nb_symbol_period;
        int carrier = nfft - half;
        for (size_t i = 0; i < sym_bytes / 2; i++)
            for (int q = 0; q < 4; q++) curr[carrier++] = PHASE_MAP[(d[i] >> (2 * q)) & 3];
        carrier = 1;
        for (size_t i = 0; i < sym_bytes / 2; i++)
            for (int q = 0; q < 4; q++) curr[carrier++] = PHASE_MAP[(d[sym_bytes / 2 + i] >> (2 * q)) & 3];
        for (int i = 0; i < half; i++) { const int j = nfft - half + i; curr[j] = c_mul(last[j], curr[j]); }
        for (int i = 0; i < half; i++) { const int j = 1 + i; curr[j] = c_mul(last[j], curr[j]); }
        orc_fft(curr, sym + cp, nfft, +1);
        for (int i = 0; i < cp; i++) sym[i] = sym[nfft + i];
        orc_c32* t = last; last = curr; curr = t;
    }
    free(last < curr ? last : curr);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ OFDM demodulator */

enum { ST_FINDING_NULL_POWER_DIP = 0, ST_READING_NULL_AND_PRS, ST_RUNNING_COARSE_FREQ_SYNC, ST_RUNNING_FINE_TIME_SYNC, ST_READING_SYMBOLS };

struct orc_ofdm {
    orc_params p;
    orc_config cfg;
    int state;
    int total_frames_read, total_frames_desync;
    int is_found_coarse;
    float freq_coarse, freq_fine;
    int fine_time_offset;
    int null_start_found, null_end_found;
    float l1_average;
    /* src/ofdm/circular_buffer.h */
    orc_c32* ring; size_t ring_index, ring_length;
    /* src/ofdm/reconstruction_buffer.h: null + PRS */
    orc_c32* corr; size_t corr_length;
    /* src/ofdm/ofdm_frame_buffer.h (no alignment padding needed here): PRS + data symbols + NULL */
    orc_c32* frame; size_t frame_fill, frame_cap;
    orc_c32 *prs_fft_ref_conj, *prs_time_ref_conj;
    orc_c32 *fft_buf, *ifft_buf;
    float *impulse_response, *freq_response;
    int* mapper;
    orc_c32 *pipe_fft, *pipe_vec;
    int8_t* bits; size_t n_bits;
    int64_t abs_consumed;
    orc_frame_info pending;
    orc_frame_cb cb; void* cb_user;
    /* collected frames */
    int8_t** frames; orc_frame_info* infos; size_t n_frames, cap_frames;
};

static float l1_average(const orc_c32* b, size_t n) { /* ofdm_demodulator.cpp:922-932 */
    float acc = 0.0f;
    for (size_t i = 0; i < n; i++) acc += fabsf(b[i].re) + fabsf(b[i].im);
    return acc / (float)n;
}

static void update_fine_offset(orc_ofdm* d, float delta) { /* ofdm_demodulator.cpp:829-840 */
    const float spacing = 1.0f / (float)d->p.nb_fft;
    const float wrap = 0.5f * spacing * 1.01f;
    d->freq_fine += delta;
    d->freq_fine = fmodf(d->freq_fine, wrap);
}

orc_ofdm* orc_ofdm_create_custom(const orc_params* pp, const orc_c32* prs_fft_ref, const int* mapper) {
    orc_ofdm* d = (orc_ofdm*)calloc(1, sizeof(orc_ofdm));
    d->p = *pp;
    const orc_params* p = &d->p;
    orc_default_config(&d->cfg);
    const size_t nfft = (size_t)p->nb_fft;
    d->ring = (orc_c32*)calloc((size_t)p->nb_null_period, sizeof(orc_c32));
    d->corr = (orc_c32*)calloc((size_t)(p->nb_null_period + p->nb_symbol_period), sizeof(orc_c32));
    d->frame_cap = (size_t)p->nb_frame_symbols * (size_t)p->nb_symbol_period + (size_t)p->nb_null_period;
    d->frame = (orc_c32*)calloc(d->frame_cap, sizeof(orc_c32));
    d->prs_fft_ref_conj = (orc_c32*)calloc(nfft, sizeof(orc_c32));
    d->prs_time_ref_conj = (orc_c32*)calloc(nfft, sizeof(orc_c32));
    d->fft_buf = (orc_c32*)calloc(nfft, sizeof(orc_c32));
    d->ifft_buf = (orc_c32*)calloc(nfft, sizeof(orc_c32));
    d->impulse_response = (float*)calloc(nfft, sizeof(float));
    d->freq_response = (float*)calloc(nfft, sizeof(float));
    d->mapper = (int*)calloc((size_t)p->nb_data_carriers, sizeof(int));
    d->pipe_fft = (orc_c32*)calloc((size_t)(p->nb_frame_symbols + 1) * nfft, sizeof(orc_c32));
    d->pipe_vec = (orc_c32*)calloc((size_t)(p->nb_frame_symbols - 1) * nfft, sizeof(orc_c32));
    d->n_bits = (size_t)(p->nb_frame_symbols - 1) * (size_t)p->nb_data_carriers * 2u;
    d->bits = (int8_t*)calloc(d->n_bits, 1);
    d->state = ST_FINDING_NULL_POWER_DIP;
    /* ofdm_demodulator.cpp:128-143 */
    for (size_t i = 0; i < nfft; i++) d->prs_fft_ref_conj[i] = c_conj(prs_fft_ref[i]);
    for (size_t i = 0; i + 1 < nfft; i++) d->prs_time_ref_conj[i] = c_mul(c_conj(prs_fft_ref[i]), prs_fft_ref[i + 1]);
    d->prs_time_ref_conj[nfft - 1].re = 0; d->prs_time_ref_conj[nfft - 1].im = 0;
    orc_fft(d->prs_time_ref_conj, d->prs_time_ref_conj, p->nb_fft, +1);
    for (size_t i = 0; i < nfft; i++) d->prs_time_ref_conj[i] = c_conj(d->prs_time_ref_conj[i]);
    memcpy(d->mapper, mapper, sizeof(int) * (size_t)p->nb_data_carriers);
    return d;
}

orc_ofdm* orc_ofdm_create(int mode) {
    orc_params p;
    if (orc_get_params(mode, &p) != 0) return NULL;
    orc_c32* prs = (orc_c32*)calloc((size_t)p.nb_fft, sizeof(orc_c32));
    int* map = (int*)calloc((size_t)p.nb_data_carriers, sizeof(int));
    orc_get_prs(mode, prs);
    orc_get_mapper(mode, map);
    orc_ofdm* d = orc_ofdm_create_custom(&p, prs, map);
    free(prs); free(map);
    return d;
}

void orc_ofdm_destroy(orc_ofdm* d) {
    if (!d) return;
    for (size_t i = 0; i < d->n_frames; i++) free(d->frames[i]);
    free(d->frames); free(d->infos);
    free(d->ring); free(d->corr); free(d->frame); free(d->prs_fft_ref_conj); free(d->prs_time_ref_conj);
    free(d->fft_buf); free(d->ifft_buf); free(d->impulse_response); free(d->freq_response); free(d->mapper);
    free(d->pipe_fft); free(d->pipe_vec); free(d->bits);
    free(d);
}

orc_config* orc_ofdm_config(orc_ofdm* d) { return &d->cfg; }
void orc_ofdm_set_callback(orc_ofdm* d, orc_frame_cb cb, void* user) { d->cb = cb; d->cb_user = user; }
size_t orc_ofdm_frames_done(const orc_ofdm* d) { return (size_t)d->total_frames_read; }
size_t orc_ofdm_frame_bits(const orc_ofdm* d) { return d->n_bits; }
const orc_c32* orc_ofdm_frame_fft(const orc_ofdm* d) { return d->pipe_fft; }
const orc_c32* orc_ofdm_frame_data_vec(const orc_ofdm* d) { return d->pipe_vec; }
const float* orc_ofdm_impulse_response(const orc_ofdm* d) { return d->impulse_response; }
const float* orc_ofdm_coarse_freq_response(const orc_ofdm* d) { return d->freq_response; }
const orc_c32* orc_ofdm_correlation_time_buffer(const orc_ofdm* d, size_t* length) {
    if (length) *length = d->corr_length;
    return d->corr;
}

int orc_ofdm_get_frame(const orc_ofdm* d, size_t index, orc_frame_info* info, int8_t* bits_out) {
    if (index >= d->n_frames) return -1;
    if (info) *info = d->infos[index];
    if (bits_out) memcpy(bits_out, d->frames[index], d->n_bits);
    return 0;
}

void orc_ofdm_get_state(const orc_ofdm* d, orc_ofdm_state* s) {
    s->state = d->state; s->fine_time_offset = d->fine_time_offset;
    s->total_frames_read = d->total_frames_read; s->total_frames_desync = d->total_frames_desync;
    s->signal_average = d->l1_average; s->fine_offset = d->freq_fine; s->coarse_offset = d->freq_coarse; s->pad = 0;
}

void orc_ofdm_reset(orc_ofdm* d) { /* ofdm_demodulator.cpp:277-289 */
    d->state = ST_FINDING_NULL_POWER_DIP;
    d->corr_length = 0;
    d->total_frames_desync++;
    d->is_found_coarse = 0;
    d->freq_coarse = 0; d->freq_fine = 0; d->fine_time_offset = 0;
}

static void update_signal_average(orc_ofdm* d, const orc_c32* block, size_t N) { /* ofdm_demodulator.cpp:934-950 */
    const size_t K = (size_t)d->cfg.signal_l1_nb_samples;
    if (N < K) return;
    const size_t M = N - K, L = K * (size_t)d->cfg.signal_l1_nb_decimate;
    const float beta = d->cfg.signal_l1_update_beta;
    for (size_t i = 0; i < M; i += L) {
        const float l1 = l1_average(block + i, K);
        d->l1_average = beta * d->l1_average + (1.0f - beta) * l1;
    }
}

static size_t find_null_power_dip(orc_ofdm* d, const orc_c32* buf, size_t n) { /* ofdm_demodulator.cpp:291-347 */
    const int N = (int)n, K = d->cfg.signal_l1_nb_samples, M = N - K;
    const float start_thresh = d->l1_average * d->cfg.thresh_null_start;
    const float end_thresh = d->l1_average * d->cfg.thresh_null_end;
    int nb_read = N;
    for (int i = 0; i < M; i += K) {
        const float l1 = l1_average(buf + i, (size_t)K);
        if (d->null_start_found) {
            if (l1 > end_thresh) { d->null_end_found = 1; nb_read = i + K; break; }
        } else if (l1 < start_thresh) {
            d->null_start_found = 1;
        }
    }
    /* circular_buffer.h:18-38, read_all = true */
    const size_t cap = (size_t)d->p.nb_null_period;
    for (int i = 0; i < nb_read; i++) { d->ring[d->ring_index++] = buf[i]; d->ring_index %= cap; }
    d->ring_length += (size_t)nb_read;
    if (d->ring_length > cap) d->ring_length = cap;
    if (!d->null_end_found) return (size_t)nb_read;
    const size_t L = d->ring_length, start = d->ring_index;
    for (size_t i = 0; i < L; i++) d->corr[i] = d->ring[(i + start) % cap];
    d->null_start_found = 0; d->null_end_found = 0;
    d->corr_length = L;
    d->ring_length = 0;
    d->state = ST_READING_NULL_AND_PRS;
    return (size_t)nb_read;
}

static size_t read_null_prs(orc_ofdm* d, const orc_c32* buf, size_t n) { /* ofdm_demodulator.cpp:349-358 */
    const size_t cap = (size_t)(d->p.nb_null_period + d->p.nb_symbol_period);
    const size_t need = cap - d->corr_length;
    const size_t nb_read = (need >= n) ? n : need;
    memcpy(d->corr + d->corr_length, buf, nb_read * sizeof(orc_c32));
    d->corr_length += nb_read;
    if (d->corr_length == cap) d->state = ST_RUNNING_COARSE_FREQ_SYNC;
    return nb_read;
}

static void run_coarse_freq_sync(orc_ofdm* d) { /* ofdm_demodulator.cpp:360-471 */
    if (!d->cfg.is_coarse_freq_correction) { d->freq_coarse = 0; d->state = ST_RUNNING_FINE_TIME_SYNC; return; }
    const int nfft = d->p.nb_fft;
    const orc_c32* prs_sym = d->corr + d->p.nb_null_period;
    orc_fft(prs_sym, d->fft_buf, nfft, -1);
    for (int i = 0; i < nfft - 1; i++) d->fft_buf[i] = c_mul(c_conj(d->fft_buf[i]), d->fft_buf[i + 1]); /* :901-909 */
    d->fft_buf[nfft - 1].re = 0; d->fft_buf[nfft - 1].im = 0;
    orc_fft(d->fft_buf, d->ifft_buf, nfft, +1);
    for (int i = 0; i < nfft; i++) d->ifft_buf[i] = c_mul(d->ifft_buf[i], d->prs_time_ref_conj[i]);
    orc_fft(d->ifft_buf, d->fft_buf, nfft, -1);
    const int M = nfft / 2;
    for (int i = 0; i < nfft; i++) { /* :911-920 */
        const orc_c32 v = d->fft_buf[(i + M) % nfft];
        d->freq_response[i] = 20.0f * log10f(sqrtf(v.re * v.re + v.im * v.im));
    }
    int max_off = (int)(d->cfg.max_coarse_freq_correction_norm * (float)nfft);
    if (max_off < 0) max_off = 0;
    if (max_off > M) max_off = M;
    int max_index = -max_off;
    float max_value = d->freq_response[max_index + M];
    for (int i = -max_off; i <= max_off; i++) {
        const int fi = i + M;
        if (fi == nfft) continue;
        if (d->freq_response[fi] > max_value) { max_value = d->freq_response[fi]; max_index = i; }
    }
    float mag[3]; int idx[3];
    for (int k = 0; k < 3; k++) {
        int index = max_index - 1 + k;
        if (index < -max_off) index = -max_off;
        if (index > max_off) index = max_off;
        int fi = index + M;
        if (fi >= nfft) fi = nfft - 1;
        mag[k] = powf(10.0f, d->freq_response[fi] / 20.0f);
        idx[k] = fi - M;
    }
    float peak_sum = 0.0f, lerp = 0.0f;
    for (int k = 0; k < 3; k++) peak_sum += mag[k];
    for (int k = 0; k < 3; k++) lerp += (float)idx[k] * mag[k] / peak_sum;
    const float predicted = -lerp / (float)nfft;
    const float error = predicted - d->freq_coarse;
    const float large_thresh = 1.5f / (float)nfft;
    const int is_fast = (fabsf(error) > large_thresh) || !d->is_found_coarse;
    const float beta = is_fast ? 1.0f : d->cfg.coarse_freq_slow_beta;
    const float delta = beta * error;
    d->freq_coarse += delta;
    d->is_found_coarse = 1;
    update_fine_offset(d, -delta);
    d->state = ST_RUNNING_FINE_TIME_SYNC;
}

static void frame_consume_reset(orc_ofdm* d) { d->frame_fill = 0; }

static void run_fine_time_sync(orc_ofdm* d) { /* ofdm_demodulator.cpp:473-548 */
    const int nfft = d->p.nb_fft, cp = d->p.nb_cyclic_prefix, sp = d->p.nb_symbol_period, np = d->p.nb_null_period;
    const float freq_offset = d->freq_coarse + d->freq_fine;
    orc_apply_pll(d->corr + np, d->ifft_buf, (size_t)nfft, freq_offset, 0.0f);
    orc_fft(d->ifft_buf, d->fft_buf, nfft, -1);
    for (int i = 0; i < nfft; i++) d->fft_buf[i] = c_mul(d->fft_buf[i], d->prs_fft_ref_conj[i]);
    orc_fft(d->fft_buf, d->ifft_buf, nfft, +1);
    for (int i = 0; i < nfft; i++) {
        const orc_c32 v = d->ifft_buf[i];
        d->impulse_response[i] = 20.0f * log10f(sqrtf(v.re * v.re + v.im * v.im));
    }
    float avg = 0.0f, max_value = d->impulse_response[0];
    int max_index = 0;
    const float decay = 1.0f - d->cfg.impulse_peak_distance_probability;
    for (int i = 0; i < nfft; i++) {
        const float peak = d->impulse_response[i];
        const int dist = abs(cp - i);
        const float norm_dist = (float)dist / (float)sp;
        const float prob = 1.0f - decay * norm_dist;
        const float weighted = prob * peak;
        avg += peak;
        if (weighted > max_value) { max_value = weighted; max_index = i; }
    }
    avg /= (float)nfft;
    if ((max_value - avg) < d->cfg.impulse_peak_threshold_db) { orc_ofdm_reset(d); return; }
    const int offset = max_index - cp;
    const int prs_start = np + offset, prs_len = sp - offset;
    frame_consume_reset(d);
    memcpy(d->frame, d->corr + prs_start, (size_t)prs_len * sizeof(orc_c32));
    d->frame_fill = (size_t)prs_len;
    d->corr_length = 0;
    d->fine_time_offset = offset;
    d->state = ST_READING_SYMBOLS;
}

/* ofdm_demodulator.cpp:867-889 + 57-72: deinterleave, L-infinity normalise, quantise with C truncation */
static void viterbi_bits(const orc_c32* vec, const int* mapper, int N, int8_t* bits) {
    for (int i = 0; i < N; i++) {
        const orc_c32 v = vec[mapper[i]];
        const float A = fmaxf(fabsf(v.re), fabsf(v.im));
        /* the reference build keeps a true division here (verified on its output: the max component is exactly +-127) */
        const float nre = v.re / A, nim = v.im / A;
        bits[i] = (int8_t)(-(nre) * 127.0f);
        bits[i + N] = (int8_t)(-(-nim) * 127.0f);
    }
}

/* ofdm_demodulator.cpp:842-865 (called with (fft[i+1], fft[i]) at :736, i.e. out = X_i * conj(X_{i+1})) */
static void dqpsk(const orc_c32* x_next, const orc_c32* x_curr, int nfft, int ncarriers, orc_c32* out) {
    const int M = ncarriers / 2;
    int j = 0;
    for (int i = -M; i <= M; i++) {
        if (i == 0) continue;
        const int fi = (nfft + i) % nfft;
        out[j++] = c_mul(x_curr[fi], c_conj(x_next[fi]));
    }
}

static float cyclic_phase_error(const orc_c32* sym, int nfft, int cp) { /* ofdm_demodulator.cpp:768-777 */
    const orc_c32 e = orc_conj_mul_sum(sym + nfft, sym, (size_t)cp);
    return atan2f(e.im, e.re);
}

/* ofdm_demodulator.cpp:650-766 for one pipeline covering every symbol, on an already aligned frame.
 * frame_fft / data_vec may be NULL. */
static void demod_frame(const orc_params* p, const int* mapper, orc_c32* frame, int n_syms_pll, float freq_offset, int8_t* bits,
                        float* phase_error_sum, orc_c32* frame_fft, orc_c32* data_vec) {
    const int S = p->nb_frame_symbols, sp = p->nb_symbol_period, nfft = p->nb_fft, cp = p->nb_cyclic_prefix, nc = p->nb_data_carriers;
    for (int i = 0; i < n_syms_pll; i++) {
        const float dt_start = (float)(i * sp) * freq_offset;
        orc_apply_pll(frame + (size_t)i * (size_t)sp, frame + (size_t)i * (size_t)sp, (size_t)sp, freq_offset, dt_start);
    }
    float total = 0.0f;
    for (int i = 0; i < S; i++) total += cyclic_phase_error(frame + (size_t)i * (size_t)sp, nfft, cp);
    *phase_error_sum = total;
    orc_c32* fft_local = frame_fft ? frame_fft : (orc_c32*)malloc(sizeof(orc_c32) * (size_t)nfft * (size_t)(S + 1));
    orc_c32* vec_local = data_vec ? data_vec : (orc_c32*)malloc(sizeof(orc_c32) * (size_t)nc);
    for (int i = 0; i < n_syms_pll; i++) orc_fft(frame + (size_t)i * (size_t)sp + cp, fft_local + (size_t)i * (size_t)nfft, nfft, -1);
    for (int i = 0; i < S - 1; i++) {
        orc_c32* vec = data_vec ? (data_vec + (size_t)i * (size_t)nc) : vec_local;
        dqpsk(fft_local + (size_t)(i + 1) * (size_t)nfft, fft_local + (size_t)i * (size_t)nfft, nfft, nc, vec);
        viterbi_bits(vec, mapper, nc, bits + (size_t)i * (size_t)nc * 2u);
    }
    if (!frame_fft) free(fft_local);
    if (!data_vec) free(vec_local);
}

void orc_ofdm_demod_frame(const orc_params* p, const int* mapper, const orc_c32* frame, float freq_offset, int8_t* bits_out,
                          float* phase_error_sum) {
    const size_t n = (size_t)p->nb_frame_symbols * (size_t)p->nb_symbol_period;
    orc_c32* tmp = (orc_c32*)malloc(n * sizeof(orc_c32));
    memcpy(tmp, frame, n * sizeof(orc_c32));
    demod_frame(p, mapper, tmp, p->nb_frame_symbols, freq_offset, bits_out, phase_error_sum, NULL, NULL);
    free(tmp);
}

static void run_pipeline(orc_ofdm* d) { /* CoordinatorThread :581-639 + PipelineThread :650-766, real-time order */
    const orc_params* p = &d->p;
    const float freq_offset = d->freq_coarse + d->freq_fine;
    d->pending.coarse_offset = d->freq_coarse;
    d->pending.fine_offset_used = d->freq_fine;
    d->pending.signal_average = d->l1_average;
    d->pending.total_desync = d->total_frames_desync;
    float total_phase_error = 0.0f;
    /* PLL and FFT also run on the first nb_symbol_period samples of the NULL symbol (symbol index S), :673-678,:701-724 */
    demod_frame(p, d->mapper, d->frame, p->nb_frame_symbols + 1, freq_offset, d->bits, &total_phase_error, d->pipe_fft, d->pipe_vec);
    const float avg_error = total_phase_error / (float)p->nb_frame_symbols;
    const float two_pi = (float)M_PI * 2.0f;
    const float fine_error = (1.0f / (float)p->nb_fft) * avg_error / two_pi; /* :779-824 */
    update_fine_offset(d, -d->cfg.fine_freq_update_beta * fine_error);
    d->pending.fine_offset_after = d->freq_fine;
    d->total_frames_read++;
    if (d->cb) {
        d->cb(d->cb_user, d->bits, d->n_bits, &d->pending);
    } else {
        if (d->n_frames == d->cap_frames) {
            d->cap_frames = d->cap_frames ? d->cap_frames * 2 : 16;
            d->frames = (int8_t**)realloc(d->frames, d->cap_frames * sizeof(int8_t*));
            d->infos = (orc_frame_info*)realloc(d->infos, d->cap_frames * sizeof(orc_frame_info));
        }
        d->frames[d->n_frames] = (int8_t*)malloc(d->n_bits);
        memcpy(d->frames[d->n_frames], d->bits, d->n_bits);
        d->infos[d->n_frames] = d->pending;
        d->n_frames++;
    }
}

static size_t read_symbols(orc_ofdm* d, const orc_c32* buf, size_t n) { /* ofdm_demodulator.cpp:550-577 */
    const size_t need = d->frame_cap - d->frame_fill;
    const size_t nb_read = (n > need) ? need : n;
    memcpy(d->frame + d->frame_fill, buf, nb_read * sizeof(orc_c32));
    d->frame_fill += nb_read;
    if (d->frame_fill != d->frame_cap) return nb_read;
    const size_t np = (size_t)d->p.nb_null_period;
    memcpy(d->corr, d->frame + d->frame_cap - np, np * sizeof(orc_c32));
    d->corr_length = np;
    run_pipeline(d);
    d->frame_fill = 0;
    d->state = ST_READING_NULL_AND_PRS;
    return nb_read;
}

void orc_ofdm_process(orc_ofdm* d, const orc_c32* buf, size_t N) { /* ofdm_demodulator.cpp:235-275 */
    update_signal_average(d, buf, N);
    size_t curr = 0;
    while (curr < N) {
        const orc_c32* block = buf + curr;
        const size_t remain = N - curr;
        switch (d->state) {
        case ST_FINDING_NULL_POWER_DIP: curr += find_null_power_dip(d, block, remain); break;
        case ST_READING_NULL_AND_PRS: curr += read_null_prs(d, block, remain); break;
        case ST_RUNNING_COARSE_FREQ_SYNC: run_coarse_freq_sync(d); break;
        case ST_RUNNING_FINE_TIME_SYNC:
            run_fine_time_sync(d);
            if (d->state == ST_READING_SYMBOLS) {
                d->pending.fine_time_offset = d->fine_time_offset;
                d->pending.frame_start = d->abs_consumed + (int64_t)curr - (int64_t)d->p.nb_symbol_period + (int64_t)d->fine_time_offset;
            }
            break;
        case ST_READING_SYMBOLS: curr += read_symbols(d, block, remain); break;
        }
    }
    d->abs_consumed += (int64_t)N;
}

/* ------------------------------------------------------------------------------------------------ Viterbi */

enum { VK = 7, VR = 4, VSTATES = 64 };
static const uint8_t V_POLY[VR] = { 109, 79, 83, 109 }; /* dab_viterbi_decoder.cpp:24 */
#define V_MAX_ERROR 1016u      /* (127 - -127) * 4, dab_viterbi_decoder.cpp:31 */
#define V_NON_START 5080u      /* 5 * max_error, :32-36 */
#define V_RENORM_THRESHOLD 60455u /* 65535 - 5080, :37 */

struct orc_viterbi {
    uint16_t metric[2][VSTATES];
    int cur;
    uint64_t* decisions; size_t decisions_len; /* traceback_length + 6 */
    size_t current_decoded_bit;
    uint64_t accumulated_error;
    int16_t* depunct; size_t depunct_cap;
    int16_t branch[VR][VSTATES / 2];
};

static int parity(unsigned v) { v ^= v >> 4; v ^= v >> 2; v ^= v >> 1; return (int)(v & 1u); }

orc_viterbi* orc_vit_create(void) {
    orc_viterbi* v = (orc_viterbi*)calloc(1, sizeof(orc_viterbi));
    for (int s = 0; s < VSTATES / 2; s++) /* viterbi_branch_table.h:34-55 */
        for (int r = 0; r < VR; r++) v->branch[r][s] = parity(((unsigned)s << 1) & V_POLY[r]) ? 127 : -127;
    orc_vit_reset(v, 0);
    orc_vit_set_traceback_length(v, 0);
    return v;
}

void orc_vit_destroy(orc_viterbi* v) { if (v) { free(v->decisions); free(v->depunct); free(v); } }

void orc_vit_set_traceback_length(orc_viterbi* v, size_t n) { /* viterbi_decoder_core.h:180-187 */
    const size_t new_len = n + (VK - 1);
    v->decisions = (uint64_t*)realloc(v->decisions, new_len * sizeof(uint64_t));
    v->decisions_len = new_len;
    if (v->current_decoded_bit > new_len) v->current_decoded_bit = new_len;
}
size_t orc_vit_get_traceback_length(const orc_viterbi* v) { return v->decisions_len - (VK - 1); }
size_t orc_vit_get_current_decoded_bit(const orc_viterbi* v) { return v->current_decoded_bit; }

void orc_vit_reset(orc_viterbi* v, size_t start_state) { /* viterbi_decoder_core.h:202-211, dab_viterbi_decoder.cpp:109-112 */
    v->current_decoded_bit = 0;
    v->cur = 0;
    for (int i = 0; i < VSTATES; i++) v->metric[0][i] = V_NON_START;
    v->metric[0][start_state & (VSTATES - 1)] = 0;
    v->accumulated_error = 0;
}

static uint16_t sat_add_u16(uint16_t a, uint16_t b) { const unsigned s = (unsigned)a + (unsigned)b; return (uint16_t)(s > 65535u ? 65535u : s); }
static uint16_t sat_sub_u16(uint16_t a, uint16_t b) { return (uint16_t)(a > b ? a - b : 0); }
static int16_t sat_sub_s16(int16_t a, int16_t b) { int d = (int)a - (int)b; if (d > 32767) d = 32767; if (d < -32768) d = -32768; return (int16_t)d; }

/* viterbi_decoder_avx_u16.h:47-71 (update), :73-136 (bfly), :138-170 (renormalise) */
static uint64_t vit_steps(orc_viterbi* v, const int16_t* sym, size_t n_sym) {
    uint64_t total_error = 0;
    for (size_t s = 0; s < n_sym; s += VR) {
        const uint16_t* old = v->metric[v->cur];
        uint16_t* neu = v->metric[1 - v->cur];
        uint64_t dec = 0;
        for (int b = 0; b < VSTATES / 2; b++) {
            uint16_t e = 0;
            for (int r = 0; r < VR; r++) {
                int16_t err = sat_sub_s16(v->branch[r][b], sym[s + (size_t)r]);
                if (err < 0) err = (int16_t)(err == -32768 ? -32768 : -err); /* _mm256_abs_epi16 */
                e = sat_add_u16(e, (uint16_t)err);
            }
            const uint16_t inv = sat_sub_u16(V_MAX_ERROR, e);
            const uint16_t e00 = sat_add_u16(old[b], e), e10 = sat_add_u16(old[b + 32], inv);
            const uint16_t e01 = sat_add_u16(old[b], inv), e11 = sat_add_u16(old[b + 32], e);
            const uint16_t m0 = e00 < e10 ? e00 : e10, m1 = e01 < e11 ? e01 : e11;
            neu[2 * b] = m0; neu[2 * b + 1] = m1;
            dec |= (uint64_t)(m0 == e10) << (2 * b);       /* tie -> 1 */
            dec |= (uint64_t)(m1 == e11) << (2 * b + 1);
        }
        v->decisions[v->current_decoded_bit] = dec;
        if (neu[0] >= V_RENORM_THRESHOLD) {
            uint16_t mn = neu[0];
            for (int i = 1; i < VSTATES; i++) if (neu[i] < mn) mn = neu[i];
            for (int i = 0; i < VSTATES; i++) neu[i] = sat_sub_u16(neu[i], mn);
            total_error += mn;
        }
        v->cur = 1 - v->cur;
        v->current_decoded_bit++;
    }
    return total_error;
}

/* dab_viterbi_decoder.cpp:114-122 + 131-181 */
size_t orc_vit_update(orc_viterbi* v, const int8_t* soft, size_t n_soft, const uint8_t* code, size_t code_len, size_t n_out) {
    if (n_out > v->depunct_cap) { v->depunct = (int16_t*)realloc(v->depunct, n_out * sizeof(int16_t)); v->depunct_cap = n_out; }
    size_t ip = 0, ic = 0, io = 0;
    while (io < n_out) {
        const size_t take = code[ic];
        if (n_soft - ip < take) return 0; /* underrun: the reference returns a zeroed result and decodes nothing (:158-162) */
        for (size_t i = 0; i < take; i++) v->depunct[io++] = (int16_t)soft[ip++];
        for (size_t i = take; i < VR; i++) v->depunct[io++] = 0;
        ic = (ic + 1) % code_len;
    }
    v->accumulated_error += vit_steps(v, v->depunct, io);
    return ip;
}

/* viterbi_decoder_core.h:214-236 with ViterbiTracebackBuffer<7> (:88-153): 8-bit register, state = reg >> 2 */
uint64_t orc_vit_chainback(orc_viterbi* v, uint8_t* out, size_t nbytes, size_t end_state) {
    const size_t total_bits = nbytes * 8u;
    unsigned reg = (unsigned)(end_state << 2);
    for (size_t i = 0; i < total_bits; i++) {
        const size_t j = (total_bits - 1) - i;
        const uint64_t dec = v->decisions[j + (VK - 1)];
        const unsigned state = reg >> 2;
        const unsigned bit = (unsigned)((dec >> state) & 1u);
        reg = (reg >> 1) | (bit << 7);
        out[j / 8] = (uint8_t)(reg & 0xFFu);
    }
    return v->accumulated_error + (uint64_t)v->metric[v->cur][0]; /* dab_viterbi_decoder.cpp:124-129 */
}

uint64_t orc_vit_decode_job(orc_viterbi* v, const int8_t* soft, size_t n_soft, const uint8_t* seg_codes, const uint32_t* seg_code_len,
                            const uint32_t* seg_n_out, uint32_t n_seg, uint8_t* out, size_t n_out_bytes, size_t* consumed) {
    orc_vit_reset(v, 0);
    size_t used = 0;
    for (uint32_t s = 0; s < n_seg; s++)
        used += orc_vit_update(v, soft + used, n_soft - used, seg_codes + 8u * s, seg_code_len[s], seg_n_out[s]);
    if (consumed) *consumed = used;
    return orc_vit_chainback(v, out, n_out_bytes, 0);
}

/* src/dab/constants/puncture_codes.h:42-74 as a rule: PI_p keeps base = (p-1)/8 + 1 bits in each of the 8 groups of 4
 * mother bits, plus one more in the first ((p-1)%8)+1 groups taken in bit-reversed order 0,4,2,6,1,5,3,7. */
static uint8_t PI_COUNTS[24][8];
static const uint8_t PI_TAIL[6] = { 2, 2, 2, 2, 2, 2 };
static pthread_once_t pi_once = PTHREAD_ONCE_INIT;
static void pi_init(void) {
    for (int p = 1; p <= 24; p++) {
        const int base = (p - 1) / 8 + 1, extra = ((p - 1) % 8) + 1;
        for (int g = 0; g < 8; g++) {
            const int rev = ((g & 1) << 2) | (g & 2) | ((g & 4) >> 2);
            PI_COUNTS[p - 1][g] = (uint8_t)(base + (rev < extra ? 1 : 0));
        }
    }
}
const uint8_t* orc_puncture_code(int pi) { pthread_once(&pi_once, pi_init); return (pi >= 1 && pi <= 24) ? PI_COUNTS[pi - 1] : NULL; }
const uint8_t* orc_puncture_code_tail(void) { return PI_TAIL; }

/* Mother code K=7, G = {109,79,83,109}: convolutional_encoder_shift_register.h:44-62 (MSB-first input bits, register
 * shifted left, output bit r = parity(G[r] & reg)), followed by 6 zero tail bits (helpers/test_helpers.h:52-59).
 * One soft symbol (+127 for 1, -127 for 0) per output bit. */
size_t orc_conv_encode(const uint8_t* bytes, size_t nbytes, int8_t* soft_out) {
    unsigned reg = 0;
    size_t n = 0;
    const size_t total_bits = nbytes * 8u + (VK - 1);
    for (size_t i = 0; i < total_bits; i++) {
        const unsigned bit = (i < nbytes * 8u) ? ((bytes[i / 8] >> (7 - (i % 8))) & 1u) : 0u;
        reg = ((reg << 1) | bit) & 0x7Fu;
        for (int r = 0; r < VR; r++) soft_out[n++] = parity(reg & V_POLY[r]) ? 127 : -127;
    }
    return n;
}

/* Transmit-side puncturing matching depuncture_symbols (dab_viterbi_decoder.cpp:131-181) */
size_t orc_puncture(const int8_t* mother, size_t n_mother, const uint8_t* seg_codes, const uint32_t* seg_code_len, const uint32_t* seg_n_out,
                    uint32_t n_seg, int8_t* out) {
    size_t im = 0, io = 0;
    for (uint32_t s = 0; s < n_seg; s++) {
        size_t ic = 0;
        for (size_t produced = 0; produced < seg_n_out[s]; produced += VR) {
            const size_t take = seg_codes[8u * s + ic];
            for (size_t i = 0; i < take && (im + i) < n_mother; i++) out[io++] = mother[im + i];
            im += VR;
            ic = (ic + 1) % seg_code_len[s];
        }
    }
    return io;
}

/* ------------------------------------------------------------------------------------------------ FIC / MSC decode
 * what sits between the OFDM soft bits and the decoded bytes, either side of the Viterbi decoder. */

/* src/dab/algorithms/additive_scrambler.h:10-35: G(x) = 1 + x^-5 + x^-9, register reloaded with the syncword on Reset() */
void orc_scrambler_bytes(uint16_t syncword, uint8_t* out, size_t n) {
    uint16_t reg = syncword;
    for (size_t k = 0; k < n; k++) {
        uint8_t b = 0;
        for (int i = 0; i < 8; i++) {
            const uint8_t v = (uint8_t)(((reg >> 8) ^ (reg >> 4)) & 1u);
            b |= (uint8_t)(v << (7 - i));
            reg = (uint16_t)((reg << 1) | v);
        }
        out[k] = b;
    }
}

/* src/dab/algorithms/crc.h:24-33 with the FIB parameters of fic_decoder.cpp:20-33: poly 0x1021, init 0xFFFF, final xor 0xFFFF */
uint16_t orc_crc16_fib(const uint8_t* x, size_t n) {
    uint16_t crc = 0xFFFF;
    for (size_t i = 0; i < n; i++) {
        crc ^= (uint16_t)((uint16_t)x[i] << 8);
        for (int j = 0; j < 8; j++) crc = (uint16_t)((crc & 0x8000u) ? (((unsigned)crc << 1) ^ 0x1021u) : ((unsigned)crc << 1));
    }
    return (uint16_t)(crc ^ 0xFFFFu);
}

/* src/dab/fic/fic_decoder.cpp:53-116.  bits: nb_encoded_bits soft bits of one FIB group; out: nb_encoded_bits/24 descrambled
 * bytes; valid[nb_fibs]: CRC16 match per FIB.  Returns the Viterbi path error, or UINT64_MAX when the group size is not the
 * Mode I size the reference accepts (:70-75; nothing is decoded then). */
uint64_t orc_fic_decode_group(orc_viterbi* v, const int8_t* bits, size_t nb_encoded_bits, size_t nb_fibs, uint8_t* out, uint8_t* valid) {
    const size_t nb_decoded_bits = nb_encoded_bits / 3, nb_decoded_bytes = nb_encoded_bits / 24;
    for (size_t i = 0; i < nb_fibs; i++) valid[i] = 0;
    if (nb_decoded_bits != (128u * 21u + 128u * 3u + 24u) / 4u - 6u) return UINT64_MAX;
    if (orc_vit_get_traceback_length(v) != nb_decoded_bits) orc_vit_set_traceback_length(v, nb_decoded_bits);
    orc_vit_reset(v, 0);
    size_t used = 0;
    used += orc_vit_update(v, bits + used, nb_encoded_bits - used, orc_puncture_code(16), 8, 128 * 21);
    used += orc_vit_update(v, bits + used, nb_encoded_bits - used, orc_puncture_code(15), 8, 128 * 3);
    used += orc_vit_update(v, bits + used, nb_encoded_bits - used, PI_TAIL, 6, 24);
    const uint64_t err = orc_vit_chainback(v, out, nb_decoded_bytes, 0);
    uint8_t prbs[1024];
    orc_scrambler_bytes(0xFFFF, prbs, nb_decoded_bytes);
    for (size_t i = 0; i < nb_decoded_bytes; i++) out[i] ^= prbs[i];
    const size_t fib_bytes = nb_decoded_bytes / nb_fibs;
    for (size_t i = 0; i < nb_fibs; i++) {
        const uint8_t* fib = out + i * fib_bytes;
        const uint16_t rx = (uint16_t)((fib[fib_bytes - 2] << 8) | fib[fib_bytes - 1]);
        valid[i] = (uint8_t)(rx == orc_crc16_fib(fib, fib_bytes - 2));
    }
    return err;
}

/* src/dab/msc/cif_deinterleaver.cpp:9-70 */
static const int CIF_OFFSETS[16] = { 0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15 };
struct orc_deint { int8_t* ring; size_t nb_bits; int curr; int stored; };
orc_deint* orc_deint_create(size_t nb_bits) {
    orc_deint* d = (orc_deint*)calloc(1, sizeof(orc_deint));
    d->ring = (int8_t*)calloc(nb_bits * 16u, 1);
    d->nb_bits = nb_bits;
    return d;
}
void orc_deint_destroy(orc_deint* d) { if (d) { free(d->ring); free(d); } }
int orc_deint_push(orc_deint* d, const int8_t* bits, int8_t* out) {
    memcpy(d->ring + d->nb_bits * (size_t)d->curr, bits, d->nb_bits);    /* Consume :21-35 */
    d->curr = (d->curr + 1) % 16;
    if (d->stored < 16) d->stored++;
    if (d->stored < 16) return 0;                                        /* Deinterleave :37-70 */
    for (size_t i = 0; i < d->nb_bits; i++) {
        const int age = 15 - CIF_OFFSETS[i % 16];                        /* BUFFER_LOOKUP index: 0 = newest */
        const int row = ((d->curr - 1) - age + 32) % 16;
        out[i] = d->ring[d->nb_bits * (size_t)row + i];
    }
    return 1;
}

/* src/dab/constants/subchannel_protection_tables.h:21-86 (UEP rows: Lx[4], PIx[4]; the two 128 kbit/s rows are kept in the
 * reference's order), :121-140 (EEP A/B as Lx = m*n + b), :145-154 (2-A special when the sub-channel has 8 CU) */
static const uint8_t UEP_TABLE[64][8] = {
    {3,4,17,0, 5,3,2,0}, {3,3,18,0, 11,6,5,0}, {3,4,14,3, 15,9,6,8}, {3,4,14,3, 22,13,8,13}, {3,5,13,3, 24,17,12,17},
    {4,3,26,3, 5,4,2,3}, {3,4,26,3, 9,6,4,6}, {3,4,26,3, 15,10,6,9}, {3,4,26,3, 24,14,8,15}, {3,5,25,3, 24,18,13,18},
    {6,10,23,3, 5,4,2,3}, {6,10,23,3, 9,6,4,5}, {6,12,21,3, 16,7,6,9}, {6,10,23,3, 23,13,8,13},
    {6,9,31,2, 5,3,2,3}, {6,9,33,0, 11,6,5,0}, {6,12,27,3, 16,8,6,9}, {6,10,29,3, 23,13,8,13}, {6,11,28,3, 24,18,12,18},
    {6,10,41,3, 6,3,2,3}, {6,10,41,3, 11,6,5,6}, {6,11,40,3, 16,8,6,7}, {6,10,41,3, 23,13,8,13}, {6,10,41,3, 24,17,12,18},
    {7,9,53,3, 5,4,2,4}, {7,10,52,3, 9,6,4,6}, {6,12,51,3, 16,9,6,10}, {6,10,53,3, 22,12,9,12}, {6,13,50,3, 24,18,13,19},
    {14,17,50,3, 5,4,2,5}, {11,21,49,3, 9,6,4,8}, {11,23,47,3, 16,8,6,9}, {11,21,49,3, 23,12,9,14},
    {12,19,62,3, 5,3,2,4}, {11,21,61,3, 11,6,5,7}, {11,22,60,3, 16,9,6,10}, {11,21,61,3, 22,12,9,14}, {11,20,62,3, 24,17,13,19},
    {11,19,87,3, 5,4,2,4}, {11,23,83,3, 11,6,5,9}, {11,24,82,3, 16,8,6,11}, {11,21,85,3, 22,11,9,13}, {11,22,84,3, 24,18,12,19},
    {11,20,110,3, 6,4,2,5}, {11,22,108,3, 10,6,4,9}, {11,24,106,3, 16,10,6,11}, {11,20,110,3, 22,13,9,13}, {11,21,109,3, 24,20,13,24},
    {12,22,131,3, 8,6,2,6}, {12,26,127,3, 12,8,4,11}, {11,20,134,3, 16,10,7,9}, {11,22,132,3, 24,16,10,15}, {11,24,130,3, 24,20,12,20},
    {11,24,154,3, 6,5,2,5}, {11,24,154,3, 12,9,5,10}, {11,27,151,3, 16,10,7,10}, {11,22,156,3, 24,14,10,13}, {11,26,152,3, 24,19,14,18},
    {11,26,200,3, 8,5,2,6}, {11,25,201,3, 13,9,5,10}, {11,26,200,3, 24,17,9,17},
    {11,27,247,3, 8,6,2,7}, {11,24,250,3, 16,9,7,10}, {12,28,245,3, 24,20,14,23},
};
static const uint16_t UEP_SIZE[64] = {
    16,21,24,29,35, 24,29,35,42,52, 29,35,42,52, 32,42,48,58,70, 40,52,58,70,84, 48,58,70,84,104, 58,70,84,104,
    84,64,96,116,140, 80,104,116,140,168, 96,116,140,168,208, 116,140,168,208,232, 128,168,192,232,280, 160,208,280, 192,280,416,
};
int orc_uep_subchannel_size(int index) { return (index >= 0 && index < 64) ? (int)UEP_SIZE[index] : -1; }
typedef struct { int cu_multiple; int m[2], b[2]; int pi[2]; } eep_desc;
static const eep_desc EEP_A[4] = { {12, {6, 0}, {-3, 3}, {24, 23}}, {8, {2, 4}, {-3, 3}, {14, 13}}, {6, {6, 0}, {-3, 3}, {8, 7}}, {4, {4, 2}, {-3, 3}, {3, 2}} };
static const eep_desc EEP_2A_SPECIAL = {8, {0, 0}, {5, 1}, {13, 12}};
static const eep_desc EEP_B[4] = { {27, {24, 0}, {-3, 3}, {10, 9}}, {21, {24, 0}, {-3, 3}, {6, 5}}, {18, {24, 0}, {-3, 3}, {4, 3}}, {15, {24, 0}, {-3, 3}, {2, 1}} };

/* The update() sequence of MSC_Decoder::DecodeEEP / DecodeUEP (msc_decoder.cpp:88-99, 136-146) for one sub-channel, as
 * (puncture index, requested output symbols) pairs; the PI_X tail (24 symbols) is pi = 0.  Returns the number of entries. */
int orc_msc_segments(const orc_subchannel* sc, int pi_out[5], uint32_t n_out[5]) {
    int k = 0;
    if (!sc->is_uep) {
        const eep_desc* d = sc->eep_type_b ? &EEP_B[sc->eep_prot_level & 3] : (sc->length == 8 ? &EEP_2A_SPECIAL : &EEP_A[sc->eep_prot_level & 3]);
        const int n = sc->length / d->cu_multiple;
        for (int i = 0; i < 2; i++) { pi_out[k] = d->pi[i]; n_out[k] = (uint32_t)(128 * (d->m[i] * n + d->b[i])); k++; }
    } else {
        const uint8_t* row = UEP_TABLE[sc->uep_prot_index & 63];
        for (int i = 0; i < 4; i++) { pi_out[k] = row[4 + i]; n_out[k] = 128u * row[i]; k++; }
    }
    pi_out[k] = 0; n_out[k] = 24; k++;
    return k;
}

/* src/dab/msc/msc_decoder.cpp:27-170 */
struct orc_msc { orc_subchannel sc; size_t nb_bits; orc_deint* deint; orc_viterbi* vit; int8_t* enc; uint8_t* dec; uint8_t* prbs; };
orc_msc* orc_msc_create(const orc_subchannel* sc) {
    orc_msc* m = (orc_msc*)calloc(1, sizeof(orc_msc));
    m->sc = *sc;
    m->nb_bits = (size_t)sc->length * 64u;
    m->deint = orc_deint_create(m->nb_bits);
    m->vit = orc_vit_create();
    orc_vit_set_traceback_length(m->vit, m->nb_bits);      /* :38 */
    m->enc = (int8_t*)calloc(m->nb_bits + 1, 1);
    m->dec = (uint8_t*)calloc(m->nb_bits / 8u + 1, 1);
    m->prbs = (uint8_t*)calloc(m->nb_bits / 8u + 1, 1);
    orc_scrambler_bytes(0xFFFF, m->prbs, m->nb_bits / 8u);
    return m;
}
void orc_msc_destroy(orc_msc* m) { if (m) { orc_deint_destroy(m->deint); orc_vit_destroy(m->vit); free(m->enc); free(m->dec); free(m->prbs); free(m); } }
/* DecodeCIF: returns the number of bytes written to out (0 while the de-interleaver fills, -1 when the sub-channel overflows
 * the CIF, :49-54); *path_error receives the Viterbi error when bytes were produced */
int64_t orc_msc_decode_cif(orc_msc* m, const int8_t* cif_bits, size_t n_bits, uint8_t* out, uint64_t* path_error) {
    const size_t start_bit = (size_t)m->sc.start_address * 64u;
    if (start_bit + m->nb_bits > n_bits) return -1;
    if (!orc_deint_push(m->deint, cif_bits + start_bit, m->enc)) return 0;
    int pi[5]; uint32_t n_out[5];
    const int n_seg = orc_msc_segments(&m->sc, pi, n_out);
    orc_vit_reset(m->vit, 0);
    size_t used = 0;
    for (int i = 0; i < n_seg; i++) {
        const uint8_t* code = pi[i] ? orc_puncture_code(pi[i]) : PI_TAIL;
        used += orc_vit_update(m->vit, m->enc + used, m->nb_bits - used, code, pi[i] ? 8 : 6, n_out[i]);
    }
    const size_t decoded_bits = orc_vit_get_current_decoded_bit(m->vit) - 6u;   /* :105-108 (24 tail symbols / code rate 4) */
    const size_t nbytes = decoded_bits / 8u;
    const uint64_t err = orc_vit_chainback(m->vit, m->dec, nbytes, 0);
    if (path_error) *path_error = err;
    for (size_t i = 0; i < nbytes; i++) out[i] = (uint8_t)(m->dec[i] ^ m->prbs[i]);   /* :112-117 */
    return (int64_t)nbytes;
}

/* ------------------------------------------------------------------------------------------------ CPU baselines */

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

typedef struct { int mode; const orc_c32* iq; size_t n, block; int repeats; uint64_t frames; } ofdm_job;
static void* ofdm_worker(void* arg) {
    ofdm_job* j = (ofdm_job*)arg;
    orc_ofdm* d = orc_ofdm_create(j->mode);
    for (int r = 0; r < j->repeats; r++)
        for (size_t off = 0; off < j->n; off += j->block) {
            const size_t len = (j->n - off < j->block) ? (j->n - off) : j->block;
            orc_ofdm_process(d, j->iq + off, len);
            /* drop collected frames so memory stays flat */
            for (size_t i = 0; i < d->n_frames; i++) free(d->frames[i]);
            d->n_frames = 0;
        }
    j->frames = (uint64_t)d->total_frames_read;
    orc_ofdm_destroy(d);
    return NULL;
}

double orc_ofdm_bench(int mode, int n_threads, const orc_c32* iq, size_t n, size_t block, int repeats, uint64_t* frames_out) {
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    ofdm_job* jobs = (ofdm_job*)calloc((size_t)n_threads, sizeof(ofdm_job));
    const double t0 = now_s();
    for (int i = 0; i < n_threads; i++) {
        jobs[i].mode = mode; jobs[i].iq = iq; jobs[i].n = n; jobs[i].block = block; jobs[i].repeats = repeats;
        pthread_create(&th[i], NULL, ofdm_worker, &jobs[i]);
    }
    uint64_t frames = 0;
    for (int i = 0; i < n_threads; i++) { pthread_join(th[i], NULL); frames += jobs[i].frames; }
    const double t1 = now_s();
    if (frames_out) *frames_out = frames;
    free(th); free(jobs);
    return t1 - t0;
}

typedef struct {
    int tid, n_threads; const int8_t* soft; size_t soft_per_job, n_jobs; const uint8_t* seg_codes; const uint32_t* seg_code_len;
    const uint32_t* seg_n_out; uint32_t n_seg; size_t traceback_bits; uint8_t* out; size_t out_bytes_per_job;
} vit_job;
static void* vit_worker(void* arg) {
    vit_job* j = (vit_job*)arg;
    orc_viterbi* v = orc_vit_create();
    orc_vit_set_traceback_length(v, j->traceback_bits);
    for (size_t k = (size_t)j->tid; k < j->n_jobs; k += (size_t)j->n_threads)
        orc_vit_decode_job(v, j->soft + k * j->soft_per_job, j->soft_per_job, j->seg_codes, j->seg_code_len, j->seg_n_out, j->n_seg,
                           j->out + k * j->out_bytes_per_job, j->out_bytes_per_job, NULL);
    orc_vit_destroy(v);
    return NULL;
}

double orc_vit_bench(int n_threads, const int8_t* soft, size_t soft_per_job, size_t n_jobs, const uint8_t* seg_codes,
                     const uint32_t* seg_code_len, const uint32_t* seg_n_out, uint32_t n_seg, size_t traceback_bits, uint8_t* out,
                     size_t out_bytes_per_job) {
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    vit_job* jobs = (vit_job*)calloc((size_t)n_threads, sizeof(vit_job));
    const double t0 = now_s();
    for (int i = 0; i < n_threads; i++) {
        vit_job j = { i, n_threads, soft, soft_per_job, n_jobs, seg_codes, seg_code_len, seg_n_out, n_seg, traceback_bits, out, out_bytes_per_job };
        jobs[i] = j;
        pthread_create(&th[i], NULL, vit_worker, &jobs[i]);
    }
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    const double t1 = now_s();
    free(th); free(jobs);
    return t1 - t0;
}
