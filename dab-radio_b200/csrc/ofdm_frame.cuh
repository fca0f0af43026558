// Frame demodulation kernel: the per-frame pipeline of OFDM_Demod::PipelineThread (reference ofdm_demodulator.cpp:650-766)
// in ONE pass over HBM -- IQ is read once, int8 soft bits are written once:
//   PLL (apply_pll.cpp:82-116)  ->  cyclic-prefix phase error (:768-777)  ->  FFT (:891-894, FFTW3 forward c2c)
//   ->  DQPSK against the previous symbol's spectrum (:842-865)  ->  frequency de-interleave + L-inf normalise + int8
//   quantise (:867-889, :57-72).
//
// Work item = (frame, chunk of consecutive symbols).  N/16 threads run one item: each thread keeps 16 FFT points and the 16
// matching bins of the previous symbol in registers, so the differential step between consecutive symbols of a chunk costs
// no HBM or shared-memory traffic.  A 128-thread CTA hosts 128/(N/16) items (1 for the 2048-point mode I FFT, 8 for mode III).
// Per symbol the CTA meets two barriers (the two FFT exchanges); soft bits are staged in shared memory in de-interleaved
// order and leave as 16-byte coalesced stores while the next symbol is in flight.
#pragma once
#include "ofdm_device.cuh"

namespace dabb200 {

struct FrameDesc {
    const void* src;       // float2 (or uchar2 for raw-u8 ingest) sample base of the stream
    uint64_t mask;         // index mask: ring size - 1, or ~0 for a linear buffer
    int64_t start;         // sample index of the PRS cyclic-prefix start
    float freq;            // net PLL frequency (coarse + fine), cycles / sample
    int32_t valid;         // 0: nothing to do for this frame slot
    int8_t* bits;          // out: (S-1) * 2 * ncarr soft bits
    float* phase_err;      // out: S cyclic-prefix phase errors (radians), one per symbol
    float2* fft_tap;       // optional GUI tap: S * NFFT spectra (natural bin order), else nullptr
    float2* vec_tap;       // optional GUI tap: (S-1) * ncarr DQPSK vectors (carrier order), else nullptr
    uint64_t limit;        // samples addressable from src without wrapping (ring size, or the length of a linear buffer):
                           // bounds the 16-byte-granular bulk copies of the v3 kernel
};

struct FrameGeom {
    int n_symbols;         // S: PRS + data symbols (the NULL symbol is not demodulated: it feeds no soft bit)
    int symbol_period;
    int cyclic_prefix;
    int n_carriers;
    int syms_per_chunk;    // DQPSK outputs per work item
    int n_chunks;          // ceil((S-1) / syms_per_chunk)
    const int16_t* bin_to_pos;      // [NFFT]: de-interleaved soft-bit position of FFT bin k, -1 for DC / guard bins
    const int16_t* bin_to_carrier;  // [NFFT]: carrier index (DQPSK vector order) of bin k, -1 if unused
    const float2* twiddles;         // [TW1_SIZE + TW2_SIZE] precomputed FFT twiddles
};

template <bool RAW_U8>
__device__ __forceinline__ float2 load_sample(const void* src, uint64_t index) {
    if (RAW_U8) {
        // examples/app_helpers/app_iq_readers.h:17-69: (u8 - 127.5) * (1 / 127.5)
        const uchar2 q = __ldg(reinterpret_cast<const uchar2*>(src) + index);
        const float scale = 1.0f / 127.5f;
        return make_float2((float(q.x) - 127.5f) * scale, (float(q.y) - 127.5f) * scale);
    }
    float2 v;
    const float2* p = reinterpret_cast<const float2*>(src) + index;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

constexpr int FRAME_CTA_THREADS = 128;

template <int NFFT>
struct FrameSmem {
    using G = FftGeom<NFFT>;
    static constexpr int GROUPS = FRAME_CTA_THREADS / G::T;
    static __host__ __device__ size_t stage_bytes(int n_carriers) { return (size_t(2 * n_carriers) + 15) & ~size_t(15); }
    static __host__ __device__ size_t group_bytes(int n_carriers) {
        return size_t(G::E1_SIZE + G::E2_SIZE) * sizeof(float2) + stage_bytes(n_carriers) + 8 * sizeof(float2);
    }
    static __host__ __device__ size_t total_bytes(int n_carriers) {
        return size_t(G::TW1_SIZE + G::TW2_SIZE) * sizeof(float2) + size_t(GROUPS) * group_bytes(n_carriers);
    }
};

template <int NFFT, bool RAW_U8>
__global__ void __launch_bounds__(FRAME_CTA_THREADS, 4)
ofdm_frame_kernel(FrameGeom geo, const FrameDesc* __restrict__ descs, int n_frames) {
    using G = FftGeom<NFFT>;
    using SM = FrameSmem<NFFT>;
    constexpr int T = G::T;
    constexpr int GROUPS = SM::GROUPS;
    constexpr int WARPS_PER_GROUP = (T + 31) / 32;
    constexpr int RED_WIDTH = (T < 32) ? T : 32;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw1 = reinterpret_cast<float2*>(smem_raw);
    float2* tw2 = tw1 + G::TW1_SIZE;
    const int group = threadIdx.x / T, t = threadIdx.x % T;
    unsigned char* gbase = reinterpret_cast<unsigned char*>(tw2 + G::TW2_SIZE) + size_t(group) * SM::group_bytes(geo.n_carriers);
    float2* e1 = reinterpret_cast<float2*>(gbase);
    float2* e2 = e1 + G::E1_SIZE;
    int8_t* stage = reinterpret_cast<int8_t*>(e2 + G::E2_SIZE);
    float2* red = reinterpret_cast<float2*>(stage + SM::stage_bytes(geo.n_carriers));

    const int n_items = n_frames * geo.n_chunks;
    const int item = blockIdx.x * GROUPS + group;
    const int frame = (item < n_items) ? item / geo.n_chunks : 0;
    const int chunk = (item < n_items) ? item % geo.n_chunks : 0;
    const FrameDesc desc = descs[frame];
    const bool active = (item < n_items) && desc.valid != 0;
    // a CTA without any frame to demodulate (streams that completed no frame in this pass) leaves at once
    if (GROUPS == 1) {
        if (!active) return;
    } else {
        if (!__syncthreads_or(active ? 1 : 0)) return;
    }
    fft_load_twiddles<NFFT>(tw1, geo.twiddles, threadIdx.x, FRAME_CTA_THREADS);

    const int S = geo.n_symbols, sp = geo.symbol_period, cp = geo.cyclic_prefix, ncarr = geo.n_carriers;
    const int s_first = chunk * geo.syms_per_chunk;                  // first symbol whose DQPSK output this item owns
    const int s_out_end = min(s_first + geo.syms_per_chunk, S - 1);  // one past the last owned output symbol

    // soft-bit positions of my 16 output bins (fixed for the whole kernel), two int16 per register
    uint32_t pos_pack[8];
#pragma unroll
    for (int r = 0; r < 16; r += 2) {
        const uint32_t lo = uint16_t(geo.bin_to_pos[fft_out_bin<NFFT>(t, r)]);
        const uint32_t hi = uint16_t(geo.bin_to_pos[fft_out_bin<NFFT>(t, r + 1)]);
        pos_pack[r / 2] = lo | (hi << 16);
    }

    const int tail0 = NFFT - cp;  // FFT-window index from which samples pair with the cyclic prefix (needs cp <= NFFT / 4)
    float2 prev[16];
#pragma unroll
    for (int r = 0; r < 16; r++) prev[r] = make_float2(0.0f, 0.0f);
    int staged = -1;  // output symbol sitting in `stage` (group-uniform), not yet written to HBM

    __syncthreads();  // twiddle tables ready

    for (int si = 0; si <= geo.syms_per_chunk; si++) {
        const int s = s_first + si;
        const bool sym_active = active && (s <= s_out_end);  // s_out_end <= S - 1
        float2 v[16];
        float2 corr = make_float2(0.0f, 0.0f);
        if (sym_active) {
            const uint64_t sym0 = uint64_t(desc.start + int64_t(s) * sp);
            const PllSymbol pll = pll_symbol(desc.freq, s * sp, sp);
            // FFT window (cyclic prefix removed, ofdm_demodulator.cpp:705) and the prefix itself: coalesced 8-byte loads
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = load_sample<RAW_U8>(desc.src, (sym0 + uint64_t(cp + t + T * j)) & desc.mask);
            float2 head[4];
#pragma unroll
            for (int j = 12; j < 16; j++) {
                const int w = t + T * j;
                head[j - 12] = (w >= tail0) ? load_sample<RAW_U8>(desc.src, (sym0 + uint64_t(w - tail0)) & desc.mask) : make_float2(0.0f, 0.0f);
            }
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = pll_rotate(pll, v[j], cp + t + T * j);
            // cyclic-prefix correlation sum x[nfft + n] * conj(x[n]), n < cp, on PLL'd samples (complex_conj_mul_sum.cpp:65-100)
#pragma unroll
            for (int j = 12; j < 16; j++) {
                const int w = t + T * j;
                if (w >= tail0) {
                    const float2 h = pll_rotate(pll, head[j - 12], w - tail0);
                    const float2 pr = cmul_conj(v[j], h);
                    corr.x += pr.x;
                    corr.y += pr.y;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = make_float2(0.0f, 0.0f);
        }
        corr = group_reduce_sum<RED_WIDTH>(corr);
        if (WARPS_PER_GROUP > 1 && (t & 31) == 0) red[t >> 5] = corr;

        fft_pass1<NFFT>(v, t, e1, tw1);
        __syncthreads();  // ---- barrier A: exchange 1 complete, correlation partials visible, previous staging complete

        if (sym_active && t == 0 && desc.phase_err != nullptr && (s < s_out_end || s == S - 1)) {
            float2 tot = corr;
            if (WARPS_PER_GROUP > 1) {
                tot = red[0];
#pragma unroll
                for (int w = 1; w < WARPS_PER_GROUP; w++) { tot.x += red[w].x; tot.y += red[w].y; }
            }
            desc.phase_err[s] = atan2f(tot.y, tot.x);  // CalculateCyclicPhaseError, ofdm_demodulator.cpp:776
        }
        if (staged >= 0) {  // soft bits of the previous output symbol leave as 16-byte stores
            const uint4* src4 = reinterpret_cast<const uint4*>(stage);
            uint4* dst4 = reinterpret_cast<uint4*>(desc.bits + size_t(staged) * size_t(2 * ncarr));
            for (int i = t; i < (2 * ncarr) / 16; i += T) dst4[i] = src4[i];
            staged = -1;
        }

        fft_pass2<NFFT>(v, t, e1, e2, tw2);
        __syncthreads();  // ---- barrier B: exchange 2 complete, staging buffer free
        fft_pass3<NFFT>(v, t, e2);

        if (sym_active && desc.fft_tap != nullptr && (s < s_out_end || s == S - 1)) {
#pragma unroll
            for (int r = 0; r < 16; r++) desc.fft_tap[size_t(s) * NFFT + fft_out_bin<NFFT>(t, r)] = v[r];
        }

        // DQPSK X_{s-1} * conj(X_s) (ofdm_demodulator.cpp:736,861), de-interleave, quantise into the staging buffer
        if (si > 0 && sym_active) {
            const int s_out = s - 1;
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int pos = int(int16_t((pos_pack[r / 2] >> (16 * (r & 1))) & 0xFFFFu));
                if (pos >= 0) {
                    const float2 d = cmul_conj(prev[r], v[r]);
                    const float a = fmaxf(fabsf(d.x), fabsf(d.y));
                    // the reference divides by A exactly (max component -> +-127); rcp.rn(A) * 127.00003 reproduces that and is
                    // within 3e-5 of x / A * 127 elsewhere; A = 0 gives NaN -> 0 like the reference's cast
                    const float ra = __frcp_rn(a) * 127.00003f;
                    stage[pos] = int8_t(__float2int_rz(-d.x * ra));
                    stage[pos + ncarr] = int8_t(__float2int_rz(d.y * ra));
                    if (desc.vec_tap != nullptr) {
                        const int c = geo.bin_to_carrier[fft_out_bin<NFFT>(t, r)];
                        desc.vec_tap[size_t(s_out) * ncarr + c] = d;
                    }
                }
            }
            staged = s_out;
        }
#pragma unroll
        for (int r = 0; r < 16; r++) prev[r] = v[r];
    }
    __syncthreads();
    if (staged >= 0) {
        const uint4* src4 = reinterpret_cast<const uint4*>(stage);
        uint4* dst4 = reinterpret_cast<uint4*>(desc.bits + size_t(staged) * size_t(2 * ncarr));
        for (int i = t; i < (2 * ncarr) / 16; i += T) dst4[i] = src4[i];
    }
}

}  // namespace dabb200
