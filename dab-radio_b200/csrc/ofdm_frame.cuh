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

// How the stream buffers hold a sample (examples/app_helpers/app_iq_readers.h:17-88): complex float, or a raw
// integer pair straight from the front end, dequantised on the way into the registers as the reference's QuantisedIQ<T>::to_c32 +
// scale does:  value = (float(raw) - BIAS) * (1 / MAX_AMPLITUDE).  The kernels are compiled per sample SIZE (8 / 2 / 4 bytes); which
// integer type a 2- or 4-byte sample is (signedness, byte order) is a runtime parameter: a signed value is read as
// unsigned ^ sign bit with the bias raised by 2^(bits-1) (exact in fp32), a big-endian 16-bit pair costs one PRMT.
struct SampleFmt {
    uint32_t flip;   // xor mask on the unsigned component: 0 (unsigned), 0x80 (s8), 0x8000 (s16)
    uint32_t prmt;   // byte selector for the 4-byte formats: 0x3210 little endian, 0x2301 big endian
    float bias;      // u8 127.5, s8 128, u16 32767.5, s16 32768
    float scale;     // 1 / MAX_AMPLITUDE: u8 1/127.5, s8 1/127, u16 1/32767.5, s16 1/32767
};

// SB = bytes per complex sample: 8 (float2), 2 (u8 / s8 pair), 4 (u16 / s16 pair)
template <int SB>
__device__ __forceinline__ float2 decode_sample(uint32_t raw, const SampleFmt& f) {
    uint32_t i, q;
    if (SB == 2) {
        i = raw & 0xFFu;
        q = (raw >> 8) & 0xFFu;
    } else {
        raw = __byte_perm(raw, 0u, f.prmt);
        i = raw & 0xFFFFu;
        q = raw >> 16;
    }
    return make_float2((float(i ^ f.flip) - f.bias) * f.scale, (float(q ^ f.flip) - f.bias) * f.scale);
}

template <int SB>
__device__ __forceinline__ float2 load_sample_ptr(const void* p, const SampleFmt& f) {
    if (SB == 2) return decode_sample<2>(uint32_t(__ldg(reinterpret_cast<const unsigned short*>(p))), f);
    if (SB == 4) return decode_sample<4>(__ldg(reinterpret_cast<const unsigned int*>(p)), f);
    float2 v;   // read-only for the kernel's lifetime: not volatile, the compiler may batch independent loads
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

template <int SB>
__device__ __forceinline__ float2 load_sample(const void* src, uint64_t index, const SampleFmt& f) {
    return load_sample_ptr<SB>(reinterpret_cast<const unsigned char*>(src) + index * uint64_t(SB), f);
}

// Two consecutive samples with one load; p is aligned to 2 * SB bytes.
template <int SB>
__device__ __forceinline__ void load_sample_pair(const void* p, const SampleFmt& f, float2& a, float2& b) {
    if (SB == 2) {
        const uint32_t raw = __ldg(reinterpret_cast<const unsigned int*>(p));
        a = decode_sample<2>(raw & 0xFFFFu, f);
        b = decode_sample<2>(raw >> 16, f);
    } else if (SB == 4) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
        a = decode_sample<4>(raw.x, f);
        b = decode_sample<4>(raw.y, f);
    } else {
        asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(b.x), "=f"(b.y) : "l"(p));
    }
}

// One work item of the frame kernels: symbols [s_begin, s_end) of one transmission frame of one stream.  The item writes the
// cyclic-prefix phase error of each of its symbols and the soft-bit rows s - 1 (DQPSK of symbols s - 1 and s) for s >= 1; for
// s_begin > 0 it transforms symbol s_begin - 1 once more as the differential reference.  A frame is covered by any tiling of
// [0, S) into items (ofdm_control.cuh splits a completed frame into n_chunks of them for load balance).
struct FrameDesc {
    const void* src;       // sample base of the stream (SampleFmt layout)
    uint64_t mask;         // index mask: ring size - 1, or ~0 for a linear buffer
    uint64_t limit;        // samples addressable from src without wrapping (ring size, or the length of a linear buffer):
                           // bounds the 16-byte-granular bulk copies of the v3 kernel
    int64_t start;         // sample index of the PRS cyclic-prefix start (symbol 0 of the frame)
    float freq;            // net PLL frequency (coarse + fine), cycles / sample
    int32_t valid;         // 0: nothing to do for this item
    int32_t s_begin, s_end;
    int8_t* bits;          // out: (S-1) * 2 * ncarr soft bits of the frame
    float* phase_err;      // out: S cyclic-prefix phase errors (radians), one per symbol
    float2* fft_tap;       // optional GUI tap: S * NFFT spectra (natural bin order), else nullptr
    float2* vec_tap;       // optional GUI tap: (S-1) * ncarr DQPSK vectors (carrier order), else nullptr
};

struct FrameGeom {
    int n_symbols;         // S: PRS + data symbols (the NULL symbol is not demodulated: it feeds no soft bit)
    int symbol_period;
    int cyclic_prefix;
    int n_carriers;
    SampleFmt fmt;
    const int16_t* bin_to_pos;      // [NFFT]: de-interleaved soft-bit position of FFT bin k, -1 for DC / guard bins
    const int16_t* bin_to_carrier;  // [NFFT]: carrier index (DQPSK vector order) of bin k, -1 if unused
    const int16_t* bin_to_slot;     // [NFFT]: v3 kernel, staging slot of bin k's soft-bit pair (stage_layout.h), -1 if unused
    const uint16_t* chunk_src;      // [n_carriers / 8]: v3 kernel, chunk slot << 2 | word rotation of output chunk i
    const float2* twiddles;         // [TW1_SIZE + TW2_SIZE] precomputed FFT twiddles
};

// The DAB transmission-mode geometry (the four modes of dab_ofdm_params_ref.cpp:10-57 share cp = 63 N / 256, symbol = N + cp,
// carriers = 3 N / 4): what the v3 kernel is specialised for.
template <int NFFT>
struct DabGeom {
    static constexpr int CP = 63 * NFFT / 256;
    static constexpr int SP = NFFT + CP;
    static constexpr int NCARR = 3 * NFFT / 4;
    static constexpr int HALF = NCARR / 2;
    static constexpr int TAIL0 = NFFT - CP;              // FFT-window index pairing with cyclic-prefix sample 0
    static __host__ __device__ constexpr bool matches(int sp, int cp, int ncarr) { return sp == SP && cp == CP && ncarr == NCARR; }
    static __host__ __device__ constexpr bool bin_used(int k) { return (k >= 1 && k <= HALF) || (k >= NFFT - HALF && k < NFFT); }
};

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

// Generic-geometry frame kernel: any OFDM_Params the reference API accepts (cp <= nfft / 4, carriers a multiple of 8).  The four
// DAB transmission modes run ofdm_frame_v3_kernel instead; DAB_B200_GENERIC_FRAME_KERNEL=1 routes them here too (cross-check in
// tests/test_ofdm_gpu.py).
template <int NFFT, int SB>
__global__ void __launch_bounds__(FRAME_CTA_THREADS, 4)
ofdm_frame_kernel(FrameGeom geo, const FrameDesc* __restrict__ descs, int n_items) {
    using G = FftGeom<NFFT>;
    using SM = FrameSmem<NFFT>;
    constexpr int T = G::T;
    constexpr int GROUPS = SM::GROUPS;
    constexpr int WARPS_PER_GROUP = (T + 31) / 32;
    constexpr int RED_WIDTH = (T < 32) ? T : 32;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int n_loop_s;
    float2* tw1 = reinterpret_cast<float2*>(smem_raw);
    float2* tw2 = tw1 + G::TW1_SIZE;
    const int group = threadIdx.x / T, t = threadIdx.x % T;
    unsigned char* gbase = reinterpret_cast<unsigned char*>(tw2 + G::TW2_SIZE) + size_t(group) * SM::group_bytes(geo.n_carriers);
    float2* e1 = reinterpret_cast<float2*>(gbase);
    float2* e2 = e1 + G::E1_SIZE;
    int8_t* stage = reinterpret_cast<int8_t*>(e2 + G::E2_SIZE);
    float2* red = reinterpret_cast<float2*>(stage + SM::stage_bytes(geo.n_carriers));

    const int item = blockIdx.x * GROUPS + group;
    const FrameDesc desc = descs[(item < n_items) ? item : 0];
    const bool active = (item < n_items) && desc.valid != 0 && desc.s_end > desc.s_begin;
    // a CTA without any symbol to demodulate leaves at once
    if (threadIdx.x == 0) n_loop_s = 0;
    if (!__syncthreads_or(active ? 1 : 0)) return;
    const int s_load0 = max(desc.s_begin - 1, 0);   // first symbol transformed: the differential reference (or the PRS itself)
    if (active && t == 0) atomicMax(&n_loop_s, desc.s_end - s_load0);
    fft_load_twiddles<NFFT>(tw1, geo.twiddles, threadIdx.x, FRAME_CTA_THREADS);

    const int sp = geo.symbol_period, cp = geo.cyclic_prefix, ncarr = geo.n_carriers;

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

    __syncthreads();  // twiddle tables and the CTA's loop count ready
    const int n_loop = n_loop_s;

    for (int si = 0; si < n_loop; si++) {
        const int s = s_load0 + si;
        const bool sym_active = active && (s < desc.s_end);
        const bool own = sym_active && (s >= desc.s_begin);
        float2 v[16];
        float2 corr = make_float2(0.0f, 0.0f);
        if (sym_active) {
            const uint64_t sym0 = uint64_t(desc.start + int64_t(s) * sp);
            const PllSymbol pll = pll_symbol(desc.freq, s * sp, sp);
            // FFT window (cyclic prefix removed, ofdm_demodulator.cpp:705) and the prefix itself: coalesced loads
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = load_sample<SB>(desc.src, (sym0 + uint64_t(cp + t + T * j)) & desc.mask, geo.fmt);
            float2 head[4];
#pragma unroll
            for (int j = 12; j < 16; j++) {
                const int w = t + T * j;
                head[j - 12] = (w >= tail0) ? load_sample<SB>(desc.src, (sym0 + uint64_t(w - tail0)) & desc.mask, geo.fmt) : make_float2(0.0f, 0.0f);
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

        if (own && t == 0 && desc.phase_err != nullptr) {
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

        if (own && desc.fft_tap != nullptr) {
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
