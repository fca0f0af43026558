// Frame demodulation kernel specialised for the DAB transmission-mode geometry (the four modes of
// dab_ofdm_params_ref.cpp:10-57 share cp = 63 N / 256, symbol = N + cp, carriers = 3 N / 4).  Same single pass over HBM and
// same arithmetic as ofdm_frame_kernel (ofdm_frame.cuh, kept as the generic-geometry fallback); what the compile-time geometry
// buys, measured with ncu on the generic kernel (profiles/r01_frame_kernel_v1.md):
//   * no per-load 64-bit ring arithmetic: one symbol base pointer, every sample load is [base + immediate]
//   * PLL phase arguments from per-thread float constants (no int->float conversions, no tail branch in modes I / IV)
//   * the quarter of the FFT bins that carry no data (guard band) is dropped at compile time: no DQPSK, no stores, and
//     only 12 instead of 16 "previous symbol" bins live in registers
//   * soft bits are staged as 16-bit (re, im) pairs: half the shared-memory stores, de-interleaved with PRMT on the way out
//   * the next symbol's samples are requested before the current symbol's FFT, so HBM latency hides behind the FFT
#pragma once
#include "ofdm_frame.cuh"

namespace dabb200 {

template <int NFFT>
struct DabGeom {
    static constexpr int CP = 63 * NFFT / 256;
    static constexpr int SP = NFFT + CP;
    static constexpr int NCARR = 3 * NFFT / 4;
    static constexpr int HALF = NCARR / 2;
    static constexpr bool HAS_TAIL = (SP % 4) != 0;     // apply_pll's scalar tail (modes II and III)
    static constexpr int TAIL0 = NFFT - CP;              // FFT-window index pairing with cyclic-prefix sample 0
    static __host__ __device__ constexpr bool matches(int sp, int cp, int ncarr) { return sp == SP && cp == CP && ncarr == NCARR; }
    static __host__ __device__ constexpr bool bin_used(int k) { return (k >= 1 && k <= HALF) || (k >= NFFT - HALF && k < NFFT); }
    // 0: no thread's bin in register slot r carries data, 1: every thread's does, 2: depends on the thread
    static __host__ __device__ constexpr int slot_kind(int r) {
        constexpr int R3 = NFFT / 256, T = NFFT / 16;
        int lo = 0;
        if (R3 == 1) {
            lo = 16 * r;  // bin = t + 16 r
        } else {
            const int m = r / R3, k3 = r % R3;
            lo = T * m + 256 * k3;  // bin = (t + T m) + 256 k3
        }
        const int span = (R3 == 1) ? 16 : T;
        int used = 0;
        for (int t = 0; t < span; t++) used += bin_used(lo + t) ? 1 : 0;
        return used == 0 ? 0 : (used == span ? 1 : 2);
    }
    static __host__ __device__ constexpr uint32_t used_mask() {
        uint32_t m = 0;
        for (int r = 0; r < 16; r++) m |= (slot_kind(r) != 0 ? 1u : 0u) << r;
        return m;
    }
    static constexpr uint32_t USED_MASK = used_mask();   // bit r: register slot r holds a data carrier for at least one thread
};

template <int NFFT>
struct FrameDabSmem {
    using G = FftGeom<NFFT>;
    using D = DabGeom<NFFT>;
    static constexpr int GROUPS = FRAME_CTA_THREADS / G::T;
    static constexpr size_t STAGE_BYTES = ((size_t(D::NCARR) + 8) * 2 + 15) & ~size_t(15);  // u16 per carrier + a dummy slot
    static constexpr size_t GROUP_BYTES = size_t(G::E1_SIZE + G::E2_SIZE) * sizeof(float2) + STAGE_BYTES + 8 * sizeof(float2);
    static constexpr size_t TOTAL_BYTES = size_t(G::TW1_SIZE + G::TW2_SIZE) * sizeof(float2) + size_t(GROUPS) * GROUP_BYTES;
};

template <bool RAW_U8>
__device__ __forceinline__ float2 load_sample_ptr(const void* p) {
    if (RAW_U8) {
        const uchar2 q = __ldg(reinterpret_cast<const uchar2*>(p));
        const float scale = 1.0f / 127.5f;
        return make_float2((float(q.x) - 127.5f) * scale, (float(q.y) - 127.5f) * scale);
    }
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

template <int NFFT, bool RAW_U8, int MIN_BLOCKS>
__global__ void __launch_bounds__(FRAME_CTA_THREADS, MIN_BLOCKS)
ofdm_frame_dab_kernel(FrameGeom geo, const FrameDesc* __restrict__ descs, int n_frames) {
    using G = FftGeom<NFFT>;
    using D = DabGeom<NFFT>;
    using SM = FrameDabSmem<NFFT>;
    constexpr int T = G::T;
    constexpr int GROUPS = SM::GROUPS;
    constexpr int WARPS_PER_GROUP = (T + 31) / 32;
    constexpr int RED_WIDTH = (T < 32) ? T : 32;
    constexpr int CP = D::CP, SP = D::SP, NCARR = D::NCARR, TAIL0 = D::TAIL0;
    constexpr int SAMPLE_BYTES = RAW_U8 ? 2 : 8;
    constexpr int N_HEAD = 4;  // cyclic-prefix samples per thread: CP <= 4 T

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw1 = reinterpret_cast<float2*>(smem_raw);
    float2* tw2 = tw1 + G::TW1_SIZE;
    const int group = threadIdx.x / T, t = threadIdx.x % T;
    unsigned char* gbase = reinterpret_cast<unsigned char*>(tw2 + G::TW2_SIZE) + size_t(group) * SM::GROUP_BYTES;
    float2* e1 = reinterpret_cast<float2*>(gbase);
    float2* e2 = e1 + G::E1_SIZE;
    uint16_t* stage = reinterpret_cast<uint16_t*>(e2 + G::E2_SIZE);
    float2* red = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(stage) + SM::STAGE_BYTES);

    const int n_items = n_frames * geo.n_chunks;
    const int item = blockIdx.x * GROUPS + group;
    const int frame = (item < n_items) ? item / geo.n_chunks : 0;
    const int chunk = (item < n_items) ? item % geo.n_chunks : 0;
    const FrameDesc desc = descs[frame];
    const bool active = (item < n_items) && desc.valid != 0;
    if (GROUPS == 1) {
        if (!active) return;
    } else {
        if (!__syncthreads_or(active ? 1 : 0)) return;
    }
    fft_load_twiddles<NFFT>(tw1, geo.twiddles, threadIdx.x, FRAME_CTA_THREADS);

    const int S = geo.n_symbols;
    const int s_first = chunk * geo.syms_per_chunk;
    const int s_out_end = min(s_first + geo.syms_per_chunk, S - 1);

    // de-interleaved position of each data-carrying register slot (two per register); unused bins of mixed slots go to the
    // dummy slot NCARR
    constexpr uint32_t USED = D::USED_MASK;
    uint32_t pos_pack[8];
#pragma unroll
    for (int r = 0; r < 16; r += 2) {
        uint32_t lo = NCARR, hi = NCARR;
        if ((USED >> r) & 1u) {
            const int p = geo.bin_to_pos[fft_out_bin<NFFT>(t, r)];
            lo = (p >= 0) ? uint32_t(p) : uint32_t(NCARR);
        }
        if ((USED >> (r + 1)) & 1u) {
            const int p = geo.bin_to_pos[fft_out_bin<NFFT>(t, r + 1)];
            hi = (p >= 0) ? uint32_t(p) : uint32_t(NCARR);
        }
        pos_pack[r / 2] = lo | (hi << 16);
    }

    // per-thread PLL constants: sample i = CP + t + T j of a symbol has i & 3 == k for every j (T is a multiple of 4)
    const int k_fft = (CP + t) & 3;
    float fi_fft = float(CP + t - k_fft);         // float(i & ~3) for j = 0; + T j is exact in float
    const int k_head = (t + T * 12 - TAIL0) & 3;  // head sample index i = t + T j - TAIL0, j >= 12
    float fi_head = float(t + T * 12 - TAIL0 - k_head);
    // keep these as float registers: otherwise the compiler re-derives every fi + T j as an integer add plus an I2FP
    asm volatile("" : "+f"(fi_fft), "+f"(fi_head));

    float2 prev[16];
#pragma unroll
    for (int r = 0; r < 16; r++) prev[r] = make_float2(0.0f, 0.0f);
    int staged = -1;

    const uint64_t ring = desc.mask + 1;  // 0 for a linear buffer
    auto symbol_ptr = [&](int s, bool& contiguous) -> const unsigned char* {
        const uint64_t p0 = uint64_t(desc.start + int64_t(s) * SP) & desc.mask;
        contiguous = (ring == 0) || (p0 + uint64_t(SP) <= ring);
        return reinterpret_cast<const unsigned char*>(desc.src) + p0 * SAMPLE_BYTES;
    };
    // raw samples of one symbol: 16 FFT-window samples + up to 4 cyclic-prefix samples
    auto load_symbol = [&](int s, float2 (&raw)[16], float2 (&head)[N_HEAD]) {
        bool contiguous;
        const unsigned char* base = symbol_ptr(s, contiguous);
        if (contiguous) {
            const unsigned char* mine = base + size_t(t) * SAMPLE_BYTES;
#pragma unroll
            for (int j = 0; j < 16; j++) raw[j] = load_sample_ptr<RAW_U8>(mine + size_t(CP + T * j) * SAMPLE_BYTES);
#pragma unroll
            for (int j = 12; j < 16; j++) {
                const int w0 = T * j - TAIL0;  // head index for t = 0 (may be negative for j = 12)
                head[j - 12] = (w0 + t >= 0) ? load_sample_ptr<RAW_U8>(mine + (ptrdiff_t(w0) * SAMPLE_BYTES)) : make_float2(0.0f, 0.0f);
            }
        } else {  // the symbol straddles the end of the stream ring: masked index per sample
            const uint64_t sym0 = uint64_t(desc.start + int64_t(s) * SP);
#pragma unroll
            for (int j = 0; j < 16; j++) raw[j] = load_sample<RAW_U8>(desc.src, (sym0 + uint64_t(CP + t + T * j)) & desc.mask);
#pragma unroll
            for (int j = 12; j < 16; j++) {
                const int w = t + T * j - TAIL0;
                head[j - 12] = (w >= 0) ? load_sample<RAW_U8>(desc.src, (sym0 + uint64_t(w)) & desc.mask) : make_float2(0.0f, 0.0f);
            }
        }
    };

    auto rotate = [&](float2 x, float base, float pk, float pkc) -> float2 {
        float ts = base + pk, tc = base + pkc;
        ts -= rint_magic(ts);
        tc -= rint_magic(tc);
        const float2 cs = sincos_turns(tc, ts);
        return make_float2(fmaf(cs.x, x.x, -(cs.y * x.y)), fmaf(cs.x, x.y, cs.y * x.x));
    };

    float2 raw[16], head[N_HEAD];
    if (active && s_first <= s_out_end) load_symbol(s_first, raw, head);

    __syncthreads();  // twiddle tables ready

    // one group per CTA: run exactly the symbols this item owns; several groups share the CTA barriers, so they all run the
    // full chunk length and idle (sym_active == false) past their own end
    const int n_iter = (GROUPS == 1) ? (s_out_end - s_first) : geo.syms_per_chunk;
    for (int si = 0; si <= n_iter; si++) {
        const int s = s_first + si;
        const bool sym_active = (GROUPS == 1) ? true : (active && (s <= s_out_end));
        float2 v[16];
        float2 corr = make_float2(0.0f, 0.0f);
        if (sym_active) {
            const float f = desc.freq;
            const float dt0 = float(s * SP) * f;
            if (!D::HAS_TAIL) {
                const float pk = float(k_fft) * f, pkc = pk + 0.25f;
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = rotate(raw[j], fmaf(fi_fft + float(T * j), f, dt0), pk, pkc);
                const float pkh = float(k_head) * f, pkhc = pkh + 0.25f;
#pragma unroll
                for (int j = 12; j < 16; j++) {
                    if (T * j - TAIL0 + t >= 0) {
                        const float2 h = rotate(head[j - 12], fmaf(fi_head + float(T * (j - 12)), f, dt0), pkh, pkhc);
                        const float2 pr = cmul_conj(v[j], h);
                        corr.x += pr.x;
                        corr.y += pr.y;
                    }
                }
            } else {
                const PllSymbol pll = pll_symbol(f, s * SP, SP);
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = pll_rotate(pll, raw[j], CP + t + T * j);
#pragma unroll
                for (int j = 12; j < 16; j++) {
                    const int w = t + T * j - TAIL0;
                    if (w >= 0) {
                        const float2 h = pll_rotate(pll, head[j - 12], w);
                        const float2 pr = cmul_conj(v[j], h);
                        corr.x += pr.x;
                        corr.y += pr.y;
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = make_float2(0.0f, 0.0f);
        }
        // request the next symbol's samples now: they travel while this symbol goes through its FFT
        if (active && (s + 1 <= s_out_end) && si < n_iter) load_symbol(s + 1, raw, head);

        corr = group_reduce_sum<RED_WIDTH>(corr);
        if (WARPS_PER_GROUP > 1 && (t & 31) == 0) red[t >> 5] = corr;

        fft_pass1<NFFT>(v, t, e1, tw1);
        __syncthreads();  // ---- barrier A

        if (sym_active && t == 0 && desc.phase_err != nullptr && (s < s_out_end || s == S - 1)) {
            float2 tot = corr;
            if (WARPS_PER_GROUP > 1) {
                tot = red[0];
#pragma unroll
                for (int w = 1; w < WARPS_PER_GROUP; w++) { tot.x += red[w].x; tot.y += red[w].y; }
            }
            desc.phase_err[s] = atan2f(tot.y, tot.x);
        }
        if (staged >= 0) {
            // 8 carriers per step: 16 bytes of (re, im) pairs -> 8 re bytes + 8 im bytes, [re half | im half] per symbol
            int8_t* out = desc.bits + size_t(staged) * size_t(2 * NCARR);
            const uint4* src4 = reinterpret_cast<const uint4*>(stage);
            for (int i = t; i < NCARR / 8; i += T) {
                const uint4 w = src4[i];
                uint2 re, im;
                re.x = __byte_perm(w.x, w.y, 0x6420);
                re.y = __byte_perm(w.z, w.w, 0x6420);
                im.x = __byte_perm(w.x, w.y, 0x7531);
                im.y = __byte_perm(w.z, w.w, 0x7531);
                *reinterpret_cast<uint2*>(out + 8 * i) = re;
                *reinterpret_cast<uint2*>(out + NCARR + 8 * i) = im;
            }
            staged = -1;
        }

        fft_pass2<NFFT>(v, t, e1, e2, tw2);
        __syncthreads();  // ---- barrier B
        fft_pass3<NFFT>(v, t, e2);

        if (sym_active && desc.fft_tap != nullptr && (s < s_out_end || s == S - 1)) {
#pragma unroll
            for (int r = 0; r < 16; r++) desc.fft_tap[size_t(s) * NFFT + fft_out_bin<NFFT>(t, r)] = v[r];
        }

        if (si > 0 && sym_active) {
            const int s_out = s - 1;
#pragma unroll
            for (int r = 0; r < 16; r++) {
                if ((USED >> r) & 1u) {
                    const uint32_t pos = (r & 1) ? (pos_pack[r / 2] >> 16) : (pos_pack[r / 2] & 0xFFFFu);
                    const float2 d = cmul_conj(prev[r], v[r]);
                    const float a = fmaxf(fabsf(d.x), fabsf(d.y));
                    const float ra = __frcp_rn(a) * 127.00003f;
                    const uint32_t bre = uint32_t(__float2int_rz(-d.x * ra)) & 0xFFu;
                    const uint32_t bim = uint32_t(__float2int_rz(d.y * ra)) & 0xFFu;
                    stage[pos] = uint16_t(bre | (bim << 8));
                    if (desc.vec_tap != nullptr && pos != uint32_t(NCARR)) {
                        const int c = geo.bin_to_carrier[fft_out_bin<NFFT>(t, r)];
                        desc.vec_tap[size_t(s_out) * NCARR + c] = d;
                    }
                }
            }
            staged = s_out;
        }
#pragma unroll
        for (int r = 0; r < 16; r++)
            if ((USED >> r) & 1u) prev[r] = v[r];
    }
    __syncthreads();
    if (staged >= 0) {
        int8_t* out = desc.bits + size_t(staged) * size_t(2 * NCARR);
        const uint4* src4 = reinterpret_cast<const uint4*>(stage);
        for (int i = t; i < NCARR / 8; i += T) {
            const uint4 w = src4[i];
            uint2 re, im;
            re.x = __byte_perm(w.x, w.y, 0x6420);
            re.y = __byte_perm(w.z, w.w, 0x6420);
            im.x = __byte_perm(w.x, w.y, 0x7531);
            im.y = __byte_perm(w.z, w.w, 0x7531);
            *reinterpret_cast<uint2*>(out + 8 * i) = re;
            *reinterpret_cast<uint2*>(out + NCARR + 8 * i) = im;
        }
    }
}

}  // namespace dabb200
