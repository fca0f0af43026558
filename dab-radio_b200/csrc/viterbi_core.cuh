// Device core of the DAB Viterbi decoder (sm_100a), shared by the plain batch decoder (viterbi.cu) and the ensemble decoder
// (ensemble.cu: CIF de-interleave fused into the symbol fetch, energy dispersal + FIB CRC fused behind the traceback).
//
// Replaces ViterbiDecoder_AVX_u16<7,4> (reference vendor/viterbi_decoder/include/viterbi/x86/viterbi_decoder_avx_u16.h:47-170),
// DAB_Viterbi_Decoder::update / depuncture_symbols (src/dab/algorithms/dab_viterbi_decoder.cpp:114-181) and
// ViterbiDecoder_Core::chainback (viterbi_decoder_core.h:214-236).  See viterbi.cu for the mapping onto a warp.
#pragma once
#include "common.cuh"

namespace dabb200 {


constexpr int VIT_WARPS_PER_CTA = 4;
constexpr uint32_t VIT_MAX_ERROR = 1016;        // (127 - -127) * 4, dab_viterbi_decoder.cpp:31
constexpr uint32_t VIT_NON_START = 5080;        // 5 * max_error, :32-36
constexpr uint32_t VIT_RENORM = 60455;          // 65535 - 5080, :37
constexpr uint32_t VIT_NEAR_SAT = 65535 - 1020; // a metric below this cannot saturate in the next step (e, 1016-e <= 1020)

// device-side form of one update() call, with the depuncture walk pre-digested on the host
struct DevSegment {
    uint32_t first_step;   // trellis step at which the segment starts
    uint32_t n_steps;      // n_out / 4
    uint32_t soft_start;   // punctured symbols consumed by the previous segments
    uint32_t period_syms;  // kept symbols per full cycle of the code
    uint32_t code_len;
    uint32_t counts;       // 8 x 4-bit kept count (1..4) per group
    uint32_t prefix_lo;    // 4 x 8-bit: kept symbols before group r within a cycle, r = 0..3
    uint32_t prefix_hi;    // r = 4..7
};

struct DevSchedule {
    DevSegment seg[DAB_VIT_MAX_SEGMENTS];
    uint32_t n_seg;
    uint32_t total_steps;
    uint32_t n_out_bits;
    uint32_t soft_symbols;
    uint32_t start_state;
    uint32_t end_state;
    uint32_t pad[2];
};

__device__ __forceinline__ uint32_t vabsdiff4_sum(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("vabsdiff4.u32.s32.s32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u));
    return d;
}

__device__ __forceinline__ int parity7(uint32_t v) { return __popc(v & 0x7Fu) & 1; }

struct StepOut {
    uint32_t pack;
    bool d0, d1;
};

// One add-compare-select for butterfly `lane`.  m0m0 / m1m1: predecessor metrics duplicated into both halves.
template <bool SATURATING>
__device__ __forceinline__ StepOut acs(uint32_t m0m0, uint32_t m1m1, uint32_t e, uint32_t inv) {
    StepOut o;
    if (!SATURATING) {
        const uint32_t a = __vadd2(m0m0, e | (inv << 16));   // lo: m0 + e      hi: m0 + (1016 - e)
        const uint32_t b = __vadd2(m1m1, inv | (e << 16));   // lo: m1 + (1016-e) hi: m1 + e
        bool p_hi, p_lo;
        o.pack = __vibmin_u16x2(b, a, &p_hi, &p_lo);         // pred = (b <= a): tie keeps the m1 branch, decision bit 1
        o.d0 = p_lo;
        o.d1 = p_hi;
    } else {
        const uint32_t m0 = m0m0 & 0xFFFFu, m1 = m1m1 & 0xFFFFu;
        const uint32_t a_lo = min(m0 + e, 65535u), b_lo = min(m1 + inv, 65535u);
        const uint32_t a_hi = min(m0 + inv, 65535u), b_hi = min(m1 + e, 65535u);
        o.d0 = b_lo <= a_lo;
        o.d1 = b_hi <= a_hi;
        o.pack = min(a_lo, b_lo) | (min(a_hi, b_hi) << 16);
    }
    return o;
}

// decisions of one step: bit s of w0 = decision of state 2s, bit s of w1 = decision of state 2s+1
__device__ __forceinline__ uint32_t decision_bit(uint2 w, uint32_t state) {
    const uint32_t word = (state & 1u) ? w.y : w.x;
    return (word >> (state >> 1)) & 1u;
}

// One 64-state trellis on one warp.  `view` supplies the punctured soft symbols and receives the decoded bytes:
//     uint32_t View::fetch(uint32_t idx) const     symbol idx of this trellis' punctured input, as an unsigned byte
//     void     View::store(uint32_t byte, uint32_t value)     decoded byte (MSB first), called by lane 0 only
// win: this warp's decision window in shared memory (window_steps entries); spill: global scratch for longer trellises.
// Returns the path error of DAB_Viterbi_Decoder::chainback (dab_viterbi_decoder.cpp:124-129), valid in every lane.
template <class View>
__device__ __forceinline__ uint64_t viterbi_trellis(const DevSchedule* sch, View& view, uint2* win, uint2* spill, uint32_t window_steps, int lane) {
    const uint32_t total_steps = sch->total_steps;
    const bool long_mode = total_steps > window_steps;

    // branch table of ViterbiBranchTable<7,4> (viterbi_branch_table.h:44-52) for butterfly `lane`, one byte per polynomial
    uint32_t table4 = 0;
    {
        const uint32_t G[4] = {109, 79, 83, 109};
#pragma unroll
        for (int r = 0; r < 4; r++) table4 |= (parity7((uint32_t(lane) << 1) & G[r]) ? 0x7Fu : 0x81u) << (8 * r);
    }
    // ViterbiDecoder_Core::reset (viterbi_decoder_core.h:202-211)
    const uint32_t start_state = sch->start_state & 63u;
    uint32_t pack = VIT_NON_START | (VIT_NON_START << 16);
    if ((start_state >> 1) == uint32_t(lane)) pack = (start_state & 1u) ? (pack & 0x0000FFFFu) : (pack & 0xFFFF0000u);
    uint64_t renorm_acc = 0;
    bool near_sat = false;

    const uint32_t sel = (lane & 1) ? 0x3232u : 0x1010u;  // pick half (lane & 1) of the fetched pair and duplicate it
    const int src_a = lane >> 1, src_b = 16 + (lane >> 1);

    for (uint32_t t0 = 0; t0 < total_steps; t0 += 32) {
        // ---- fused depuncture (dab_viterbi_decoder.cpp:131-181): lane i fetches the kept symbols of step t0 + i
        uint32_t my_syms = 0;
        {
            const uint32_t t = t0 + uint32_t(lane);
            if (t < total_steps) {
                int k = 0;
#pragma unroll
                for (int i = 1; i < DAB_VIT_MAX_SEGMENTS; i++)
                    if (i < int(sch->n_seg) && t >= sch->seg[i].first_step) k = i;
                const DevSegment& sg = sch->seg[k];
                const uint32_t g = t - sg.first_step;
                const uint32_t q = g / sg.code_len, r = g - q * sg.code_len;
                const uint32_t prefix = ((r < 4 ? sg.prefix_lo : sg.prefix_hi) >> (8 * (r & 3))) & 0xFFu;
                const uint32_t cnt = (sg.counts >> (4 * r)) & 0xFu;
                const uint32_t idx = sg.soft_start + q * sg.period_syms + prefix;
#pragma unroll
                for (uint32_t j = 0; j < 4; j++)
                    if (j < cnt) my_syms |= view.fetch(idx + j) << (8 * j);
            }
        }
        const uint32_t n_here = min(32u, total_steps - t0);
        for (uint32_t i = 0; i < n_here; i++) {
            const uint32_t sym4 = __shfl_sync(0xFFFFFFFFu, my_syms, int(i));
            const uint32_t e = vabsdiff4_sum(table4, sym4);          // adds_epu16 never saturates: e <= 1020
            const uint32_t inv = (e > VIT_MAX_ERROR) ? 0u : (VIT_MAX_ERROR - e);  // subs_epu16(max_error, e)
            const uint32_t va = __shfl_sync(0xFFFFFFFFu, pack, src_a);
            const uint32_t vb = __shfl_sync(0xFFFFFFFFu, pack, src_b);
            const uint32_t m0m0 = __byte_perm(va, 0, sel), m1m1 = __byte_perm(vb, 0, sel);
            StepOut o = near_sat ? acs<true>(m0m0, m1m1, e, inv) : acs<false>(m0m0, m1m1, e, inv);
            pack = o.pack;
            const uint32_t w0 = __ballot_sync(0xFFFFFFFFu, o.d0);
            const uint32_t w1 = __ballot_sync(0xFFFFFFFFu, o.d1);
            if (lane == 0) win[long_mode ? i : (t0 + i)] = make_uint2(w0, w1);
            // lane 0 votes on the renormalisation test of metric[0]; every other lane on "could saturate next step".
            // (new[1] <= new[0] + 1020, so lane 0's high half is covered by the renormalisation test.)
            const uint32_t lo = pack & 0xFFFFu, hi = pack >> 16;
            const bool pred = (lane == 0) ? (lo >= VIT_RENORM) : (max(lo, hi) >= VIT_NEAR_SAT);
            const uint32_t vote = __ballot_sync(0xFFFFFFFFu, pred);
            near_sat = (vote >> 1) != 0u;
            if (vote & 1u) {  // renormalise (viterbi_decoder_avx_u16.h:138-170)
                uint32_t mn = min(lo, hi);
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
                pack -= mn | (mn << 16);
                renorm_acc += mn;
            }
        }
        if (long_mode) {
            __syncwarp();
            if (uint32_t(lane) < n_here) spill[t0 + uint32_t(lane)] = win[lane];
            __syncwarp();
        }
    }
    __syncwarp();

    // ---- DAB_Viterbi_Decoder::chainback (dab_viterbi_decoder.cpp:124-129): error = sum of renormalisations + metric[0]
    const uint64_t path_error = renorm_acc + uint64_t(__shfl_sync(0xFFFFFFFFu, pack, 0) & 0xFFFFu);

    // ---- ViterbiDecoder_Core::chainback (viterbi_decoder_core.h:214-236) with ViterbiTracebackBuffer<7>
    const uint32_t n_bits = sch->n_out_bits;
    uint32_t reg = (sch->end_state & 63u) << 2;
    if (!long_mode) {
        if (lane == 0) {
            for (int32_t j = int32_t(n_bits) - 1; j >= 0; j--) {
                const uint32_t bit = decision_bit(win[uint32_t(j) + 6u], reg >> 2);
                reg = (reg >> 1) | (bit << 7);
                if ((j & 7) == 0) view.store(uint32_t(j) >> 3, reg & 0xFFu);
            }
        }
    } else {
        // page the decisions back window by window, newest first
        int64_t hi_step = int64_t(n_bits) + 5;  // decision index of bit n_bits-1
        while (hi_step >= 6) {
            const int64_t lo_step = max(int64_t(6), hi_step - int64_t(window_steps) + 1);
            for (int64_t s = lo_step + lane; s <= hi_step; s += 32) win[s - lo_step] = spill[s];
            __syncwarp();
            if (lane == 0) {
                for (int64_t s = hi_step; s >= lo_step; s--) {
                    const uint32_t bit = decision_bit(win[s - lo_step], reg >> 2);
                    reg = (reg >> 1) | (bit << 7);
                    const int64_t j = s - 6;
                    if ((j & 7) == 0) view.store(uint32_t(j >> 3), reg & 0xFFu);
                }
            }
            reg = __shfl_sync(0xFFFFFFFFu, reg, 0);
            __syncwarp();
            hi_step = lo_step - 1;
        }
    }
    __syncwarp();
    return path_error;
}

// host: digest one dab_vit_schedule (the update() calls of one decode) into its device form
inline int digest_schedule(const dab_vit_schedule* s, DevSchedule* d) {
    if (!s || s->n_seg == 0 || s->n_seg > DAB_VIT_MAX_SEGMENTS) return set_error(DAB_ERR_INVALID, "schedule needs 1..%d segments", DAB_VIT_MAX_SEGMENTS);
    memset(d, 0, sizeof(*d));
    uint32_t step = 0, soft = 0;
    for (uint32_t i = 0; i < s->n_seg; i++) {
        const dab_vit_segment& sg = s->seg[i];
        if (sg.code_len < 1 || sg.code_len > 8) return set_error(DAB_ERR_INVALID, "segment %u: code_len %u outside 1..8", i, sg.code_len);
        if (sg.n_out % 4 != 0) return set_error(DAB_ERR_INVALID, "segment %u: requested_output_symbols %u is not a multiple of the code rate", i, sg.n_out);
        DevSegment& o = d->seg[i];
        o.first_step = step;
        o.n_steps = sg.n_out / 4;
        o.soft_start = soft;
        o.code_len = sg.code_len;
        uint32_t prefix = 0;
        for (uint32_t r = 0; r < sg.code_len; r++) {
            if (sg.counts[r] > 4) return set_error(DAB_ERR_INVALID, "segment %u: puncture count %u > 4", i, unsigned(sg.counts[r]));
            o.counts |= uint32_t(sg.counts[r]) << (4 * r);
            if (r < 4) o.prefix_lo |= prefix << (8 * r); else o.prefix_hi |= prefix << (8 * (r - 4));
            prefix += sg.counts[r];
        }
        o.period_syms = prefix;
        const uint32_t full = o.n_steps / sg.code_len, rem = o.n_steps % sg.code_len;
        uint32_t used = full * prefix;
        for (uint32_t r = 0; r < rem; r++) used += sg.counts[r];
        soft += used;
        step += o.n_steps;
    }
    d->n_seg = s->n_seg;
    d->total_steps = step;
    d->n_out_bits = s->n_out_bytes * 8u;
    d->soft_symbols = soft;
    d->start_state = s->start_state;
    d->end_state = s->end_state;
    if (d->n_out_bits + 6u > step) return set_error(DAB_ERR_TRACEBACK, "chainback of %u bits needs %u trellis steps, schedule has %u", d->n_out_bits, d->n_out_bits + 6u, step);
    return DAB_OK;
}

}  // namespace dabb200
