// Lane-per-trellis form of the DAB Viterbi decoder (sm_100a): the bulk path for batches of thousands of trellises.
//
// Same arithmetic as viterbi_core.cuh (the warp-per-trellis form), i.e. ViterbiDecoder_AVX_u16<7,4>
// (reference vendor/viterbi_decoder/include/viterbi/x86/viterbi_decoder_avx_u16.h:47-170), DAB_Viterbi_Decoder::update /
// depuncture_symbols (src/dab/algorithms/dab_viterbi_decoder.cpp:114-181) and ViterbiDecoder_Core::chainback
// (viterbi_decoder_core.h:214-236) -- bit-exact bytes and path error -- but laid out the other way round:
//
//   * one THREAD runs one 64-state trellis; a warp runs 32 trellises in lock step.  The 64 path metrics of a trellis live in
//     32 registers as u16x2 pairs.  Layout L_i pairs the states that differ in bit i: register remove_bit_i(x) = (metric[x],
//     metric[x | 1 << i]).  A trellis step shifts state indices left by one, so two registers of L_i that hold
//     A = (m0 of butterfly j, m0 of butterfly j | 1 << i) and B = (the two m1) produce, with 4 VIADD.16x2 + 2 VIMNMX.U16x2,
//     (new[2j], new[2j | 1 << (i+1)]) = min(A + E, B + E') and (new[2j+1], ...) = min(A + E', B + E): two registers of
//     L_{i+1}, no data movement at all.  That works for i = 1..4; L_5 pairs (x, x + 32), the m0 and m1 of ONE butterfly, so the
//     step out of L_5 regroups two registers with two PRMTs and lands in L_1.  The loop is the 5-step cycle
//     L_5 -> L_1 -> L_2 -> L_3 -> L_4 -> L_5, each phase a separate straight-line instance: 6.4 PRMTs per step on average
//     instead of 32, no shuffles, no shared memory.
//   * the 32 butterflies see only 8 distinct branch-metric patterns (G = {109,79,83,109}: polynomials 0 and 3 coincide), and
//     the partner butterfly j | 1 << i has the pattern of j XOR a per-phase constant, so a step needs 8 VABSDIFF4 and 8 packs.
//     subs(1016, e[p]) is derived from the complementary pattern e[~p] (exact also for -128 symbols, see the step).
//   * the packed adds wrap where the reference saturates.  All metrics lie within 6 * 1020 of metric[0] (every state is
//     reachable from the best one in 6 steps), so a lane looks at the other 63 metrics only while metric[0] >= 58000; it then
//     computes the exact maximum and runs the saturating form of the step when a metric could reach 65535.
//   * survivor decisions (64 bits per step per trellis) stream to a global scratch as [warp][step][lane] uint2: every warp
//     store / traceback load is one coalesced 256-byte row.  The traceback is 32 lanes wide (each lane walks its own trellis).
//   * depuncturing walks the cyclic count table per lane; the punctured symbols are read as aligned 32-bit words and
//     funnel-shifted to the lane's byte position (L1 holds the lane's 128-byte line between refills).
//
// ~7 warp instructions per trellis step instead of ~65 for the warp-per-trellis form, so this is the form used when a batch
// is large enough to fill the GPU with one trellis per thread (viterbi.cu: launch()).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "viterbi_core.cuh"

namespace dabb200 {

constexpr int VITL_THREADS = 32;              // one warp per CTA: the finest grain the block scheduler can balance
constexpr int VITL_MAX_REGS = 128;            // 4 warps per scheduler (16384 registers each): 96 would buy a fifth warp at the price of spills
constexpr int VITL_TB_BYTES = 5;               // traceback: decoded bytes (x 8 decision rows) buffered per lane
#ifndef VITL_TB_PREFETCH
#define VITL_TB_PREFETCH 16                    // traceback: bytes (x 8 rows x 256 B per warp) prefetched into L2 ahead of the row loads
#endif
constexpr uint32_t VITL_CAREFUL = 58000;      // < 65535 - 1020 - 6 * 1020: below this no metric can be near saturation

// Batches of at least this many trellises run one trellis per thread; smaller ones one per warp (viterbi_core.cuh), which
// finishes a handful of trellises sooner.  DAB_B200_VITERBI_LANES=0/1 forces the choice (read at every launch).
constexpr long long VITL_MIN_JOBS = 4096;
inline bool vitl_use_lanes(long long n_jobs) {
    const char* e = getenv("DAB_B200_VITERBI_LANES");
    if (e && *e) return atoi(e) != 0;
    return n_jobs >= VITL_MIN_JOBS;
}

// ---- which warp runs which trellises
// A "group" is 32 trellises that one warp (= one CTA) runs in lock step; its cost is the longest trellis in it (steps).  Warps
// are launched longest group first: the block scheduler hands a freed warp slot to the next CTA in line, so the short groups
// at the end of the line fill the gaps (longest-processing-time-first list scheduling, done by the hardware).  For 1024 Mode I
// ensembles that is 2304 sub-channel groups (1542 steps) and 128 FIC groups (774 steps) on 2368 resident warps: the last 64
// FIC groups start when the first 64 finish, half-way through, and everything ends together -- instead of 64 sub-channel
// groups starting late behind FIC groups that happened to be first in the job list.
struct VitlPlan {
    std::vector<int32_t> groups;              // [n_warps] group id of each warp, longest first
    std::vector<unsigned long long> warp_row; // [n_warps + 1] first decision row (of 32 uint2) of each warp's scratch
    int n_warps() const { return int(groups.size()); }
    unsigned long long rows() const { return warp_row.empty() ? 0ull : warp_row.back(); }
};

inline void vitl_plan(const std::vector<uint32_t>& group_cost, VitlPlan& plan) {
    const int n = int(group_cost.size());
    plan.groups.resize(static_cast<size_t>(n));
    for (int i = 0; i < n; i++) plan.groups[size_t(i)] = i;
    std::stable_sort(plan.groups.begin(), plan.groups.end(), [&](int32_t a, int32_t b) { return group_cost[size_t(a)] > group_cost[size_t(b)]; });
    plan.warp_row.assign(1, 0ull);
    for (int32_t g : plan.groups) plan.warp_row.push_back(plan.warp_row.back() + std::max(8u, (group_cost[size_t(g)] + 1u) & ~1u));
}

__host__ __device__ constexpr uint32_t vitl_par(uint32_t v) { return (v ^ (v >> 1) ^ (v >> 2) ^ (v >> 3) ^ (v >> 4) ^ (v >> 5) ^ (v >> 6)) & 1u; }
// branch pattern of butterfly s: bit 0 = polynomials 0 and 3 (109), bit 1 = polynomial 1 (79), bit 2 = polynomial 2 (83)
// (ViterbiBranchTable<7,4>, viterbi_branch_table.h:44-52)
__host__ __device__ constexpr uint32_t vitl_pattern(uint32_t s) {
    return vitl_par((s << 1) & 109u) | (vitl_par((s << 1) & 79u) << 1) | (vitl_par((s << 1) & 83u) << 2);
}
// the 4 table bytes (+127 -> 0x7F, -127 -> 0x81) of pattern p, one byte per polynomial
__host__ __device__ constexpr uint32_t vitl_table4(uint32_t p) {
    return ((p & 1u) ? 0x7F00007Fu : 0x81000081u) | ((p & 2u) ? 0x00007F00u : 0x00008100u) | ((p & 4u) ? 0x007F0000u : 0x00810000u);
}
static_assert(vitl_pattern(16) == (vitl_pattern(0) ^ 1u) && vitl_pattern(21) == (vitl_pattern(5) ^ 1u), "butterfly k+16 flips pattern bit 0");

__device__ __forceinline__ uint32_t vitl_min_halves(uint32_t x) { return min(x & 0xFFFFu, x >> 16); }
__device__ __forceinline__ uint32_t vitl_max_halves(uint32_t x) { return max(x & 0xFFFFu, x >> 16); }

// Decision gather on the FP32 pipe: acc += 2^k under a predicate, as ONE predicated FADD.  The integer alternatives (@P
// VIADD / LOP3) land on the fma-heavy or alu pipe, which the packed adds, PRMTs and min/max already fill; the fma-lite pipe
// is otherwise idle in this kernel.  acc starts at 2^23, so with k < 16 the sum is exact and the low 16 mantissa bits of the
// float ARE the gathered bits (no conversion: two PRMTs assemble the 64-bit decision word from four accumulators).
__device__ __forceinline__ void vitl_fadd_if(float& acc, bool p, float bit) {
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q add.rn.f32 %0, %0, %2;\n\t}" : "+f"(acc) : "r"(uint32_t(p)), "f"(bit));
}

__host__ __device__ constexpr uint32_t vitl_insert0(uint32_t k, int i) { return ((k >> i) << (i + 1)) | (k & ((1u << i) - 1u)); }
__host__ __device__ constexpr uint32_t vitl_remove(uint32_t x, int i) { return ((x >> (i + 1)) << i) | (x & ((1u << i) - 1u)); }
// pattern difference between butterflies j and j | 1 << i (the parity of (s << 1) & G is linear in s)
__host__ __device__ constexpr uint32_t vitl_flip(int i) { return vitl_pattern(1u << i); }
static_assert(vitl_flip(4) == 1u && vitl_pattern(21) == (vitl_pattern(5) ^ vitl_flip(4)), "partner pattern = pattern ^ flip");

// One trellis step for all 64 states of this lane's trellis, phase PH of the 5-step layout cycle: `in` is in layout L_PH
// (PH = 0: L_5), `out` in L_{PH+1}.  E[p] = (e[p], e[p ^ flip(PH)]), Ei the packed subs(1016, .) of the same.  Decision
// bits: state k < 32 -> bit k of dlo, else bit k - 32 of dhi; 1 = predecessor k/2 + 32 survived, ties included
// (viterbi_decoder_avx_u16.h:114-115).
template <int PH, bool SATURATING>
__device__ __forceinline__ void vitl_acs(const uint32_t (&in)[32], uint32_t (&out)[32], const uint32_t (&E)[8], const uint32_t (&Ei)[8],
                                         uint32_t& dlo, uint32_t& dhi) {
    float f[4] = {8388608.0f, 8388608.0f, 8388608.0f, 8388608.0f};   // decisions of states 0-15, 16-31, 32-47, 48-63
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const uint32_t j = vitl_insert0(uint32_t(q), PH);            // butterfly j, partner j | 1 << PH
        const uint32_t p = vitl_pattern(j);
        uint32_t A, B;
        if (PH == 0) {
            A = __byte_perm(in[j], in[j + 1], 0x5410);                // (metric[j],      metric[j + 1])
            B = __byte_perm(in[j], in[j + 1], 0x7632);                // (metric[j + 32], metric[j + 33])
        } else {
            A = in[vitl_remove(j, PH)];                               // (metric[j],      metric[j | 1 << PH])
            B = in[vitl_remove(j + 32u, PH)];                         // (metric[j + 32], metric[(j | 1 << PH) + 32])
        }
        uint32_t a0, b0, a1, b1;
        if (SATURATING) {
            a0 = __vaddus2(A, E[p]);  b0 = __vaddus2(B, Ei[p]);
            a1 = __vaddus2(A, Ei[p]); b1 = __vaddus2(B, E[p]);
        } else {
            a0 = __vadd2(A, E[p]);  b0 = __vadd2(B, Ei[p]);
            a1 = __vadd2(A, Ei[p]); b1 = __vadd2(B, E[p]);
        }
        const uint32_t s0 = 2u * j, s1 = 2u * j + 1u, partner = 1u << (PH + 1);   // new states in the low halves; high = | partner
        bool ph, pl;
        out[vitl_remove(s0, PH + 1)] = __vibmin_u16x2(b0, a0, &ph, &pl);           // pred = (b <= a)
        vitl_fadd_if(f[s0 >> 4], pl, float(1u << (s0 & 15u)));
        vitl_fadd_if(f[(s0 | partner) >> 4], ph, float(1u << ((s0 | partner) & 15u)));
        out[vitl_remove(s1, PH + 1)] = __vibmin_u16x2(b1, a1, &ph, &pl);
        vitl_fadd_if(f[s1 >> 4], pl, float(1u << (s1 & 15u)));
        vitl_fadd_if(f[(s1 | partner) >> 4], ph, float(1u << ((s1 | partner) & 15u)));
    }
    dlo = __byte_perm(__float_as_uint(f[0]), __float_as_uint(f[1]), 0x5410);
    dhi = __byte_perm(__float_as_uint(f[2]), __float_as_uint(f[3]), 0x5410);
}

template <int N>
struct VitlPhase { static constexpr int value = N; };

// per-lane reader of the punctured symbols of one job + the depuncture walk (dab_viterbi_decoder.cpp:131-181)
struct VitlFeed {
    const uint32_t* words;    // aligned base
    uint32_t bytepos;         // byte position of the next symbol from `words`
    uint32_t widx;            // index of w0
    uint32_t last_word;       // last word that may be read
    uint32_t w0, w1;
    // segment walk
    const DevSegment* segs;
    uint32_t n_seg, k, seg_end, code_len, counts, r;

    __device__ __forceinline__ void open(const int8_t* soft, uint32_t soft_symbols, const DevSchedule* sch) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(soft);
        words = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
        bytepos = uint32_t(a & 3u);
        widx = 0;
        last_word = (bytepos + max(soft_symbols, 1u) - 1u) >> 2;
        w0 = __ldg(words);
        w1 = __ldg(words + min(1u, last_word));
        segs = sch->seg;
        n_seg = sch->n_seg;
        k = 0;
        load_segment();
    }
    __device__ __forceinline__ void load_segment() {
        const DevSegment& sg = segs[k];
        seg_end = sg.first_step + sg.n_steps;
        code_len = sg.code_len;
        counts = sg.counts;
        r = 0;
    }
    // the 4 depunctured symbols of trellis step t as packed bytes (punctured positions = 0)
    __device__ __forceinline__ uint32_t next(uint32_t t) {
        while (t >= seg_end && k + 1 < n_seg) { k++; load_segment(); }
        const uint32_t cnt = (counts >> (4u * r)) & 0xFu;
        r = (r + 1u == code_len) ? 0u : r + 1u;
        const uint32_t raw = __funnelshift_r(w0, w1, (bytepos & 3u) * 8u);
        const uint32_t sym4 = raw & __funnelshift_rc(0xFFFFFFFFu, 0u, 32u - 8u * cnt);
        bytepos += cnt;
        if ((bytepos >> 2) != widx) {
            widx++;
            w0 = w1;
            w1 = __ldg(words + min(widx + 1u, last_word));
        }
        return sym4;
    }
};

// One trellis per lane.  `active` lanes decode (sch, view); the others idle through the warp's loop.
//     View::soft_base()                       first punctured symbol of this lane's job
//     View::store(uint32_t byte, uint32_t v)  decoded byte (MSB first)
// dec: this warp's decision rows, [step][32] uint2.  Returns the path error (dab_viterbi_decoder.cpp:124-129).
template <class View>
__device__ __forceinline__ uint64_t viterbi_lane_trellis(const DevSchedule* sch, View& view, uint2* __restrict__ dec, int lane, bool active) {
    const uint32_t total_steps = active ? sch->total_steps : 0u;

    uint32_t S[32];           // path metrics, layout L_5 at every multiple of 5 steps: S[k] = (metric[k], metric[k + 32])
    uint64_t renorm_acc = 0;
    uint32_t last0 = 0;       // metric[0] after the most recent step
    bool near_sat = false;
    VitlFeed feed;
    if (active) {
        // ViterbiDecoder_Core::reset (viterbi_decoder_core.h:202-211)
        const uint32_t start_state = sch->start_state & 63u;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            uint32_t v = VIT_NON_START | (VIT_NON_START << 16);
            if (start_state == uint32_t(k)) v &= 0xFFFF0000u;
            if (start_state == uint32_t(k + 32)) v &= 0x0000FFFFu;
            S[k] = v;
        }
        last0 = S[0] & 0xFFFFu;
        feed.open(view.soft_base(), sch->soft_symbols, sch);
    }

    uint2* drow = dec + lane;   // this lane's slot in the decision row of the current step
    auto step = [&](auto phase, uint32_t t, const uint32_t (&in)[32], uint32_t (&out)[32]) {
        constexpr int PH = decltype(phase)::value;
        constexpr uint32_t FL = vitl_flip(PH);
        const uint32_t sym4 = feed.next(t);
        uint32_t e8[8], E[8], Ei[8];
#pragma unroll
        for (int p = 0; p < 8; p++) e8[p] = vabsdiff4_sum(vitl_table4(uint32_t(p)), sym4);   // <= 1020: adds_epu16 never saturates
#pragma unroll
        for (int p = 0; p < 8; p++) E[p] = e8[p ^ FL] * 65536u + e8[p];
        // subs_epu16(1016, e[p]) in terms of the complementary pattern: a symbol s contributes |127 - s| + |-127 - s| = 254 to
        // e[p] + e[~p], except s = -128 which contributes 256.  So e[p] + e[~p] = 1016 + 2 * (number of -128 symbols) =: 1016 + c
        // and subs(1016, e[p]) = subs(e[~p], c): one packed max + add per pair, nothing at all to branch on.
        const uint32_t c = e8[0] + e8[7] - VIT_MAX_ERROR;
        const uint32_t c2 = c * 0x00010001u;                          // (c, c)
        const uint32_t nc2 = ((0u - c) & 0xFFFFu) * 0x00010001u;      // (-c, -c) mod 2^16
#pragma unroll
        for (int p = 0; p < 8; p++) Ei[p] = __vadd2(__vmaxu2(E[p ^ 7], c2), nc2);   // max(x, c) - c = subs(x, c)
        uint32_t dlo, dhi;
        if (__builtin_expect(near_sat, 0)) vitl_acs<PH, true>(in, out, E, Ei, dlo, dhi);
        else vitl_acs<PH, false>(in, out, E, Ei, dlo, dhi);
        *drow = make_uint2(dlo, dhi);   // default caching: the newest rows are the first the traceback asks for
        drow += 32;
        near_sat = false;
        uint32_t new0 = out[0] & 0xFFFFu;   // state 0 is the low half of register 0 in every layout
        if (__builtin_expect(new0 >= VITL_CAREFUL, 0)) {
            if (new0 >= VIT_RENORM) {   // renormalise (viterbi_decoder_avx_u16.h:138-170)
                uint32_t m2 = out[0];
#pragma unroll
                for (int k = 1; k < 32; k++) m2 = __vminu2(m2, out[k]);
                const uint32_t mn = vitl_min_halves(m2);
                const uint32_t nmn2 = ((0u - mn) & 0xFFFFu) * 0x00010001u;   // (-mn, -mn): no half goes below zero, mn is the minimum
#pragma unroll
                for (int k = 0; k < 32; k++) out[k] = __vadd2(out[k], nmn2);
                renorm_acc += mn;
                new0 -= mn;
            }
            uint32_t x2 = out[0];
#pragma unroll
            for (int k = 1; k < 32; k++) x2 = __vmaxu2(x2, out[k]);
            near_sat = vitl_max_halves(x2) >= VIT_NEAR_SAT;
        }
        last0 = new0;
    };

    // the 5-phase layout cycle; a lane leaves the loop at its own last step (no warp-wide operation inside)
    for (uint32_t t = 0; t < total_steps; t += 5) {
        uint32_t A1[32], A2[32], A3[32], A4[32];
        step(VitlPhase<0>{}, t, S, A1);
        if (t + 1 >= total_steps) break;
        step(VitlPhase<1>{}, t + 1, A1, A2);
        if (t + 2 >= total_steps) break;
        step(VitlPhase<2>{}, t + 2, A2, A3);
        if (t + 3 >= total_steps) break;
        step(VitlPhase<3>{}, t + 3, A3, A4);
        if (t + 4 >= total_steps) break;
        step(VitlPhase<4>{}, t + 4, A4, S);
    }
    __syncwarp();

    // ---- DAB_Viterbi_Decoder::chainback (dab_viterbi_decoder.cpp:124-129): error = sum of renormalisations + metric[0]
    const uint64_t path_error = active ? renorm_acc + uint64_t(last0) : 0;

    // ---- ViterbiDecoder_Core::chainback (viterbi_decoder_core.h:214-236) with ViterbiTracebackBuffer<7>, a byte at a time:
    // the 8 decision rows of a byte are requested before the serial walk through them
    const uint32_t n_bytes = active ? sch->n_out_bits / 8u : 0u;
    const uint32_t warp_bytes = __reduce_max_sync(0xFFFFFFFFu, n_bytes);
    uint32_t reg = active ? (sch->end_state & 63u) << 2 : 0u;
    auto load_rows = [&](uint2 (&w)[8], int32_t b) {
        if (b >= 0 && uint32_t(b) < n_bytes) {
            const uint2* row = dec + (size_t(b) * 8u + 6u) * 32u + uint32_t(lane);
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = __ldcs(row + size_t(i) * 32u);
        }
    };
    // The traceback starts when every warp of the GPU has just finished its trellis: for a while all that runs is 2368 warps streaming
    // their decision rows back from HBM (the newest ~13 % are still in L2), and what a lane can keep in flight in registers
    // (VITL_TB_BYTES - 1 bytes = 32 rows) left that phase at ~3.6 TB/s, 20 % of the kernel's time for 1 % of its instructions
    // (profiles/r02_viterbi_lanes_ncu.md).  Rows VITL_TB_PREFETCH bytes further down the walk are pulled into L2 by prefetches, one
    // per 32-byte sector: no registers, and the row loads behind them find their data on chip.
    auto prefetch_rows = [&](int32_t b) {
        if (b >= 0 && uint32_t(b) < warp_bytes && (lane & 3) == 0) {   // the sector of lanes lane .. lane + 3, whichever of them are active
            const uint2* row = dec + (size_t(b) * 8u + 6u) * 32u + uint32_t(lane);
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + size_t(i) * 32u));
        }
    };
    auto walk_byte = [&](const uint2 (&w)[8], int32_t b) {
        if (b >= 0 && uint32_t(b) < n_bytes) {
#pragma unroll
            for (int i = 7; i >= 0; i--) {
                const uint32_t state = reg >> 2;
                const uint32_t word = (state & 32u) ? w[i].y : w[i].x;
                const uint32_t bit = (word >> (state & 31u)) & 1u;
                reg = (reg >> 1) | (bit << 7);
            }
            view.store(uint32_t(b), reg & 0xFFu);
        }
    };
    // VITL_TB_BYTES - 1 bytes (8 decision rows each) are in flight per lane while one is walked: the walk is a handful of
    // instructions, so the traceback runs at whatever rate the memory system returns 256-byte rows
    uint2 w[VITL_TB_BYTES][8];
    int32_t b = int32_t(warp_bytes) - 1;
#pragma unroll
    for (int i = 0; i < VITL_TB_BYTES - 1; i++) load_rows(w[i], b - i);
    for (; b >= 0; b -= VITL_TB_BYTES) {
#pragma unroll
        for (int i = 0; i < VITL_TB_BYTES; i++) {
            load_rows(w[(i + VITL_TB_BYTES - 1) % VITL_TB_BYTES], b - i - (VITL_TB_BYTES - 1));
            prefetch_rows(b - i - (VITL_TB_BYTES - 1) - VITL_TB_PREFETCH);
            walk_byte(w[i], b - i);
        }
    }
    return path_error;
}

}  // namespace dabb200
