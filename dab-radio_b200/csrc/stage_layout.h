// Host side: where the frame kernel stages the soft bits of one symbol in shared memory.
//
// The kernel's threads hold FFT bins in natural order (thread t: bins t + T j) and scatter their (re, im) byte pairs to the
// frequency-de-interleaved soft-bit positions (carrier_mapper, reference ofdm_demodulator.cpp:867-889) of a staging row, which
// then leaves as 16-byte chunks of 8 positions.  With the row in position order a warp's 32 two-byte stores land on pseudo-random
// banks: 3.4 wavefronts per store instead of 1, 11 % of all the kernel's shared-memory wavefronts (profiles/r02_frame_kernel_ncu.md).
// The flush only needs every chunk of 8 positions to stay a 16-byte unit, so the chunks may sit anywhere in the row (a
// permutation of chunk slots = a choice of bank quad per chunk) and their four words may be rotated (the flush rotates them back
// in registers).  stage_layout_optimise() picks, per chunk, the quad and the rotation that minimise the bank conflicts of the
// kernel's store instructions, by a seeded local search (deterministic; a few milliseconds, cached per map by the caller).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace dabb200 {

struct StageLayout {
    std::vector<int16_t> bin_to_slot;   // [nfft] staging slot (2 bytes each) of the bin's soft-bit pair, -1: bin carries no data
    std::vector<uint16_t> chunk_src;    // [ncarr / 8] output chunk i is staged at chunk slot (v >> 2), words rotated by (v & 3)
    int wavefronts_before = 0, wavefronts_after = 0, store_instructions = 0;
};

// groups: for every store instruction of the kernel, the bins its lanes hold (-1: lane idle).  bin_to_pos: [nfft] soft-bit
// position of the bin or -1 (then the lane writes the dummy slot behind the row).  ncarr must be a multiple of 64.
inline StageLayout stage_layout_optimise(const std::vector<std::vector<int>>& groups, const std::vector<int16_t>& bin_to_pos, int ncarr, bool optimise, int max_iter = 60000) {
    const int n_chunks = ncarr / 8;
    const size_t nc = size_t(n_chunks);
    std::vector<int> quad(nc, 0), row(nc, 0), rot(nc, 0);
    for (int c = 0; c < n_chunks; c++) { quad[size_t(c)] = c & 7; row[size_t(c)] = c >> 3; }
    const int dummy_bank = ((ncarr * 2) >> 2) & 31;
    // lanes of a group as (chunk, word-in-chunk) or the dummy
    struct Lane { int chunk, word; };
    std::vector<std::vector<Lane>> glanes;
    std::vector<std::vector<int>> groups_of_chunk(static_cast<size_t>(n_chunks));
    for (const auto& g : groups) {
        std::vector<Lane> lanes;
        for (int bin : g) {
            if (bin < 0) continue;
            const int p = bin_to_pos[size_t(bin)];
            if (p < 0) lanes.push_back({-1, 0});
            else lanes.push_back({p >> 3, (p & 7) >> 1});
        }
        if (lanes.empty()) continue;
        const int gi = int(glanes.size());
        for (const Lane& l : lanes)
            if (l.chunk >= 0) {
                auto& v = groups_of_chunk[size_t(l.chunk)];
                if (v.empty() || v.back() != gi) v.push_back(gi);
            }
        glanes.push_back(std::move(lanes));
    }
    for (auto& v : groups_of_chunk) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    // wavefronts of one store instruction = the largest number of distinct words that fall on one bank.  The search minimises the
    // number of surplus words per bank instead (a smoother objective with the same optimum, every bank hit once) and breaks ties
    // by the wavefront count: 64 * surplus + wavefronts.
    auto group_eval = [&](int gi, bool wavefronts_only) {
        int words[32][8];
        int n[32] = {0};
        int worst = 0, surplus = 0;
        bool dummy_seen = false;
        for (const Lane& l : glanes[size_t(gi)]) {
            int bank, word;
            if (l.chunk < 0) {
                if (dummy_seen) continue;
                dummy_seen = true;
                bank = dummy_bank;
                word = -1;
            } else {
                const int w = (l.word + rot[size_t(l.chunk)]) & 3;
                bank = quad[size_t(l.chunk)] * 4 + w;
                word = (row[size_t(l.chunk)] * 8 + quad[size_t(l.chunk)]) * 4 + w;
            }
            bool dup = false;
            for (int k = 0; k < n[bank] && k < 8; k++) dup = dup || (words[bank][k] == word);
            if (dup) continue;
            if (n[bank] < 8) words[bank][n[bank]] = word;
            n[bank]++;
            if (n[bank] > 1) surplus++;
            worst = std::max(worst, n[bank]);
        }
        return wavefronts_only ? worst : 64 * worst + surplus;
    };
    auto group_cost = [&](int gi) { return group_eval(gi, false); };
    auto total_wavefronts = [&]() {
        int tw = 0;
        for (size_t g = 0; g < glanes.size(); g++) tw += group_eval(int(g), true);
        return tw;
    };
    std::vector<int> cost(glanes.size());
    int total = 0;
    for (size_t g = 0; g < glanes.size(); g++) total += (cost[g] = group_cost(int(g)));
    StageLayout out;
    out.wavefronts_before = total_wavefronts();
    out.store_instructions = int(glanes.size());
    if (optimise && n_chunks >= 8 && n_chunks % 8 == 0) {
        uint64_t rng = 0x9E3779B97F4A7C15ull;
        auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return uint32_t(rng >> 11); };
        std::vector<int> touched;
        for (int it = 0; it < max_iter && total > int(glanes.size()); it++) {   // total == groups: one wavefront each
            const int a = int(next() % uint32_t(n_chunks));
            const bool swap = (next() & 1) != 0;
            const int b = swap ? int(next() % uint32_t(n_chunks)) : a;
            const int ra = rot[size_t(a)], rb = rot[size_t(b)], qa = quad[size_t(a)], qb = quad[size_t(b)], wa = row[size_t(a)], wb = row[size_t(b)];
            touched = groups_of_chunk[size_t(a)];
            if (b != a) touched.insert(touched.end(), groups_of_chunk[size_t(b)].begin(), groups_of_chunk[size_t(b)].end());
            std::sort(touched.begin(), touched.end());
            touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
            int before = 0;
            for (int g : touched) before += cost[size_t(g)];
            if (b != a) {   // the two chunks trade slots; both get a fresh rotation
                std::swap(quad[size_t(a)], quad[size_t(b)]);
                std::swap(row[size_t(a)], row[size_t(b)]);
                rot[size_t(b)] = int(next() & 3);
            }
            rot[size_t(a)] = int(next() & 3);
            int after = 0;
            for (int g : touched) after += group_cost(g);
            if (after <= before) {
                for (int g : touched) cost[size_t(g)] = group_cost(g);
                total += after - before;
            } else {
                quad[size_t(a)] = qa; quad[size_t(b)] = qb; row[size_t(a)] = wa; row[size_t(b)] = wb; rot[size_t(a)] = ra; rot[size_t(b)] = rb;
            }
        }
    }
    out.wavefronts_after = total_wavefronts();
    out.bin_to_slot.assign(bin_to_pos.size(), int16_t(-1));
    for (size_t bin = 0; bin < bin_to_pos.size(); bin++) {
        const int p = bin_to_pos[bin];
        if (p < 0) continue;
        const int c = p >> 3, w = ((p & 7) >> 1), chunk_slot = row[size_t(c)] * 8 + quad[size_t(c)];
        out.bin_to_slot[bin] = int16_t(chunk_slot * 8 + ((w + rot[size_t(c)]) & 3) * 2 + (p & 1));
    }
    out.chunk_src.resize(size_t(n_chunks));
    for (int c = 0; c < n_chunks; c++) out.chunk_src[size_t(c)] = uint16_t(((row[size_t(c)] * 8 + quad[size_t(c)]) << 2) | rot[size_t(c)]);
    return out;
}

}  // namespace dabb200
