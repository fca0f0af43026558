// Host-only check of csrc/stage_layout.h on the four DAB carrier maps: the layout is a bijection between soft-bit positions and
// staging slots, every 8-position chunk stays one 16-byte unit (its four words rotated), chunk_src names the chunk's slot and
// rotation, and the searched layout needs no more wavefronts than position order.  Prints one line per mode.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "stage_layout.h"

// ETSI EN 300 401 14.6 frequency interleaving: position n -> carrier k (restated for the test)
static std::vector<int> dab_carrier_order(int nfft, int ncarr) {
    std::vector<int> pi(static_cast<size_t>(nfft), 0);
    const int add = nfft / 4 - 1;
    for (int i = 1; i < nfft; i++) pi[size_t(i)] = (13 * pi[size_t(i - 1)] + add) % nfft;
    const int lo = (nfft - ncarr) / 2, hi = nfft - lo, mid = nfft / 2;
    std::vector<int> ks;
    for (int i = 0; i < nfft; i++) {
        const int d = pi[size_t(i)];
        if (d >= lo && d <= hi && d != mid) ks.push_back(d - mid);
    }
    return ks;
}

int main() {
    const int modes[4][2] = {{2048, 1536}, {512, 384}, {256, 192}, {1024, 768}};
    int bad = 0;
    for (const auto& m : modes) {
        const int nfft = m[0], ncarr = m[1];
        const std::vector<int> ks = dab_carrier_order(nfft, ncarr);
        if (int(ks.size()) != ncarr) { printf("nfft %d: %zu carriers\n", nfft, ks.size()); return 1; }
        std::vector<int16_t> b2p(static_cast<size_t>(nfft), int16_t(-1));
        for (int n = 0; n < ncarr; n++) b2p[size_t((nfft + ks[size_t(n)]) % nfft)] = int16_t(n);
        const int T = nfft / 16, R3 = nfft / 256, lanes = T < 32 ? T : 32;
        std::vector<std::vector<int>> groups;
        for (int t0 = 0; t0 < T; t0 += lanes)
            for (int r = 0; r < 16; r++) {
                std::vector<int> g;
                bool any = false;
                for (int l = 0; l < lanes; l++) {
                    const int t = t0 + l;
                    const int bin = (R3 == 1) ? t + 16 * r : (t + T * (r / R3)) + 256 * (r % R3);
                    g.push_back(bin);
                    any = any || b2p[size_t(bin)] >= 0;
                }
                if (any) groups.push_back(g);
            }
        const dabb200::StageLayout id = dabb200::stage_layout_optimise(groups, b2p, ncarr, false);
        const dabb200::StageLayout lay = dabb200::stage_layout_optimise(groups, b2p, ncarr, true, 20000);
        // identity layout = position order
        for (int bin = 0; bin < nfft; bin++) bad += (id.bin_to_slot[size_t(bin)] != b2p[size_t(bin)]);
        std::vector<int> slot_of_pos(static_cast<size_t>(ncarr), -1), seen(static_cast<size_t>(ncarr), 0);
        for (int bin = 0; bin < nfft; bin++) {
            const int p = b2p[size_t(bin)], s = lay.bin_to_slot[size_t(bin)];
            if ((p < 0) != (s < 0)) { bad++; continue; }
            if (p < 0) continue;
            if (s >= ncarr || seen[size_t(s)]++) bad++;
            slot_of_pos[size_t(p)] = s;
        }
        for (int c = 0; c < ncarr / 8; c++) {
            const int chunk_slot = lay.chunk_src[size_t(c)] >> 2, rot = lay.chunk_src[size_t(c)] & 3;
            for (int i = 0; i < 8; i++) {
                const int want = chunk_slot * 8 + ((((i >> 1) + rot) & 3) << 1) + (i & 1);
                bad += (slot_of_pos[size_t(8 * c + i)] != want);
            }
        }
        bad += (lay.wavefronts_after > lay.wavefronts_before) + (id.wavefronts_after != id.wavefronts_before) + (lay.wavefronts_before != id.wavefronts_before);
        printf("nfft %d: %d store instructions, %d wavefronts in position order, %d as laid out, errors so far %d\n", nfft, lay.store_instructions,
               lay.wavefronts_before, lay.wavefronts_after, bad);
    }
    return bad ? 1 : 0;
}
