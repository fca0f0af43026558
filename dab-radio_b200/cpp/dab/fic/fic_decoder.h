// Drop-in mirror of the reference's FIC_Decoder (src/dab/fic/fic_decoder.h:15-38) on top of the libdab_b200 C ABI: same
// constructor, DecodeFIBGroup and OnFIB observable, so BasicRadio (src/basic_radio/basic_radio.cpp:51-56, 83-91) compiles and
// behaves unchanged.  One call = one launch of the ensemble decoder's FIC stage for one group (Viterbi PI_16 x21 + PI_15 x3 +
// PI_X, energy dispersal, CRC16 per FIB fused on the GPU: fic_decoder.cpp:53-116); the batched dab_ensemble_* entry points are
// the throughput path, this class is the compatibility path.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "utility/span.h"
#include "utility/observable.h"
#include "viterbi_config.h"

struct dab_ensemble;   // include/dab_b200.h

class FIC_Decoder {
public:
    // nb_encoded_bits: soft bits of one FIB group (one CIF's share of the FIC); nb_fibs_per_group: FIBs in it
    FIC_Decoder(const size_t nb_encoded_bits, const size_t nb_fibs_per_group);
    ~FIC_Decoder();
    FIC_Decoder(const FIC_Decoder&) = delete;
    FIC_Decoder& operator=(const FIC_Decoder&) = delete;

    void DecodeFIBGroup(tcb::span<const viterbi_bit_t> encoded_bits, const size_t cif_index);
    auto& OnFIB(void) { return m_on_fib; }
    // not in the reference (it only logs the value, fic_decoder.cpp:89-90): Viterbi path error of the last group
    uint64_t GetLastPathError() const { return m_path_error; }

private:
    dab_ensemble* m_ensemble = nullptr;            // one stream, one CIF, FIC only
    const size_t m_group_bits;                     // soft bits per group
    const size_t m_fibs;                           // FIBs per group
    const size_t m_group_bytes;                    // decoded bytes per group (rate 1/3 after puncturing)
    std::vector<uint8_t> m_bytes;                  // descrambled group, host copy
    std::vector<uint8_t> m_crc_ok;                 // per FIB: CRC16 verdict computed on the device
    uint64_t m_path_error = 0;
    Observable<tcb::span<const uint8_t>> m_on_fib;
};
