// Drop-in mirror of the reference's FIC_Decoder (src/dab/fic/fic_decoder.h:15-38) on top of the libdab_b200 C ABI: same
// constructor, DecodeFIBGroup and OnFIB observable, so BasicRadio (src/basic_radio/basic_radio.cpp:51-56, 83-91) compiles and
// behaves unchanged.  One call = one launch of the ensemble decoder's FIC stage for one group (Viterbi PI_16 x21 + PI_15 x3 +
// PI_X, energy dispersal, CRC16 per FIB fused on the GPU: fic_decoder.cpp:53-116); the batched dab_ensemble_* entry points are
// the throughput path, this class is the compatibility path.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <vector>
#include "utility/observable.h"
#include "utility/span.h"
#include "viterbi_config.h"

struct dab_ensemble;

class FIC_Decoder {
private:
    dab_ensemble* m_handle = nullptr;
    std::vector<uint8_t> m_decoded_bytes;
    std::vector<uint8_t> m_fib_valid;
    const size_t m_nb_fibs_per_group;
    const size_t m_nb_encoded_bits;
    const size_t m_nb_decoded_bytes;
    const size_t m_nb_decoded_bits;
    uint64_t m_last_error = 0;
    Observable<tcb::span<const uint8_t>> obs_on_fib;
public:
    // number of bits in FIB (fast information block) group per CIF (common interleaved frame)
    FIC_Decoder(const size_t nb_encoded_bits, const size_t nb_fibs_per_group);
    ~FIC_Decoder();
    FIC_Decoder(const FIC_Decoder&) = delete;
    FIC_Decoder& operator=(const FIC_Decoder&) = delete;
    void DecodeFIBGroup(tcb::span<const viterbi_bit_t> encoded_bits, const size_t cif_index);
    auto& OnFIB(void) { return obs_on_fib; }
    // not in the reference (it only logs the value, fic_decoder.cpp:89-90): Viterbi path error of the last group
    uint64_t GetLastPathError() const { return m_last_error; }
};
