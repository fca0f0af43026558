// Drop-in mirror of the reference's MSC_Decoder (src/dab/msc/msc_decoder.h:17-41) on top of the libdab_b200 C ABI: same
// constructor from a Subchannel and the same DecodeCIF contract (empty span while the 16-CIF de-interleaver fills or when the
// sub-channel overflows the CIF), so Basic_Audio_Channel / Basic_Data_Packet_Channel (src/basic_radio/) compile unchanged.
// One call = the ensemble decoder's MSC stages for one sub-channel of one CIF on the GPU: CIF_Deinterleaver ring
// (cif_deinterleaver.cpp:21-70), EEP / UEP schedule (msc_decoder.cpp:78-170), Viterbi, energy dispersal.  The batched
// dab_ensemble_* entry points are the throughput path; this class is the compatibility path.
#pragma once

#include <stdint.h>
#include <vector>
#if __has_include("../database/dab_database_entities.h")
#include "../database/dab_database_entities.h"
#else
#include "dab/database/dab_database_entities.h"
#endif
#include "utility/span.h"
#include "viterbi_config.h"

struct dab_ensemble;

class MSC_Decoder {
private:
    const Subchannel m_subchannel;
    const int m_nb_encoded_bits;
    const int m_nb_encoded_bytes;
    std::vector<uint8_t> m_decoded_bytes_buf;
    dab_ensemble* m_handle = nullptr;
    int m_cif_bits = -1;       // size of the CIF the handle was created for
    uint64_t m_last_error = 0;
public:
    explicit MSC_Decoder(const Subchannel subchannel);
    ~MSC_Decoder();
    MSC_Decoder(const MSC_Decoder&) = delete;
    MSC_Decoder& operator=(const MSC_Decoder&) = delete;
    // Returns the number of bytes decoded
    // NOTE: the number of bytes decoded can be 0 if the deinterleaver is still collecting frames
    tcb::span<uint8_t> DecodeCIF(tcb::span<const viterbi_bit_t> buf);
    // not in the reference (it only logs the value, msc_decoder.cpp:106): Viterbi path error of the last CIF
    uint64_t GetLastPathError() const { return m_last_error; }
};
