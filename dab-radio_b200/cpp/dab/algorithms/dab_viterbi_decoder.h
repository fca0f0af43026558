// Drop-in mirror of the reference's DAB_Viterbi_Decoder (src/dab/algorithms/dab_viterbi_decoder.h:12-45) on top of the
// libdab_b200 C ABI: same constants and the same reset / update / chainback protocol, so FIC_Decoder (fic_decoder.cpp:74-87)
// and MSC_Decoder (msc_decoder.cpp:86-104, 124-145) compile and behave unchanged.
//
// update() only records the segment (and copies the punctured symbols it consumes, as the caller's span dies after the call);
// the trellis runs on the GPU when chainback() is called -- one warp-cooperative decode of everything since reset().
// The batched entry points of the C ABI (dab_viterbi_decode_batch*) are the throughput path; this class is the compatibility path.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <vector>
#include "utility/span.h"
#include "viterbi_config.h"

struct dab_viterbi;

class DAB_Viterbi_Decoder {
public:
    static constexpr size_t m_constraint_length = 7;
    static constexpr size_t m_code_rate = 4;
    DAB_Viterbi_Decoder();
    ~DAB_Viterbi_Decoder();
    DAB_Viterbi_Decoder(const DAB_Viterbi_Decoder&) = delete;
    DAB_Viterbi_Decoder& operator=(const DAB_Viterbi_Decoder&) = delete;
    void set_traceback_length(const size_t traceback_length);
    size_t get_traceback_length() const;
    size_t get_current_decoded_bit() const;
    void reset(const size_t starting_state = 0u);
    size_t update(tcb::span<const viterbi_bit_t> punctured_symbols, tcb::span<const uint8_t> puncture_code, const size_t requested_output_symbols);
    uint64_t chainback(tcb::span<uint8_t> bytes_out, const size_t end_state = 0u);

private:
    struct Segment {
        uint8_t counts[8];
        uint32_t code_len;
        uint32_t n_out;
    };
    dab_viterbi* m_handle = nullptr;
    std::vector<Segment> m_segments;
    std::vector<viterbi_bit_t> m_soft;   // punctured symbols consumed since reset()
    size_t m_traceback_length = 0;
    size_t m_current_decoded_bit = 0;
    size_t m_start_state = 0;
};
