// DAB_Viterbi_Decoder mirror class over the libdab_b200 C ABI.  See dab_viterbi_decoder.h.
#include "./dab_viterbi_decoder.h"

#include <assert.h>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "dab_b200.h"

DAB_Viterbi_Decoder::DAB_Viterbi_Decoder() {
    const char* e = std::getenv("DAB_B200_DEVICE");
    int status = DAB_OK;
    m_handle = dab_viterbi_create(e ? std::atoi(e) : 0, &status);
    if (!m_handle) throw std::runtime_error(std::string("DAB_Viterbi_Decoder: ") + dab_last_error());  // no CPU fallback
}

DAB_Viterbi_Decoder::~DAB_Viterbi_Decoder() { dab_viterbi_destroy(m_handle); }

void DAB_Viterbi_Decoder::set_traceback_length(const size_t traceback_length) {
    // ViterbiDecoder_Core::set_traceback_length (viterbi_decoder_core.h:180-187)
    m_traceback_length = traceback_length;
    const size_t new_length = traceback_length + (m_constraint_length - 1);
    if (m_current_decoded_bit > new_length) m_current_decoded_bit = new_length;
}

size_t DAB_Viterbi_Decoder::get_traceback_length() const { return m_traceback_length; }
size_t DAB_Viterbi_Decoder::get_current_decoded_bit() const { return m_current_decoded_bit; }

void DAB_Viterbi_Decoder::reset(const size_t starting_state) {
    m_segments.clear();
    m_soft.clear();
    m_current_decoded_bit = 0;
    m_start_state = starting_state;
}

size_t DAB_Viterbi_Decoder::update(tcb::span<const viterbi_bit_t> punctured_symbols, tcb::span<const uint8_t> puncture_code,
                                   const size_t requested_output_symbols) {
    assert(requested_output_symbols % m_code_rate == 0);
    assert(puncture_code.size() >= 1 && puncture_code.size() <= 8);
    // depuncture_symbols (dab_viterbi_decoder.cpp:131-181): how many punctured symbols this segment consumes
    size_t consumed = 0, code_index = 0;
    for (size_t out = 0; out < requested_output_symbols; out += m_code_rate) {
        const size_t take = puncture_code[code_index];
        // input underrun: the reference returns an all-zero result and decodes nothing for this call (:158-162)
        if (punctured_symbols.size() - consumed < take) return 0;
        consumed += take;
        code_index = (code_index + 1) % puncture_code.size();
    }
    assert(m_current_decoded_bit + requested_output_symbols / m_code_rate <= m_traceback_length + (m_constraint_length - 1));
    Segment seg{};
    for (size_t i = 0; i < puncture_code.size(); i++) seg.counts[i] = puncture_code[i];
    seg.code_len = uint32_t(puncture_code.size());
    seg.n_out = uint32_t(requested_output_symbols);
    // The reference accepts any number of update() calls.  A call that continues the previous call's code where that one stopped
    // (same code, previous call ended on a code boundary) extends the previous segment; anything else starts a new one.
    bool merged = false;
    if (!m_segments.empty()) {
        Segment& last = m_segments.back();
        const bool same_code = last.code_len == seg.code_len && std::memcmp(last.counts, seg.counts, sizeof(seg.counts)) == 0;
        if (same_code && (last.n_out / m_code_rate) % last.code_len == 0) {
            last.n_out += seg.n_out;
            merged = true;
        }
    }
    if (!merged) {
        if (m_segments.size() >= DAB_VIT_MAX_SEGMENTS)
            throw std::runtime_error("DAB_Viterbi_Decoder: more than " + std::to_string(DAB_VIT_MAX_SEGMENTS) +
                                     " distinct puncturing segments between reset() and chainback()");
        m_segments.push_back(seg);
    }
    m_soft.insert(m_soft.end(), punctured_symbols.begin(), punctured_symbols.begin() + consumed);
    m_current_decoded_bit += requested_output_symbols / m_code_rate;
    return consumed;
}

uint64_t DAB_Viterbi_Decoder::chainback(tcb::span<uint8_t> bytes_out, const size_t end_state) {
    const size_t total_bits = bytes_out.size() * 8u;
    assert(m_traceback_length >= total_bits);                                     // viterbi_decoder_core.h:216-218
    assert(m_current_decoded_bit >= total_bits + (m_constraint_length - 1));
    if (m_segments.empty() || bytes_out.empty()) {
        // nothing decoded since reset(): the reference chains back over zeroed decisions (all predecessors 0) and reports the
        // start metric of state 0 -- zero bytes, error 0 when the trellis started there, the initial penalty otherwise
        // (dab_viterbi_decoder.cpp:31-41: non-start states begin at 5080)
        std::memset(bytes_out.data(), 0, bytes_out.size());
        return (m_start_state == 0) ? 0u : 5080u;
    }
    dab_vit_schedule s{};
    s.n_seg = uint32_t(m_segments.size());
    for (size_t i = 0; i < m_segments.size(); i++) {
        std::memcpy(s.seg[i].counts, m_segments[i].counts, 8);
        s.seg[i].code_len = m_segments[i].code_len;
        s.seg[i].n_out = m_segments[i].n_out;
    }
    s.n_out_bytes = uint32_t(bytes_out.size());
    s.start_state = uint32_t(m_start_state);
    s.end_state = uint32_t(end_state);
    uint64_t error = 0;
    const int rc = dab_viterbi_decode_one(m_handle, &s, m_soft.data(), m_soft.size(), bytes_out.data(), &error);
    if (rc != DAB_OK) throw std::runtime_error(std::string("DAB_Viterbi_Decoder::chainback: ") + dab_last_error());
    return error;
}
