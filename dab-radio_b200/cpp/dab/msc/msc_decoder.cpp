// MSC_Decoder mirror class over the libdab_b200 C ABI.  See msc_decoder.h.
#include "./msc_decoder.h"

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "dab_b200.h"

constexpr int TOTAL_CAPACITY_UNIT_BITS = 64;
constexpr int TOTAL_CAPACITY_UNIT_BYTES = TOTAL_CAPACITY_UNIT_BITS / 8;

MSC_Decoder::MSC_Decoder(const Subchannel subchannel)
    : m_subchannel(subchannel),
      m_nb_encoded_bits(m_subchannel.length * TOTAL_CAPACITY_UNIT_BITS),
      m_nb_encoded_bytes(m_subchannel.length * TOTAL_CAPACITY_UNIT_BYTES) {
    m_decoded_bytes_buf.resize(size_t(m_nb_encoded_bytes));
    if (dab_device_count() < 1) throw std::runtime_error("MSC_Decoder: no usable sm_100 device (there is no CPU fallback)");
}

MSC_Decoder::~MSC_Decoder() { dab_ensemble_destroy(m_handle); }

tcb::span<uint8_t> MSC_Decoder::DecodeCIF(tcb::span<const viterbi_bit_t> buf) {
    const int N = int(buf.size());
    const int start_bit = m_subchannel.start_address * TOTAL_CAPACITY_UNIT_BITS;
    const int end_bit = start_bit + m_nb_encoded_bits;
    if (end_bit > N) return {};   // sub-channel overflows the MSC channel (msc_decoder.cpp:50-54)
    if (m_handle == nullptr || N != m_cif_bits) {
        // the handle is shaped by the CIF size, which the reference learns only here; a CIF of another size restarts the
        // de-interleaver history (the reference never changes it: 55296 bits in every transmission mode)
        dab_ensemble_destroy(m_handle);
        m_handle = nullptr;
        dab_parameters p{};
        p.nb_frame_bits = N;
        p.nb_cifs = 1;
        p.nb_msc_bits = N;
        p.nb_cif_bits = N - N % 64;
        dab_ensemble_options o{};
        o.n_streams = 1;
        const char* e = std::getenv("DAB_B200_DEVICE");
        o.device = e ? std::atoi(e) : 0;
        o.max_subchannels = 1;
        int status = DAB_OK;
        m_handle = dab_ensemble_create(&p, &o, &status);
        if (!m_handle) throw std::runtime_error(std::string("MSC_Decoder: ") + dab_last_error());
        dab_subchannel s{};
        s.id = m_subchannel.id;
        s.start_address = m_subchannel.start_address;
        s.length = m_subchannel.length;
        s.is_uep = m_subchannel.is_uep ? 1 : 0;
        s.uep_prot_index = m_subchannel.uep_prot_index;
        s.eep_prot_level = m_subchannel.eep_prot_level;
        s.eep_type_b = (m_subchannel.eep_type == EEP_Type::TYPE_B) ? 1 : 0;
        if (dab_ensemble_set_subchannels(m_handle, 0, &s, 1) != DAB_OK) throw std::runtime_error(std::string("MSC_Decoder: ") + dab_last_error());
        m_cif_bits = N;
    }
    // a failing call is a CUDA / library error, not "the de-interleaver is still filling": report it
    if (dab_ensemble_decode_frames(m_handle, buf.data(), nullptr) != DAB_OK)
        throw std::runtime_error(std::string("MSC_Decoder::DecodeCIF: ") + dab_last_error());
    int32_t n_bytes = 0;
    if (dab_ensemble_read_msc(m_handle, 0, 0, 0, m_decoded_bytes_buf.data(), m_decoded_bytes_buf.size(), &n_bytes, &m_last_error) != DAB_OK)
        throw std::runtime_error(std::string("MSC_Decoder::DecodeCIF: ") + dab_last_error());
    if (n_bytes <= 0) return {};   // the de-interleaver is still collecting CIFs (msc_decoder.cpp:60-63)
    return {m_decoded_bytes_buf.data(), size_t(n_bytes)};
}
