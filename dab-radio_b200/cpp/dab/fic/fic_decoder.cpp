// FIC_Decoder mirror class over the libdab_b200 C ABI.  See fic_decoder.h.
#include "./fic_decoder.h"

#include <assert.h>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "dab_b200.h"

FIC_Decoder::FIC_Decoder(const size_t nb_encoded_bits, const size_t nb_fibs_per_group)
    : m_group_bits(nb_encoded_bits), m_fibs(nb_fibs_per_group), m_group_bytes(nb_encoded_bits / (8 * 3)) {
    // a one-stream, one-CIF, FIC-only ensemble: the "frame" of this handle is exactly one FIB group
    dab_parameters p{};
    p.nb_frame_bits = int(nb_encoded_bits);
    p.nb_cifs = 1;
    p.nb_fibs = int(nb_fibs_per_group);
    p.nb_fibs_per_cif = int(nb_fibs_per_group);
    p.nb_fic_bits = int(nb_encoded_bits);
    p.nb_fib_cif_bits = int(nb_encoded_bits);
    p.nb_fib_bits = nb_fibs_per_group ? int(nb_encoded_bits / nb_fibs_per_group) : 0;
    dab_ensemble_options o{};
    o.n_streams = 1;
    const char* e = std::getenv("DAB_B200_DEVICE");
    o.device = e ? std::atoi(e) : 0;
    o.max_subchannels = 1;
    int status = DAB_OK;
    m_ensemble = dab_ensemble_create(&p, &o, &status);
    if (!m_ensemble) throw std::runtime_error(std::string("FIC_Decoder: ") + dab_last_error());   // no CPU fallback
    m_bytes.resize(m_group_bytes);
    m_crc_ok.resize(nb_fibs_per_group ? nb_fibs_per_group : 1);
}

FIC_Decoder::~FIC_Decoder() { dab_ensemble_destroy(m_ensemble); }

void FIC_Decoder::DecodeFIBGroup(tcb::span<const viterbi_bit_t> encoded_bits, const size_t cif_index) {
    (void)cif_index;
    assert(encoded_bits.size() >= m_group_bits);
    // the reference only knows the Mode I puncturing and decodes nothing for other group sizes (fic_decoder.cpp:68-75)
    const size_t nb_decoded_bits_mode_I = (128 * 21 + 128 * 3 + 24) / 4 - 6;
    if (m_group_bits / 3 != nb_decoded_bits_mode_I) return;
    // a failing call is a CUDA / library error, not "no valid FIB": report it like the OFDM_Demod and DAB_Viterbi_Decoder mirrors do
    if (dab_ensemble_decode_frames(m_ensemble, encoded_bits.data(), nullptr) != DAB_OK)
        throw std::runtime_error(std::string("FIC_Decoder::DecodeFIBGroup: ") + dab_last_error());
    if (dab_ensemble_read_fic(m_ensemble, 0, m_bytes.data(), m_crc_ok.data(), &m_path_error) != DAB_OK)
        throw std::runtime_error(std::string("FIC_Decoder::DecodeFIBGroup: ") + dab_last_error());
    const size_t nb_fib_bytes = m_group_bytes / m_fibs;
    const size_t nb_crc16_bytes = 2;
    assert(nb_fib_bytes >= nb_crc16_bytes);
    const size_t nb_data_bytes = nb_fib_bytes - nb_crc16_bytes;
    for (size_t i = 0; i < m_fibs; i++) {
        if (!m_crc_ok[i]) continue;   // CRC16 computed on the device (fic_decoder.cpp:98-115)
        m_on_fib.Notify(tcb::span<const uint8_t>(m_bytes.data() + i * nb_fib_bytes, nb_data_bytes));
    }
}
