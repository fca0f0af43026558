// FIC_Decoder mirror class over the libdab_b200 C ABI.  See fic_decoder.h.
#include "./fic_decoder.h"

#include <assert.h>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "dab_b200.h"

FIC_Decoder::FIC_Decoder(const size_t nb_encoded_bits, const size_t nb_fibs_per_group)
    : m_nb_fibs_per_group(nb_fibs_per_group),
      m_nb_encoded_bits(nb_encoded_bits),
      m_nb_decoded_bytes(nb_encoded_bits / (8 * 3)),
      m_nb_decoded_bits(nb_encoded_bits / 3) {
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
    m_handle = dab_ensemble_create(&p, &o, &status);
    if (!m_handle) throw std::runtime_error(std::string("FIC_Decoder: ") + dab_last_error());   // no CPU fallback
    m_decoded_bytes.resize(m_nb_decoded_bytes);
    m_fib_valid.resize(nb_fibs_per_group ? nb_fibs_per_group : 1);
}

FIC_Decoder::~FIC_Decoder() { dab_ensemble_destroy(m_handle); }

void FIC_Decoder::DecodeFIBGroup(tcb::span<const viterbi_bit_t> encoded_bits, const size_t cif_index) {
    (void)cif_index;
    assert(encoded_bits.size() >= m_nb_encoded_bits);
    // the reference only knows the Mode I puncturing and decodes nothing for other group sizes (fic_decoder.cpp:68-75)
    const size_t nb_decoded_bits_mode_I = (128 * 21 + 128 * 3 + 24) / 4 - 6;
    if (m_nb_decoded_bits != nb_decoded_bits_mode_I) return;
    if (dab_ensemble_decode_frames(m_handle, encoded_bits.data(), nullptr) != DAB_OK) return;
    if (dab_ensemble_read_fic(m_handle, 0, m_decoded_bytes.data(), m_fib_valid.data(), &m_last_error) != DAB_OK) return;
    const size_t nb_fib_bytes = m_nb_decoded_bytes / m_nb_fibs_per_group;
    const size_t nb_crc16_bytes = 2;
    assert(nb_fib_bytes >= nb_crc16_bytes);
    const size_t nb_data_bytes = nb_fib_bytes - nb_crc16_bytes;
    for (size_t i = 0; i < m_nb_fibs_per_group; i++) {
        if (!m_fib_valid[i]) continue;   // CRC16 computed on the device (fic_decoder.cpp:98-115)
        obs_on_fib.Notify(tcb::span<const uint8_t>(m_decoded_bytes.data() + i * nb_fib_bytes, nb_data_bytes));
    }
}
