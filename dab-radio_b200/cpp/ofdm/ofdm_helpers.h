// Mirror of the reference's src/ofdm/ofdm_helpers.h:12-20 (Create_OFDM_Demodulator) plus the three table getters
// (dab_ofdm_params_ref.h, dab_prs_ref.h, dab_mapper_ref.h) for builds outside the reference tree.
#pragma once
#include <complex>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "dab_b200.h"
#include "./ofdm_demodulator.h"

static inline OFDM_Params get_DAB_OFDM_params(const int transmission_mode) {
    dab_ofdm_params p;
    if (dab_get_ofdm_params(transmission_mode, &p) != DAB_OK) throw std::runtime_error(dab_last_error());  // dab_ofdm_params_ref.cpp:54
    OFDM_Params out;
    out.nb_frame_symbols = p.nb_frame_symbols; out.nb_symbol_period = p.nb_symbol_period; out.nb_null_period = p.nb_null_period;
    out.nb_cyclic_prefix = p.nb_cyclic_prefix; out.nb_fft = p.nb_fft; out.nb_data_carriers = p.nb_data_carriers;
    return out;
}
static inline void get_DAB_PRS_reference(const int transmission_mode, tcb::span<std::complex<float>> buf) {
    if (dab_get_prs_reference(transmission_mode, reinterpret_cast<dab_c32*>(buf.data()), buf.size()) != DAB_OK)
        throw std::runtime_error(dab_last_error());                                                         // dab_prs_ref.cpp:142-151
}
static inline void get_DAB_mapper_ref(tcb::span<int> carrier_map, const size_t nb_fft) {
    if (dab_get_mapper_reference(carrier_map.data(), carrier_map.size(), nb_fft) != DAB_OK) throw std::runtime_error(dab_last_error());
}
static inline std::unique_ptr<OFDM_Demod> Create_OFDM_Demodulator(const int transmission_mode, const int total_threads = 0) {
    const OFDM_Params ofdm_params = get_DAB_OFDM_params(transmission_mode);
    auto ofdm_prs_ref = std::vector<std::complex<float>>(ofdm_params.nb_fft);
    get_DAB_PRS_reference(transmission_mode, ofdm_prs_ref);
    auto ofdm_mapper_ref = std::vector<int>(ofdm_params.nb_data_carriers);
    get_DAB_mapper_ref(ofdm_mapper_ref, ofdm_params.nb_fft);
    return std::make_unique<OFDM_Demod>(ofdm_params, ofdm_prs_ref, ofdm_mapper_ref, total_threads);
}
