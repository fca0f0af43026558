// OFDM_Demod mirror class over the libdab_b200 C ABI (include/dab_b200.h).  See ofdm_demodulator.h.
#include "./ofdm_demodulator.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "dab_b200.h"

static_assert(sizeof(OFDM_Params) == sizeof(dab_ofdm_params), "OFDM_Params and dab_ofdm_params share one layout");
static_assert(sizeof(std::complex<float>) == sizeof(dab_c32), "std::complex<float> and dab_c32 share one layout");

namespace {
bool g_gui_taps = false;
int g_device = -1;

dab_ofdm_config to_c(const OFDM_Demod_Config& c) {
    dab_ofdm_config o;
    o.signal_l1_update_beta = c.signal_l1.update_beta;
    o.signal_l1_nb_samples = c.signal_l1.nb_samples;
    o.signal_l1_nb_decimate = c.signal_l1.nb_decimate;
    o.null_l1_thresh_null_start = c.null_l1_search.thresh_null_start;
    o.null_l1_thresh_null_end = c.null_l1_search.thresh_null_end;
    o.sync_fine_freq_update_beta = c.sync.fine_freq_update_beta;
    o.sync_is_coarse_freq_correction = c.sync.is_coarse_freq_correction ? 1 : 0;
    o.sync_max_coarse_freq_correction_norm = c.sync.max_coarse_freq_correction_norm;
    o.sync_coarse_freq_slow_beta = c.sync.coarse_freq_slow_beta;
    o.sync_impulse_peak_threshold_db = c.sync.impulse_peak_threshold_db;
    o.sync_impulse_peak_distance_probability = c.sync.impulse_peak_distance_probability;
    return o;
}

bool same(const OFDM_Demod_Config& a, const OFDM_Demod_Config& b) {
    const dab_ofdm_config x = to_c(a), y = to_c(b);
    return std::memcmp(&x, &y, sizeof(x)) == 0;
}

[[noreturn]] void fail(const char* what, int status) {
    throw std::runtime_error(std::string(what) + ": " + dab_last_error() + " (status " + std::to_string(status) + ")");
}
}  // namespace

struct OFDM_Demod::Snapshot {
    dab_ofdm_state state{};
};

void OFDM_Demod::EnableGuiTaps(bool enable) { g_gui_taps = enable; }
void OFDM_Demod::SetDevice(int cuda_ordinal) { g_device = cuda_ordinal; }

OFDM_Demod::OFDM_Demod(const OFDM_Params& params, const tcb::span<const std::complex<float>> prs_fft_ref,
                       const tcb::span<const int> carrier_mapper, int /*nb_desired_threads*/)
    : m_params(params), m_snapshot(std::make_unique<Snapshot>()) {
    if (prs_fft_ref.size() < params.nb_fft || carrier_mapper.size() < params.nb_data_carriers)
        throw std::runtime_error("OFDM_Demod: PRS reference / carrier mapper shorter than the OFDM parameters require");
    int device = g_device;
    if (device < 0) {
        const char* e = std::getenv("DAB_B200_DEVICE");
        device = e ? std::atoi(e) : 0;
    }
    dab_ofdm_params p;
    std::memcpy(&p, &params, sizeof(p));
    dab_ofdm_options opt{};
    opt.n_streams = 1;
    opt.device = device;
    m_max_block = size_t(1) << 20;
    opt.max_block_samples = m_max_block;
    opt.keep_debug_taps = g_gui_taps ? 1 : 0;
    opt.sample_format = DAB_IQ_F32;
    int status = DAB_OK;
    m_handle = dab_ofdm_create(&p, reinterpret_cast<const dab_c32*>(prs_fft_ref.data()), carrier_mapper.data(), &opt, &status);
    if (!m_handle) fail("OFDM_Demod: dab_ofdm_create", status);  // no CPU fallback: a missing B200 is a hard error
    dab_ofdm_set_frame_callback(m_handle, reinterpret_cast<dab_ofdm_frame_cb>(&OFDM_Demod::FrameTrampoline), this);
    m_cfg_pushed = m_cfg;
    m_frame_bits.assign((params.nb_frame_symbols - 1) * params.nb_data_carriers * 2, 0);
    m_frame_fft.assign((params.nb_frame_symbols + 1) * params.nb_fft, {0, 0});
    m_frame_vec.assign((params.nb_frame_symbols - 1) * params.nb_fft, {0, 0});
    m_impulse.assign(params.nb_fft, 0.0f);
    m_coarse_response.assign(params.nb_fft, 0.0f);
    m_corr_time.assign(params.nb_null_period + params.nb_symbol_period, {0, 0});
}

OFDM_Demod::~OFDM_Demod() { dab_ofdm_destroy(m_handle); }

void OFDM_Demod::FrameTrampoline(void* user, int, const int8_t* bits, size_t n_bits, const void*) {
    auto* self = static_cast<OFDM_Demod*>(user);
    self->m_obs_on_ofdm_frame.Notify(tcb::span<const viterbi_bit_t>(bits, n_bits));
}

void OFDM_Demod::PushConfigIfChanged() {
    if (same(m_cfg, m_cfg_pushed)) return;  // GetConfig() hands out a mutable reference (examples/basic_radio_app.cpp:268-269)
    const dab_ofdm_config c = to_c(m_cfg);
    const int rc = dab_ofdm_set_config(m_handle, 0, &c);
    if (rc != DAB_OK) fail("OFDM_Demod: dab_ofdm_set_config", rc);
    m_cfg_pushed = m_cfg;
}

void OFDM_Demod::Process(tcb::span<const std::complex<float>> block) {
    PushConfigIfChanged();
    const auto* data = reinterpret_cast<const dab_c32*>(block.data());
    size_t done = 0;
    // one C-ABI call per Process() call; blocks beyond the ring capacity are fed in max_block pieces
    while (done < block.size() || (block.empty() && done == 0)) {
        const size_t n = std::min(m_max_block, block.size() - done);
        const int rc = dab_ofdm_process(m_handle, 0, data + done, n);
        if (rc != DAB_OK) fail("OFDM_Demod::Process", rc);
        done += n;
        if (block.empty()) break;
    }
}

void OFDM_Demod::Reset() {
    const int rc = dab_ofdm_reset(m_handle, 0);
    if (rc != DAB_OK) fail("OFDM_Demod::Reset", rc);
}

const OFDM_Demod::Snapshot& OFDM_Demod::Refresh() const {
    const int rc = dab_ofdm_get_state(m_handle, 0, &m_snapshot->state);
    if (rc != DAB_OK) fail("OFDM_Demod: dab_ofdm_get_state", rc);
    return *m_snapshot;
}

OFDM_Demod::State OFDM_Demod::GetState() const { return State(Refresh().state.state); }
float OFDM_Demod::GetSignalAverage() const { return Refresh().state.signal_average; }
float OFDM_Demod::GetFineFrequencyOffset() const { return Refresh().state.fine_frequency_offset; }
float OFDM_Demod::GetCoarseFrequencyOffset() const { return Refresh().state.coarse_frequency_offset; }
float OFDM_Demod::GetNetFrequencyOffset() const {
    const auto& s = Refresh().state;
    return s.fine_frequency_offset + s.coarse_frequency_offset;
}
int OFDM_Demod::GetFineTimeOffset() const { return Refresh().state.fine_time_offset; }
int OFDM_Demod::GetTotalFramesRead() const { return Refresh().state.total_frames_read; }
int OFDM_Demod::GetTotalFramesDesync() const { return Refresh().state.total_frames_desync; }

tcb::span<const viterbi_bit_t> OFDM_Demod::GetFrameDataBits() const {
    dab_ofdm_get_frame_data_bits(m_handle, 0, m_frame_bits.data(), m_frame_bits.size());
    return m_frame_bits;
}
tcb::span<const float> OFDM_Demod::GetImpulseResponse() const {
    dab_ofdm_get_impulse_response(m_handle, 0, m_impulse.data(), m_impulse.size());
    return m_impulse;
}
tcb::span<const float> OFDM_Demod::GetCoarseFrequencyResponse() const {
    dab_ofdm_get_coarse_frequency_response(m_handle, 0, m_coarse_response.data(), m_coarse_response.size());
    return m_coarse_response;
}
tcb::span<const std::complex<float>> OFDM_Demod::GetCorrelationTimeBuffer() const {
    dab_ofdm_get_correlation_time_buffer(m_handle, 0, reinterpret_cast<dab_c32*>(m_corr_time.data()), m_corr_time.size());
    return m_corr_time;
}
tcb::span<const std::complex<float>> OFDM_Demod::GetFrameFFT() const {
    // the library keeps the S demodulated symbols; the reference's buffer has one more slot for the NULL symbol (left zero)
    if (g_gui_taps)
        dab_ofdm_get_frame_fft(m_handle, 0, reinterpret_cast<dab_c32*>(m_frame_fft.data()), m_params.nb_frame_symbols * m_params.nb_fft);
    return m_frame_fft;
}
tcb::span<const std::complex<float>> OFDM_Demod::GetFrameDataVec() const {
    // reference layout: symbol i's nb_data_carriers vectors at i * nb_data_carriers (ofdm_demodulator.cpp:734), buffer sized with nb_fft
    if (g_gui_taps)
        dab_ofdm_get_frame_data_vec(m_handle, 0, reinterpret_cast<dab_c32*>(m_frame_vec.data()),
                                    (m_params.nb_frame_symbols - 1) * m_params.nb_data_carriers);
    return m_frame_vec;
}
