// Drop-in mirror of the reference's OFDM_Demod (src/ofdm/ofdm_demodulator.h:47-168) on top of the libdab_b200 C ABI.
// Same class name, constructor, Process / Reset / On_OFDM_Frame / GetConfig and getters, so callers such as OFDM_Block
// (examples/app_helpers/app_ofdm_blocks.h:25-58) and the GUI (examples/gui/ofdm/render_ofdm_demod.cpp) compile unchanged when
// this directory precedes the reference's src/ on the include path.
//
// Differences a maintainer should know about (all stated in DESIGN.md):
//   * nb_desired_threads is accepted and ignored: symbol-level parallelism is the GPU's.
//   * the On_OFDM_Frame callback fires on the thread that called Process(), before Process() returns, in frame order
//     (the reference fires it on its coordinator thread, possibly after Process() returned).  The span is valid only during
//     the callback, exactly like the reference's.
//   * update ordering is the reference's real-time order (frame k's fine-frequency update precedes frame k+1's PRS sync);
//     the reference's file-mode race between its reader and coordinator threads does not exist here.
//   * GetFrameFFT / GetFrameDataVec carry data only when the GUI taps are enabled (OFDM_Demod::EnableGuiTaps(true) before
//     construction): storing 77 spectra per frame would double the HBM traffic of the hot path.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <complex>
#include <memory>
#include <vector>
#include "utility/observable.h"
#include "utility/span.h"
#include "viterbi_config.h"
#include "ofdm/ofdm_params.h"

struct dab_ofdm;

// field-for-field the reference's OFDM_Demod_Config (ofdm_demodulator.h:24-45)
struct OFDM_Demod_Config {
    struct {
        float update_beta = 0.95f;
        int nb_samples = 100;
        int nb_decimate = 5;
    } signal_l1;
    struct {
        float thresh_null_start = 0.35f;
        float thresh_null_end = 0.75f;
    } null_l1_search;
    struct {
        float fine_freq_update_beta = 0.9f;
        bool is_coarse_freq_correction = true;
        float max_coarse_freq_correction_norm = 0.5f;
        float coarse_freq_slow_beta = 0.1f;
        float impulse_peak_threshold_db = 20.0f;
        float impulse_peak_distance_probability = 0.15f;
    } sync;
};

class OFDM_Demod {
public:
    enum State {
        FINDING_NULL_POWER_DIP,
        READING_NULL_AND_PRS,
        RUNNING_COARSE_FREQ_SYNC,
        RUNNING_FINE_TIME_SYNC,
        READING_SYMBOLS,
    };
    OFDM_Demod(const OFDM_Params& params, const tcb::span<const std::complex<float>> prs_fft_ref, const tcb::span<const int> carrier_mapper,
               int nb_desired_threads = 0);
    ~OFDM_Demod();
    OFDM_Demod(OFDM_Demod&) = delete;
    OFDM_Demod(OFDM_Demod&&) = delete;
    OFDM_Demod& operator=(OFDM_Demod&) = delete;
    OFDM_Demod& operator=(OFDM_Demod&&) = delete;
    void Process(tcb::span<const std::complex<float>> block);
    void Reset();

    OFDM_Params GetOFDMParams() const { return m_params; }
    State GetState() const;
    auto& GetConfig() { return m_cfg; }
    const auto& GetConfig() const { return m_cfg; }
    float GetSignalAverage() const;
    float GetFineFrequencyOffset() const;
    float GetCoarseFrequencyOffset() const;
    float GetNetFrequencyOffset() const;
    int GetFineTimeOffset() const;
    int GetTotalFramesRead() const;
    int GetTotalFramesDesync() const;
    tcb::span<const std::complex<float>> GetFrameFFT() const;
    tcb::span<const std::complex<float>> GetFrameDataVec() const;
    tcb::span<const viterbi_bit_t> GetFrameDataBits() const;
    tcb::span<const float> GetImpulseResponse() const;
    tcb::span<const float> GetCoarseFrequencyResponse() const;
    tcb::span<const std::complex<float>> GetCorrelationTimeBuffer() const;
    auto& On_OFDM_Frame() { return m_obs_on_ofdm_frame; }

    // libdab_b200 specific knobs (process wide, read at construction)
    static void EnableGuiTaps(bool enable);   // default: off
    static void SetDevice(int cuda_ordinal);  // default: 0 (or $DAB_B200_DEVICE)

private:
    struct Snapshot;
    void PushConfigIfChanged();
    const Snapshot& Refresh() const;
    static void FrameTrampoline(void* user, int stream, const int8_t* bits, size_t n_bits, const void* info);

    const OFDM_Params m_params;
    OFDM_Demod_Config m_cfg;
    OFDM_Demod_Config m_cfg_pushed;
    dab_ofdm* m_handle = nullptr;
    size_t m_max_block = 0;
    Observable<tcb::span<const viterbi_bit_t>> m_obs_on_ofdm_frame;
    mutable std::unique_ptr<Snapshot> m_snapshot;
    mutable std::vector<std::complex<float>> m_frame_fft, m_frame_vec, m_corr_time;
    mutable std::vector<viterbi_bit_t> m_frame_bits;
    mutable std::vector<float> m_impulse, m_coarse_response;
};
