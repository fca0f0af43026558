// TEST INFRASTRUCTURE ONLY (oracle/). Not shipped, not on the product path.
//
// C-ABI harness around the UNMODIFIED reference sources, which are compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libdabref.so (git-ignored).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// What it wraps (reference file:line):
//   OFDM_Demod                src/ofdm/ofdm_demodulator.h:47-168, ofdm_demodulator.cpp:235-275
//   OFDM_Modulator            src/ofdm/ofdm_modulator.cpp:49-93
//   tables                    src/ofdm/dab_ofdm_params_ref.cpp:10, dab_prs_ref.cpp:140, dab_mapper_ref.cpp:10
//   dsp                       src/ofdm/dsp/apply_pll.cpp:121, complex_conj_mul_sum.cpp:104
//   DAB_Viterbi_Decoder       src/dab/algorithms/dab_viterbi_decoder.h:12-45
//   FIC_Decoder               src/dab/fic/fic_decoder.h:17-40, fic_decoder.cpp:53-116
//   MSC_Decoder               src/dab/msc/msc_decoder.h:18-43, msc_decoder.cpp:46-170 (with CIF_Deinterleaver, AdditiveScrambler)
//
// Determinism ("real-time order"): the reference's Process() lets the reader thread
// race the pipeline/coordinator threads on m_freq_fine_offset / m_freq_coarse_offset.  In real-time operation
// frame k's fine-frequency update lands before frame k+1's PRS sync.  ref_ofdm_process(..., realtime=1)
// replays the body of OFDM_Demod::Process (ofdm_demodulator.cpp:241-274) verbatim through the private
// stage functions and, after a ReadSymbols() call that dispatched a frame, blocks until the On_OFDM_Frame
// callback has fired.  Nothing else is altered: the stage functions executed are the reference's own.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <complex>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

// The harness TU alone needs to reach the private stage functions.  The reference's own TUs are built untouched.
#define private public
#include <fftw3.h>
#include "ofdm/ofdm_demodulator.h"
#undef private
#include "ofdm/ofdm_demodulator_threads.h"
#include "ofdm/dab_mapper_ref.h"
#include "ofdm/dab_ofdm_params_ref.h"
#include "ofdm/dab_prs_ref.h"
#include "ofdm/dsp/apply_pll.h"
#include "ofdm/dsp/complex_conj_mul_sum.h"
#include "ofdm/ofdm_helpers.h"
#include "ofdm/ofdm_modulator.h"
#include "dab/algorithms/dab_viterbi_decoder.h"
#define private public
#include "dab/fic/fic_decoder.h"
#undef private
#include "dab/msc/msc_decoder.h"
#include "dab/msc/cif_deinterleaver.h"
#include "dab/algorithms/additive_scrambler.h"
#include "dab/database/dab_database_entities.h"

extern "C" {

struct ref_frame_info {
    int64_t frame_start;        // absolute sample index of the PRS cyclic-prefix start chosen by fine time sync
    int32_t fine_time_offset;   // m_fine_time_offset for this frame
    int32_t total_desync;       // m_total_frames_desync when the frame was dispatched
    float coarse_offset;        // m_freq_coarse_offset when the frame was dispatched
    float fine_offset_used;     // m_freq_fine_offset when the frame was dispatched (what the PLL used)
    float fine_offset_after;    // m_freq_fine_offset after this frame's cyclic-prefix update
    float signal_average;       // m_signal_l1_average when the frame was dispatched
};

struct ref_ofdm_state {
    int32_t state;
    int32_t fine_time_offset;
    int32_t total_frames_read;
    int32_t total_frames_desync;
    float signal_average;
    float fine_offset;
    float coarse_offset;
    int32_t pad;
};

}  // extern "C"

namespace {

struct OfdmCtx {
    int mode = 0;
    OFDM_Params params{};
    std::unique_ptr<OFDM_Demod> demod;
    bool collect = true;
    int64_t abs_base = 0;       // samples consumed by previous Process calls
    size_t frame_bits = 0;
    // callback side
    std::mutex mtx;
    std::condition_variable cv;
    size_t frames_done = 0;
    std::vector<std::vector<int8_t>> bits;
    // reader side
    size_t frames_dispatched = 0;
    std::vector<ref_frame_info> infos;
    ref_frame_info pending{};
};

void wait_frames(OfdmCtx& c, size_t n) {
    std::unique_lock<std::mutex> lock(c.mtx);
    c.cv.wait(lock, [&]() { return c.frames_done >= n; });
}

// Body of OFDM_Demod::Process (ofdm_demodulator.cpp:241-274) with the real-time-order wait inserted.
int process_realtime(OfdmCtx& c, tcb::span<const std::complex<float>> buf) {
    OFDM_Demod& d = *c.demod;
    int frames = 0;
    d.UpdateSignalAverage(buf);
    const size_t N = buf.size();
    size_t curr_index = 0;
    while (curr_index < N) {
        auto* block = &buf[curr_index];
        const size_t N_remain = N - curr_index;
        switch (d.m_state) {
        case OFDM_Demod::State::FINDING_NULL_POWER_DIP:
            curr_index += d.FindNullPowerDip({block, N_remain});
            break;
        case OFDM_Demod::State::READING_NULL_AND_PRS:
            curr_index += d.ReadNullPRS({block, N_remain});
            break;
        case OFDM_Demod::State::RUNNING_COARSE_FREQ_SYNC:
            curr_index += d.RunCoarseFreqSync({block, N_remain});
            break;
        case OFDM_Demod::State::RUNNING_FINE_TIME_SYNC:
            curr_index += d.RunFineTimeSync({block, N_remain});
            if (d.m_state == OFDM_Demod::State::READING_SYMBOLS) {
                c.pending.fine_time_offset = d.m_fine_time_offset;
                c.pending.frame_start =
                    c.abs_base + int64_t(curr_index) - int64_t(c.params.nb_symbol_period) + int64_t(d.m_fine_time_offset);
            }
            break;
        case OFDM_Demod::State::READING_SYMBOLS: {
            // no pipeline is in flight here (we waited after the previous dispatch), so these reads are race free
            c.pending.coarse_offset = d.m_freq_coarse_offset;
            c.pending.fine_offset_used = d.m_freq_fine_offset;
            c.pending.signal_average = d.m_signal_l1_average;
            c.pending.total_desync = d.m_total_frames_desync;
            curr_index += d.ReadSymbols({block, N_remain});
            if (d.m_state == OFDM_Demod::State::READING_NULL_AND_PRS) {
                c.frames_dispatched++;
                wait_frames(c, c.frames_dispatched);
                c.pending.fine_offset_after = d.m_freq_fine_offset;
                c.infos.push_back(c.pending);
                frames++;
            }
            break;
        }
        }
    }
    c.abs_base += int64_t(N);
    return frames;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- tables
int ref_get_params(int mode, uint64_t out[6]) {
    try {
        const OFDM_Params p = get_DAB_OFDM_params(mode);
        out[0] = p.nb_frame_symbols; out[1] = p.nb_symbol_period; out[2] = p.nb_null_period;
        out[3] = p.nb_cyclic_prefix; out[4] = p.nb_fft; out[5] = p.nb_data_carriers;
        return 0;
    } catch (const std::exception&) { return -1; }
}

int ref_get_prs(int mode, float* out_interleaved) {
    try {
        const OFDM_Params p = get_DAB_OFDM_params(mode);
        std::vector<std::complex<float>> prs(p.nb_fft);
        get_DAB_PRS_reference(mode, prs);
        std::memcpy(out_interleaved, prs.data(), sizeof(float) * 2 * p.nb_fft);
        return 0;
    } catch (const std::exception&) { return -1; }
}

int ref_get_mapper(int mode, int* out) {
    try {
        const OFDM_Params p = get_DAB_OFDM_params(mode);
        std::vector<int> m(p.nb_data_carriers);
        get_DAB_mapper_ref(m, p.nb_fft);
        std::memcpy(out, m.data(), sizeof(int) * p.nb_data_carriers);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// ---------------------------------------------------------------- dsp
void ref_apply_pll(const float* x, float* y, uint64_t n, float freq_norm, float dt_norm) {
    auto* xi = reinterpret_cast<const std::complex<float>*>(x);
    auto* yo = reinterpret_cast<std::complex<float>*>(y);
    apply_pll_auto({xi, size_t(n)}, {yo, size_t(n)}, freq_norm, dt_norm);
}

void ref_conj_mul_sum(const float* x0, const float* x1, uint64_t n, float out[2]) {
    auto* a = reinterpret_cast<const std::complex<float>*>(x0);
    auto* b = reinterpret_cast<const std::complex<float>*>(x1);
    const auto y = complex_conj_mul_sum_auto({a, size_t(n)}, {b, size_t(n)});
    out[0] = y.real(); out[1] = y.imag();
}

// ---------------------------------------------------------------- modulator (test-vector generator)
// bytes: (nb_frame_symbols-1)*nb_data_carriers*2/8, frame_out: nb_null_period + nb_symbol_period*nb_frame_symbols complex
int ref_modulate(int mode, const uint8_t* bytes, uint64_t nbytes, float* frame_out, uint64_t nsamples) {
    try {
        const OFDM_Params p = get_DAB_OFDM_params(mode);
        std::vector<std::complex<float>> prs(p.nb_fft);
        get_DAB_PRS_reference(mode, prs);
        OFDM_Modulator mod(p, prs);
        auto* out = reinterpret_cast<std::complex<float>*>(frame_out);
        return mod.ProcessBlock({out, size_t(nsamples)}, {bytes, size_t(nbytes)}) ? 0 : -2;
    } catch (const std::exception&) { return -1; }
}

// ---------------------------------------------------------------- OFDM demodulator
void* ref_ofdm_create(int mode, int nb_threads, int collect) {
    try {
        auto* c = new OfdmCtx();
        c->mode = mode;
        c->params = get_DAB_OFDM_params(mode);
        c->collect = collect != 0;
        c->frame_bits = (c->params.nb_frame_symbols - 1) * c->params.nb_data_carriers * 2;
        c->demod = Create_OFDM_Demodulator(mode, nb_threads);
        c->demod->On_OFDM_Frame().Attach([c](tcb::span<const viterbi_bit_t> b) {
            std::lock_guard<std::mutex> lock(c->mtx);
            if (c->collect) c->bits.emplace_back(b.begin(), b.end());
            c->frames_done++;
            c->cv.notify_all();
        });
        return c;
    } catch (const std::exception&) { return nullptr; }
}

void ref_ofdm_destroy(void* h) {
    auto* c = static_cast<OfdmCtx*>(h);
    if (!c) return;
    c->demod.reset();  // joins threads
    delete c;
}

// One Process() call.  realtime=1: deterministic real-time order (see header); realtime=0: the stock racy call.
int ref_ofdm_process(void* h, const float* iq_interleaved, uint64_t n, int realtime) {
    auto* c = static_cast<OfdmCtx*>(h);
    auto* x = reinterpret_cast<const std::complex<float>*>(iq_interleaved);
    if (realtime) return process_realtime(*c, {x, size_t(n)});
    c->demod->Process({x, size_t(n)});
    c->abs_base += int64_t(n);
    return 0;
}

uint64_t ref_ofdm_frames_done(void* h) {
    auto* c = static_cast<OfdmCtx*>(h);
    std::lock_guard<std::mutex> lock(c->mtx);
    return c->frames_done;
}

uint64_t ref_ofdm_frame_bits(void* h) { return static_cast<OfdmCtx*>(h)->frame_bits; }

int ref_ofdm_get_frame(void* h, uint64_t index, ref_frame_info* info, int8_t* bits_out) {
    auto* c = static_cast<OfdmCtx*>(h);
    std::lock_guard<std::mutex> lock(c->mtx);
    if (index >= c->bits.size()) return -1;
    if (info) {
        if (index < c->infos.size()) *info = c->infos[index];
        else std::memset(info, 0, sizeof(*info));
    }
    if (bits_out) std::memcpy(bits_out, c->bits[index].data(), c->bits[index].size());
    return 0;
}

void ref_ofdm_get_state(void* h, ref_ofdm_state* s) {
    auto* c = static_cast<OfdmCtx*>(h);
    const OFDM_Demod& d = *c->demod;
    s->state = int32_t(d.GetState());
    s->fine_time_offset = d.GetFineTimeOffset();
    s->total_frames_read = d.GetTotalFramesRead();
    s->total_frames_desync = d.GetTotalFramesDesync();
    s->signal_average = d.GetSignalAverage();
    s->fine_offset = d.GetFineFrequencyOffset();
    s->coarse_offset = d.GetCoarseFrequencyOffset();
    s->pad = 0;
}

// stage vectors of the most recent frame / sync (GUI getters, ofdm_demodulator.h:133-139)
void ref_ofdm_get_frame_fft(void* h, float* out) {
    auto v = static_cast<OfdmCtx*>(h)->demod->GetFrameFFT();
    std::memcpy(out, v.data(), v.size() * sizeof(std::complex<float>));
}
void ref_ofdm_get_frame_data_vec(void* h, float* out) {
    auto* c = static_cast<OfdmCtx*>(h);
    auto v = c->demod->GetFrameDataVec();
    const size_t n = (c->params.nb_frame_symbols - 1) * c->params.nb_data_carriers;
    std::memcpy(out, v.data(), n * sizeof(std::complex<float>));
}
void ref_ofdm_get_impulse_response(void* h, float* out) {
    auto v = static_cast<OfdmCtx*>(h)->demod->GetImpulseResponse();
    std::memcpy(out, v.data(), v.size() * sizeof(float));
}
void ref_ofdm_get_coarse_freq_response(void* h, float* out) {
    auto v = static_cast<OfdmCtx*>(h)->demod->GetCoarseFrequencyResponse();
    std::memcpy(out, v.data(), v.size() * sizeof(float));
}
// GetCorrelationTimeBuffer() (ofdm_demodulator.h:139): the whole NULL + PRS buffer, nb_null_period + nb_symbol_period samples
void ref_ofdm_get_correlation_time_buffer(void* h, float* out) {
    auto v = static_cast<OfdmCtx*>(h)->demod->GetCorrelationTimeBuffer();
    std::memcpy(out, v.data(), v.size() * sizeof(std::complex<float>));
}

// CPU baseline: n_instances independent demodulators, each fed the same IQ `repeats` times in `block` sample calls
// from its own thread with the reference's stock Process() (file mode).  Returns wall seconds; *frames_out gets the
// total number of frames produced.
double ref_ofdm_bench(int mode, int n_instances, int threads_each, const float* iq_interleaved, uint64_t n,
                      uint64_t block, int repeats, uint64_t* frames_out) {
    std::vector<void*> hs;
    for (int i = 0; i < n_instances; i++) hs.push_back(ref_ofdm_create(mode, threads_each, 0));
    auto* x = reinterpret_cast<const std::complex<float>*>(iq_interleaved);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> workers;
    for (int i = 0; i < n_instances; i++) {
        workers.emplace_back([&, i]() {
            auto* c = static_cast<OfdmCtx*>(hs[size_t(i)]);
            for (int r = 0; r < repeats; r++) {
                for (uint64_t off = 0; off < n; off += block) {
                    const uint64_t len = std::min<uint64_t>(block, n - off);
                    c->demod->Process({x + off, size_t(len)});
                }
            }
        });
    }
    for (auto& w : workers) w.join();
    // destroying joins the pipelines: every dispatched frame is finished before the clock stops
    uint64_t frames = 0;
    std::vector<OfdmCtx*> cs;
    for (auto* h : hs) cs.push_back(static_cast<OfdmCtx*>(h));
    for (auto* c : cs) c->demod.reset();
    const auto t1 = std::chrono::steady_clock::now();
    for (auto* c : cs) { frames += c->frames_done; delete c; }
    if (frames_out) *frames_out = frames;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Persistent pool for bench.py --impl reference: instances stay locked between timed steps.
struct OfdmPool {
    std::vector<void*> hs;
};
void* ref_ofdm_pool_create(int mode, int n_instances, int threads_each) {
    auto* p = new OfdmPool();
    for (int i = 0; i < n_instances; i++) {
        void* h = ref_ofdm_create(mode, threads_each, 0);
        if (!h) { for (auto* x : p->hs) ref_ofdm_destroy(x); delete p; return nullptr; }
        p->hs.push_back(h);
    }
    return p;
}
void ref_ofdm_pool_destroy(void* pool) {
    auto* p = static_cast<OfdmPool*>(pool);
    if (!p) return;
    for (auto* h : p->hs) ref_ofdm_destroy(h);
    delete p;
}
uint64_t ref_ofdm_pool_frames(void* pool) {
    auto* p = static_cast<OfdmPool*>(pool);
    uint64_t n = 0;
    for (auto* h : p->hs) n += ref_ofdm_frames_done(h);
    return n;
}
// every instance consumes the n samples `repeats` times in `block`-sample Process() calls (stock, file-mode Process) from its
// own thread; the clock stops when every dispatched frame's callback has fired.  Returns wall seconds.
double ref_ofdm_pool_run(void* pool, const float* iq_interleaved, uint64_t n, uint64_t block, int repeats) {
    auto* p = static_cast<OfdmPool*>(pool);
    auto* x = reinterpret_cast<const std::complex<float>*>(iq_interleaved);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> workers;
    for (size_t i = 0; i < p->hs.size(); i++) {
        workers.emplace_back([&, i]() {
            auto* c = static_cast<OfdmCtx*>(p->hs[i]);
            for (int r = 0; r < repeats; r++)
                for (uint64_t off = 0; off < n; off += block) c->demod->Process({x + off, size_t(std::min<uint64_t>(block, n - off))});
            // wait for the pipeline of the last dispatched frame (the same wait ReadSymbols performs, ofdm_demodulator.cpp:565),
            // hand the "ended" token back, then let the coordinator's Notify (:632-635) land
            c->demod->m_coordinator->WaitEnd();
            c->demod->m_coordinator->SignalEnd();
            for (int spin = 0; spin < 5000; spin++) {
                {
                    std::lock_guard<std::mutex> lock(c->mtx);
                    if (c->frames_done >= size_t(c->demod->GetTotalFramesRead())) break;
                }
                std::this_thread::sleep_for(std::chrono::microseconds(100));
            }
        });
    }
    for (auto& w : workers) w.join();
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// the same with one IQ buffer per instance (bench.py --impl reference: every instance its own stream, CFO and start offset, as
// the GPU arm's streams)
double ref_ofdm_pool_run_multi(void* pool, const float* const* iq_interleaved, uint64_t n, uint64_t block, int repeats) {
    auto* p = static_cast<OfdmPool*>(pool);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> workers;
    for (size_t i = 0; i < p->hs.size(); i++) {
        workers.emplace_back([&, i]() {
            auto* c = static_cast<OfdmCtx*>(p->hs[i]);
            auto* x = reinterpret_cast<const std::complex<float>*>(iq_interleaved[i]);
            for (int r = 0; r < repeats; r++)
                for (uint64_t off = 0; off < n; off += block) c->demod->Process({x + off, size_t(std::min<uint64_t>(block, n - off))});
            c->demod->m_coordinator->WaitEnd();
            c->demod->m_coordinator->SignalEnd();
            for (int spin = 0; spin < 5000; spin++) {
                {
                    std::lock_guard<std::mutex> lock(c->mtx);
                    if (c->frames_done >= size_t(c->demod->GetTotalFramesRead())) break;
                }
                std::this_thread::sleep_for(std::chrono::microseconds(100));
            }
        });
    }
    for (auto& w : workers) w.join();
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// seconds for `reps` forward transforms of `nfft` points through the FFT the reference was linked against here (the stand-in of
// oracle/fftw3_shim; FFTW3 itself is absent from the image): bounds what the real FFTW could change in the CPU baseline
double ref_fft_bench(int nfft, int reps) {
    const size_t n_pts = static_cast<size_t>(nfft);
    std::vector<std::complex<float>> in(n_pts), out(n_pts);
    for (int i = 0; i < nfft; i++) in[size_t(i)] = {float((i * 37) % 101) / 101.0f - 0.5f, float((i * 53) % 97) / 97.0f - 0.5f};
    fftwf_plan plan = fftwf_plan_dft_1d(nfft, nullptr, nullptr, FFTW_FORWARD, FFTW_ESTIMATE);
    volatile float sink = 0.0f;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) {
        fftwf_execute_dft(plan, reinterpret_cast<fftwf_complex*>(in.data()), reinterpret_cast<fftwf_complex*>(out.data()));
        sink = sink + out[size_t(r % nfft)].real();
    }
    const auto t1 = std::chrono::steady_clock::now();
    fftwf_destroy_plan(plan);
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---------------------------------------------------------------- Viterbi
void* ref_vit_create() { return new DAB_Viterbi_Decoder(); }
void ref_vit_destroy(void* h) { delete static_cast<DAB_Viterbi_Decoder*>(h); }
void ref_vit_set_traceback_length(void* h, uint64_t n) { static_cast<DAB_Viterbi_Decoder*>(h)->set_traceback_length(size_t(n)); }
uint64_t ref_vit_get_traceback_length(void* h) { return static_cast<DAB_Viterbi_Decoder*>(h)->get_traceback_length(); }
uint64_t ref_vit_get_current_decoded_bit(void* h) { return static_cast<DAB_Viterbi_Decoder*>(h)->get_current_decoded_bit(); }
void ref_vit_reset(void* h, uint64_t start_state) { static_cast<DAB_Viterbi_Decoder*>(h)->reset(size_t(start_state)); }
uint64_t ref_vit_update(void* h, const int8_t* soft, uint64_t n_soft, const uint8_t* code, uint64_t code_len, uint64_t n_out) {
    return static_cast<DAB_Viterbi_Decoder*>(h)->update({soft, size_t(n_soft)}, {code, size_t(code_len)}, size_t(n_out));
}
uint64_t ref_vit_chainback(void* h, uint8_t* out, uint64_t nbytes, uint64_t end_state) {
    return static_cast<DAB_Viterbi_Decoder*>(h)->chainback({out, size_t(nbytes)}, size_t(end_state));
}

// One complete job (reset; update per segment; PI_X-style tail included by the caller as a segment; chainback).
// seg_code: n_seg pointers flattened as [n_seg][8] counts with seg_code_len giving the cyclic length.
uint64_t ref_vit_decode_job(void* h, const int8_t* soft, uint64_t n_soft, const uint8_t* seg_codes, const uint32_t* seg_code_len,
                            const uint32_t* seg_n_out, uint32_t n_seg, uint8_t* out, uint64_t n_out_bytes, uint64_t* consumed) {
    auto* d = static_cast<DAB_Viterbi_Decoder*>(h);
    d->reset();
    tcb::span<const int8_t> buf(soft, size_t(n_soft));
    uint64_t used = 0;
    for (uint32_t s = 0; s < n_seg; s++) {
        const size_t n = d->update(buf, {seg_codes + 8 * s, size_t(seg_code_len[s])}, size_t(seg_n_out[s]));
        buf = buf.subspan(n);
        used += n;
    }
    if (consumed) *consumed = used;
    return d->chainback({out, size_t(n_out_bytes)});
}

// CPU baseline: n_threads decoders, each decoding `n_jobs` identical-schedule jobs laid out back to back.
double ref_vit_bench(int n_threads, const int8_t* soft, uint64_t soft_per_job, uint64_t n_jobs, const uint8_t* seg_codes,
                     const uint32_t* seg_code_len, const uint32_t* seg_n_out, uint32_t n_seg, uint64_t traceback_bits,
                     uint8_t* out, uint64_t out_bytes_per_job) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> workers;
    for (int t = 0; t < n_threads; t++) {
        workers.emplace_back([&, t]() {
            DAB_Viterbi_Decoder d;
            d.set_traceback_length(size_t(traceback_bits));
            for (uint64_t j = uint64_t(t); j < n_jobs; j += uint64_t(n_threads)) {
                ref_vit_decode_job(&d, soft + j * soft_per_job, soft_per_job, seg_codes, seg_code_len, seg_n_out, n_seg,
                                   out + j * out_bytes_per_job, out_bytes_per_job, nullptr);
            }
        });
    }
    for (auto& w : workers) w.join();
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---------------------------------------------------------------- FIC / MSC decoders
// FIC_Decoder::DecodeFIBGroup: `out` receives the descrambled group (m_decoded_bytes, nb_encoded_bits / 24 bytes) whether or not
// the CRCs match; valid[i] = 1 when FIB i was passed to OnFIB.  Returns the number of valid FIBs.
struct FicCtx {
    std::unique_ptr<FIC_Decoder> dec;
    size_t nb_fibs = 0;
    std::vector<uint8_t> valid;
    std::vector<const uint8_t*> seen;
};
void* ref_fic_create(uint64_t nb_encoded_bits, uint64_t nb_fibs_per_group) {
    auto* c = new FicCtx();
    c->dec = std::make_unique<FIC_Decoder>(size_t(nb_encoded_bits), size_t(nb_fibs_per_group));
    c->nb_fibs = size_t(nb_fibs_per_group);
    c->dec->OnFIB().Attach([c](tcb::span<const uint8_t> buf) { c->seen.push_back(buf.data()); });
    return c;
}
void ref_fic_destroy(void* h) { delete static_cast<FicCtx*>(h); }
int ref_fic_decode_group(void* h, const int8_t* bits, uint64_t n_bits, uint64_t cif_index, uint8_t* out, uint8_t* valid) {
    auto* c = static_cast<FicCtx*>(h);
    c->seen.clear();
    std::fill(c->dec->m_decoded_bytes.begin(), c->dec->m_decoded_bytes.end(), uint8_t(0));
    c->dec->DecodeFIBGroup({bits, size_t(n_bits)}, size_t(cif_index));
    const auto& bytes = c->dec->m_decoded_bytes;
    std::memcpy(out, bytes.data(), bytes.size());
    const size_t fib_bytes = bytes.size() / c->nb_fibs;
    for (size_t i = 0; i < c->nb_fibs; i++) valid[i] = 0;
    for (const uint8_t* p : c->seen) valid[size_t(p - bytes.data()) / fib_bytes] = 1;
    return int(c->seen.size());
}

// MSC_Decoder for one sub-channel; DecodeCIF returns the descrambled bytes (0 while the de-interleaver is still filling)
void* ref_msc_create(int start_address, int length, int is_uep, int uep_prot_index, int eep_prot_level, int eep_type_b) {
    Subchannel sc(0);
    sc.start_address = subchannel_addr_t(start_address);
    sc.length = subchannel_size_t(length);
    sc.is_uep = is_uep != 0;
    sc.uep_prot_index = uep_protection_index_t(uep_prot_index);
    sc.eep_prot_level = eep_protection_level_t(eep_prot_level);
    sc.eep_type = eep_type_b ? EEP_Type::TYPE_B : EEP_Type::TYPE_A;
    sc.is_complete = true;
    return new MSC_Decoder(sc);
}
void ref_msc_destroy(void* h) { delete static_cast<MSC_Decoder*>(h); }
int64_t ref_msc_decode_cif(void* h, const int8_t* cif_bits, uint64_t n_bits, uint8_t* out, uint64_t out_capacity) {
    auto res = static_cast<MSC_Decoder*>(h)->DecodeCIF({cif_bits, size_t(n_bits)});
    if (res.size() > out_capacity) return -1;
    std::memcpy(out, res.data(), res.size());
    return int64_t(res.size());
}
// CIF_Deinterleaver alone (cif_deinterleaver.cpp:21-70): Consume then Deinterleave; returns 1 when output was produced
void* ref_deint_create(int nb_bytes) { return new CIF_Deinterleaver(nb_bytes); }
void ref_deint_destroy(void* h) { delete static_cast<CIF_Deinterleaver*>(h); }
int ref_deint_push(void* h, const int8_t* bits, int8_t* out, uint64_t n_bits) {
    auto* d = static_cast<CIF_Deinterleaver*>(h);
    d->Consume({bits, size_t(n_bits)});
    return d->Deinterleave({out, size_t(n_bits)}) ? 1 : 0;
}
// AdditiveScrambler with syncword 0xFFFF (additive_scrambler.h:10-35)
void ref_scrambler_bytes(uint16_t syncword, uint8_t* out, uint64_t n) {
    AdditiveScrambler s;
    s.SetSyncword(syncword);
    s.Reset();
    for (uint64_t i = 0; i < n; i++) out[i] = s.Process();
}

const char* ref_build_info() {
    return "reference sources compiled unmodified; FFT = oracle/fftw3_shim (double radix-2, rounded once to float; FFTW3 absent); "
           "flags -O2 -march=x86-64-v3 -ffast-math; Viterbi = ViterbiDecoder_AVX_u16<7,4>";
}

}  // extern "C"
