// Per-stream receive state machine on the device: OFDM_Demod::Process and its five states (reference
// ofdm_demodulator.cpp:235-577), one CTA per stream.  All per-stream state lives in HBM (StreamState); the samples are
// never copied into frame / correlation buffers as the reference does (ofdm_frame_buffer.h, reconstruction_buffer.h) --
// the stream ring (or the caller's resident buffer) is addressed by absolute sample index instead:
//   correlation buffer element i  ->  explicit copy (first corr_explicit_len elements, cold start only) or stream[corr_base + i]
//   frame buffer element j        ->  stream[frame_start + j]
// Ordering is the reference's real-time order: the fine-frequency update of frame k (CoordinatorThread,
// ofdm_demodulator.cpp:608-618) is applied before frame k+1's PRS synchronisation.  A control pass therefore stops at a
// frame dispatch; the host launches  control -> frame kernel -> control ...  and the next control pass first folds the
// frame kernel's per-symbol phase errors into the fine frequency offset, then continues with the remaining samples.
//
// What the frame kernel is handed is a RANGE of symbols of a frame (FrameDesc): once a frame's PRS is synchronised its net
// frequency offset is fixed (the only update in between is the previous frame's, which precedes the sync), so its symbols can be
// demodulated in the call they arrive in instead of all at once when the frame completes.  Every sample then passes through
// shared memory in the call that delivered it, and the frame kernel sums that call's UpdateSignalAverage windows
// (ofdm_demodulator.cpp:934-950) on the way; the control kernel only evaluates the windows no item covered (NULL symbol,
// the unfinished symbol at the end of a block, streams that are not locked) and folds all of them into the running average in
// window order at the end of the call -- or earlier, the moment FindNullPowerDip needs the value.
#pragma once
#include "ofdm_device.cuh"
#include "ofdm_frame.cuh"
#include "ofdm_frame_v3.cuh"

namespace dabb200 {

constexpr int CTRL_MAX_OWNED = 8;    // UpdateSignalAverage window ranges per call that frame-kernel items may own (one per dispatch)
constexpr int FRAME_MAX_CHUNKS = 4;  // work items one dispatch is split into (load balance: a frame is 76 - 153 symbols long)

struct StreamState {
    // --- mirrors of the OFDM_Demod members (ofdm_demodulator.h:58-75)
    int32_t state;
    int32_t total_frames_read;
    int32_t total_frames_desync;
    int32_t is_found_coarse;
    float freq_coarse;
    float freq_fine;
    int32_t fine_time_offset;
    int32_t null_start_found;
    int32_t null_end_found;
    float l1_average;
    // --- CircularBuffer m_null_power_dip_buffer (circular_buffer.h): index persists across SetLength(0)
    uint32_t ring_index;
    uint32_t ring_length;
    // --- ReconstructionBuffer m_correlation_time_buffer (reconstruction_buffer.h)
    uint32_t corr_length;
    uint32_t corr_explicit_len;
    int64_t corr_base;       // absolute sample index of correlation-buffer element 0
    // --- OFDM_Frame_Buffer (ofdm_frame_buffer.h): virtual, [frame_start, frame_start + frame_cap)
    int64_t frame_start;
    float frame_freq;        // net PLL frequency of the frame being received (fixed once its PRS is synchronised)
    int32_t symbols_done;    // symbols [0, symbols_done) of that frame are demodulated or queued for the frame kernel
    int32_t frame_slot;      // soft-bit buffer (ring slot) the frame is written to
    // --- cursor / current Process() call
    int64_t consumed;        // absolute index of the next unread sample
    int64_t call_begin;
    int64_t call_end;
    int32_t avg_pending;         // UpdateSignalAverage of this call not folded into l1_average yet (see fold_average)
    int32_t pipeline_pending;    // a frame was dispatched; its phase errors still have to update the fine offset
    int32_t pending_slot;        // soft-bit ring slot of the most recently completed frame
    int32_t frames_in_call;      // frames completed during this call
    int32_t syncs_in_call;       // frames whose PRS was synchronised during this call (ring slot choice)
    int32_t own_n;               // window ranges of this call owned by frame-kernel items
    int32_t own_lo[CTRL_MAX_OWNED], own_hi[CTRL_MAX_OWNED];
    dab_ofdm_config cfg;
    dab_ofdm_frame_info pending_info;
};

struct ControlGeom {
    int n_symbols, symbol_period, null_period, cyclic_prefix, n_carriers;
    int slots;               // frames a stream can complete per call
    int ring_slots;          // soft-bit buffers per stream: slots + 1 (the frame being received has one too)
    int n_streams;           // streams of the handle (row length of descs)
    int stream0;             // first stream this launch covers (launches are split into pipeline ways, see run_call)
    int n_chunks;            // work items per dispatch (<= FRAME_MAX_CHUNKS)
    int syms_per_chunk;      // target symbols per work item
    int frame_passes;        // control passes of this call that are followed by a frame-kernel launch
    int eager;               // 1: symbols of a frame still being received are demodulated in the call they arrive in
    int frame_owns_l1;       // 1: the frame kernel sums the UpdateSignalAverage windows inside the symbols it reads
    int l1_per_symbol;       // ... at most this many per symbol
    uint32_t call_index;     // calls made on this handle so far (soft-bit ring slot choice)
    size_t frame_bits;
    uint64_t mask;           // stream index mask (ring size - 1 or ~0)
    uint64_t limit;          // samples addressable per stream (ring size, or the attached buffer's length)
    size_t stream_stride;    // samples between consecutive streams' bases
    const void* samples;     // base of stream 0
    SampleFmt fmt;
    float2* ring;            // [n_streams][null_period] null-power-dip ring
    float2* corr_explicit;   // [n_streams][null_period] cold-start copy of the ring into the correlation buffer
    const float2* prs_fft_ref_conj;   // [NFFT]  conj(PRS)                               (ofdm_demodulator.cpp:130-132)
    const float2* prs_time_ref_conj;  // [NFFT]  conj(IFFT(relative phase of PRS))       (ofdm_demodulator.cpp:134-140)
    float* impulse_response;          // [n_streams][NFFT]
    float* freq_response;             // [n_streams][NFFT]
    StreamState* states;
    FrameDesc* descs;        // [frame_passes][n_streams][n_chunks]
    dab_ofdm_frame_info* infos;  // [n_streams][slots]
    int32_t* frames_in_call;     // [n_streams]
    int32_t* frame_slots;        // [n_streams][slots]: ring slot of the f-th frame completed in this call
    int8_t* bits;            // [n_streams][ring_slots][frame_bits]
    float* phase_err;        // [n_streams][n_symbols]
    const float2* twiddles;  // precomputed FFT twiddles (fft_twiddle_init_kernel)
    float* l1_windows;       // [n_streams][l1_windows_stride] window averages of the current call summed by the frame kernel
    int l1_windows_stride;
    float2* fft_tap;         // optional [n_streams][n_symbols * NFFT]
    float2* vec_tap;         // optional [n_streams][(n_symbols-1) * n_carriers]
    const uint64_t* n_per_stream;  // samples of this call per stream, nullptr: n_uniform
    uint64_t n_uniform;
};

constexpr int CTRL_L1_BATCH = 1024;  // windows whose L1 averages are computed in parallel before the sequential scan

template <int NFFT>
struct ControlSmem {
    using G = FftGeom<NFFT>;
    // at least four warps whatever the FFT size: the transforms use the first N/16 threads, the L1 window scans (FindNullPowerDip
    // walks whole blocks while a stream is unlocked) use all of them -- with one warp per stream a 1024-stream Mode II / III step
    // spent 0.5 ms in the scans of the unlocked streams (profiles/r02a_mode_probe_baseline.txt)
    static constexpr int THREADS = (G::T < 128) ? 128 : G::T;
    // `nat` (natural-order spectrum between two transforms) aliases exchange 1 and the L1 window batch aliases the (contiguous) exchanges: both
    // are only alive while no transform is in flight (every hand-over is separated by a CTA barrier)
    static_assert(G::E1_SIZE >= NFFT, "nat must fit into exchange 1");
    static_assert(size_t(G::E1_SIZE + G::E2_SIZE) * sizeof(float2) >= size_t(CTRL_L1_BATCH) * sizeof(float), "L1 batch must fit into the exchanges");
    static constexpr size_t bytes() { return size_t(G::TW1_SIZE + G::TW2_SIZE + G::E1_SIZE + G::E2_SIZE) * sizeof(float2) + 32 * sizeof(float2) + 64; }
};

struct ArgMax {
    float value;
    int index;
};
// larger value wins, equal values keep the smaller index (the reference scans upwards with a strict '>')
__device__ __forceinline__ ArgMax argmax_combine(ArgMax a, ArgMax b) {
    const bool take_b = (b.value > a.value) || (b.value == a.value && b.index < a.index);
    return take_b ? b : a;
}

template <int NFFT, int SB>
struct Control {
    using G = FftGeom<NFFT>;
    static constexpr int T = G::T;
    static constexpr int THREADS = ControlSmem<NFFT>::THREADS;

    const ControlGeom& geo;
    StreamState& st;  // shared-memory copy, mutated by thread 0 between barriers
    const int stream, tid;
    float2 *tw1, *tw2, *e1, *e2, *nat;
    float* l1buf;
    float2* red;
    const void* src;

    __device__ float2 sample(int64_t abs_index) const { return load_sample<SB>(src, uint64_t(abs_index) & geo.mask, geo.fmt); }
    // element i of the reference's correlation buffer (null + PRS)
    __device__ float2 corr_at(int i) const {
        if (uint32_t(i) < st.corr_explicit_len) return geo.corr_explicit[size_t(stream) * geo.null_period + i];
        return sample(st.corr_base + i);
    }

    // ---- CalculateL1Average (ofdm_demodulator.cpp:922-932) of `count` windows of K samples, window w starting at first + w * step,
    // into out[0 .. count).  A warp takes 8 windows at a time and walks them together, so that 8 independent loads are in flight
    // per lane: one window after the other costs a full memory latency per window (~1 us), which made a single unlocked stream
    // (FindNullPowerDip scans whole blocks) the long pole of a 1024-stream step.  Per window the additions keep their order
    // (lane-strided partial sums, then the butterfly): the frame kernel sums its windows the same way.  No barrier inside.
    __device__ void l1_windows(float* out, int64_t first, int step, int K, int count) {
        constexpr int U = 8;
        const int lane = tid & 31, warp = tid >> 5, n_warps = THREADS / 32;
        for (int w = warp * U; w < count; w += n_warps * U) {
            const int64_t base = first + int64_t(w) * step;
            float acc[U];
#pragma unroll
            for (int u = 0; u < U; u++) acc[u] = 0.0f;
            // four strides of 32 samples per round: with the default 100-sample windows every load of the 8 windows is issued
            // before the first addition (one memory round trip per 8 windows instead of four)
            for (int i0 = lane; i0 < K; i0 += 128) {
                float2 v[4][U];
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int u = 0; u < U; u++)
                        v[q][u] = (w + u < count && i0 + 32 * q < K) ? sample(base + int64_t(u) * step + i0 + 32 * q) : make_float2(0.0f, 0.0f);
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int u = 0; u < U; u++) acc[u] += fabsf(v[q][u].x) + fabsf(v[q][u].y);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                float a = acc[u];
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, d);
                if (lane == 0 && w + u < count) out[w + u] = a / float(K);
            }
        }
    }

    // ---- UpdateSignalAverage (ofdm_demodulator.cpp:934-950) of the current call, folded into l1_average in window order.  The
    // window averages come from the frame kernel where one of its items had the samples in shared memory (own_lo/own_hi, recorded
    // at dispatch; those launches have completed: a dispatch ends its control pass) and are summed here otherwise.  Runs once per
    // call: at the end of the stream's last active pass, or the moment FindNullPowerDip needs the average.
    __device__ void fold_average() {
        const int64_t N = st.call_end - st.call_begin;
        const int K = st.cfg.signal_l1_nb_samples;
        if (N >= K && K > 0) {
            const int64_t M = N - K;
            const int L = K * st.cfg.signal_l1_nb_decimate;
            const int64_t n_windows = (L > 0) ? (M + L - 1) / L : 0;
            const float* win = geo.l1_windows + size_t(stream) * geo.l1_windows_stride;
            for (int64_t w0 = 0; w0 < n_windows; w0 += CTRL_L1_BATCH) {
                const int count = int(min(int64_t(CTRL_L1_BATCH), n_windows - w0));
                int64_t cursor = w0;
                for (int r = 0; r < st.own_n && cursor < w0 + count; r++) {
                    const int64_t lo = max(int64_t(st.own_lo[r]), cursor), hi = min(int64_t(st.own_hi[r]), w0 + count - 1);
                    if (hi < lo) continue;
                    if (lo > cursor) l1_windows(l1buf + (cursor - w0), st.call_begin + cursor * L, L, K, int(lo - cursor));
                    for (int64_t w = lo + tid; w <= hi; w += THREADS) l1buf[w - w0] = win[w];
                    cursor = hi + 1;
                }
                if (cursor < w0 + count) l1_windows(l1buf + (cursor - w0), st.call_begin + cursor * L, L, K, int(w0 + count - cursor));
                __syncthreads();
                if (tid == 0) {
                    const float beta = st.cfg.signal_l1_update_beta;
                    float avg = st.l1_average;
                    for (int w = 0; w < count; w++) avg = beta * avg + (1.0f - beta) * l1buf[w];
                    st.l1_average = avg;
                }
                __syncthreads();
            }
        }
        if (tid == 0) {
            st.avg_pending = 0;
            // every frame of this call reports the average of the whole block, as the reference's Process() has it before the first
            // frame is dispatched
            for (int f = 0; f < st.frames_in_call; f++) geo.infos[size_t(stream) * geo.slots + f].signal_average = st.l1_average;
            st.pending_info.signal_average = st.l1_average;
        }
        __syncthreads();
    }

    // ---- FindNullPowerDip (ofdm_demodulator.cpp:291-347)
    __device__ void find_null_power_dip() {
        if (st.avg_pending) fold_average();  // the thresholds use the average over the whole block (UpdateSignalAverage runs first)
        const int64_t c0 = st.consumed;
        const int64_t N = st.call_end - c0;
        const int K = st.cfg.signal_l1_nb_samples;
        const int64_t M = N - K;
        const int64_t n_windows = (M > 0 && K > 0) ? (M + K - 1) / K : 0;
        __shared__ int64_t nb_read_s;
        if (tid == 0) nb_read_s = N;
        __syncthreads();
        for (int64_t w0 = 0; w0 < n_windows; w0 += CTRL_L1_BATCH) {
            const int count = int(min(int64_t(CTRL_L1_BATCH), n_windows - w0));
            l1_windows(l1buf, c0 + w0 * K, K, K, count);
            __syncthreads();
            if (tid == 0) {
                const float start_thresh = st.l1_average * st.cfg.null_l1_thresh_null_start;
                const float end_thresh = st.l1_average * st.cfg.null_l1_thresh_null_end;
                for (int w = 0; w < count; w++) {
                    const float l1 = l1buf[w];
                    if (st.null_start_found) {
                        if (l1 > end_thresh) {
                            st.null_end_found = 1;
                            nb_read_s = (w0 + w) * K + K;
                            break;
                        }
                    } else if (l1 < start_thresh) {
                        st.null_start_found = 1;
                    }
                }
            }
            __syncthreads();
            if (st.null_end_found) break;
        }
        // CircularBuffer::ConsumeBuffer(read_all = true) (circular_buffer.h:18-38): only the last `cap` samples survive
        const int64_t nb_read = nb_read_s;
        const uint32_t cap = uint32_t(geo.null_period);
        float2* ring = geo.ring + size_t(stream) * cap;
        const int64_t keep = min(nb_read, int64_t(cap));
        const uint32_t first_slot = uint32_t((int64_t(st.ring_index) + (nb_read - keep)) % cap);
        for (int64_t j = tid; j < keep; j += THREADS) ring[(first_slot + uint32_t(j)) % cap] = sample(c0 + (nb_read - keep) + j);
        __syncthreads();
        if (tid == 0) {
            st.ring_index = uint32_t((int64_t(st.ring_index) + nb_read) % cap);
            st.ring_length = uint32_t(min(int64_t(st.ring_length) + nb_read, int64_t(cap)));
            st.consumed = c0 + nb_read;
        }
        __syncthreads();
        if (!st.null_end_found) return;
        // copy the ring, oldest slot first, into the head of the correlation buffer (:333-338)
        const uint32_t L = st.ring_length, start = st.ring_index;
        float2* dst = geo.corr_explicit + size_t(stream) * cap;
        for (uint32_t i = tid; i < L; i += THREADS) dst[i] = ring[(i + start) % cap];
        __syncthreads();
        if (tid == 0) {
            st.null_start_found = 0;
            st.null_end_found = 0;
            st.corr_length = L;
            st.corr_explicit_len = L;
            st.corr_base = st.consumed - int64_t(L);
            st.ring_length = 0;
            st.state = DAB_OFDM_READING_NULL_AND_PRS;
        }
        __syncthreads();
    }

    // ---- OFDM_Demod::Reset (ofdm_demodulator.cpp:277-289)
    __device__ void reset_thread0() {
        st.state = DAB_OFDM_FINDING_NULL_POWER_DIP;
        st.corr_length = 0;
        st.corr_explicit_len = 0;
        st.total_frames_desync++;
        st.is_found_coarse = 0;
        st.freq_coarse = 0.0f;
        st.freq_fine = 0.0f;
        st.fine_time_offset = 0;
    }

    // ---- UpdateFineFrequencyOffset (ofdm_demodulator.cpp:829-840)
    __device__ void update_fine_thread0(float delta) {
        const float spacing = 1.0f / float(NFFT);
        const float wrap = 0.5f * spacing * 1.01f;
        st.freq_fine += delta;
        st.freq_fine = fmodf(st.freq_fine, wrap);
    }

    __device__ void store_natural(const float2 (&v)[16], bool conjugate) {
        if (tid < T) {
#pragma unroll
            for (int r = 0; r < 16; r++) nat[fft_out_bin<NFFT>(tid, r)] = conjugate ? cconj(v[r]) : v[r];
        }
        __syncthreads();
    }

    template <typename F>
    __device__ ArgMax block_argmax(int n, F value_at) {
        ArgMax best{-INFINITY, 0x7FFFFFFF};
        for (int i = tid; i < n; i += THREADS) best = argmax_combine(best, ArgMax{value_at(i), i});
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            ArgMax o;
            o.value = __shfl_xor_sync(0xFFFFFFFFu, best.value, d);
            o.index = __shfl_xor_sync(0xFFFFFFFFu, best.index, d);
            best = argmax_combine(best, o);
        }
        if ((tid & 31) == 0) red[tid >> 5] = make_float2(best.value, __int_as_float(best.index));
        __syncthreads();
        ArgMax r{red[0].x, __float_as_int(red[0].y)};
        for (int w = 1; w < THREADS / 32; w++) r = argmax_combine(r, ArgMax{red[w].x, __float_as_int(red[w].y)});
        __syncthreads();
        return r;
    }

    // ---- RunCoarseFreqSync (ofdm_demodulator.cpp:360-471) followed by RunFineTimeSync (:473-548).
    // The reference runs the two back to back inside one Process() call (neither consumes samples, :256-262), so they are one
    // five-transform sequence here:  stage 0-2 = coarse frequency (FFT, IFFT, FFT), stage 3-4 = fine time (FFT, IFFT).  The
    // transform is inlined ONCE, inside the stage loop, with its 16 points per thread in registers: five inlined copies made
    // the kernel wait on instruction fetches for a third of its cycles (profiles/r01d), and one out-of-line copy sent the
    // points through local memory, which thrashes the 22 KB of L1 left beside four CTAs' shared memory (profiles/r01f:
    // 176 us per pass against 80 us).
    __device__ void run_sync() {
        int stage = 3;
        if (st.state == DAB_OFDM_RUNNING_COARSE_FREQ_SYNC) {
            if (st.cfg.sync_is_coarse_freq_correction) {
                stage = 0;
            } else {
                if (tid == 0) { st.freq_coarse = 0.0f; st.state = DAB_OFDM_RUNNING_FINE_TIME_SYNC; }
                __syncthreads();
            }
        }
        const int cp = geo.cyclic_prefix, sp = geo.symbol_period, np = geo.null_period;
        float2 v[16];
#pragma unroll 1
        for (; stage < 5; stage++) {
            // ---- input of the transform, v[n1] = x[n1 T + t]
            if (tid < T) {
                if (stage == 0) {
                    // FFT of the first NFFT samples of the received PRS symbol window (:383-387)
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) v[n1] = corr_at(np + n1 * T + tid);
                } else if (stage == 1) {
                    // relative phase conj(X[i]) X[i+1], last bin zero (:901-909); IFFT = conj(FFT(conj(.)))
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) {
                        const int i = n1 * T + tid;
                        const float2 rel = (i < NFFT - 1) ? cmul(cconj(nat[i]), nat[i + 1]) : make_float2(0.0f, 0.0f);
                        v[n1] = cconj(rel);
                    }
                } else if (stage == 2) {
                    // multiply by conj(IFFT(relative phase of the reference)), then FFT (:402-414)
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) {
                        const int i = n1 * T + tid;
                        v[n1] = cmul(nat[i], geo.prs_time_ref_conj[i]);
                    }
                } else if (stage == 3) {
                    // fine time: PRS window with the PLL at the (just updated) net frequency offset (:482-489)
                    const PllSymbol pll = pll_symbol(st.freq_coarse + st.freq_fine, 0, NFFT);
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) {
                        const int i = n1 * T + tid;
                        v[n1] = pll_rotate(pll, corr_at(np + i), i);
                    }
                } else {
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) v[n1] = nat[n1 * T + tid];
                }
            }
            __syncthreads();  // `nat` aliases exchange 1
            if (tid < T) fft_pass1<NFFT>(v, tid, e1, tw1);
            __syncthreads();
            if (tid < T) fft_pass2<NFFT>(v, tid, e1, e2, tw2);
            __syncthreads();
            if (tid < T) fft_pass3<NFFT>(v, tid, e2);
            // ---- what becomes of the spectrum (bin of register slot r = fft_out_bin(t, r))
            if (stage == 0) {
                store_natural(v, false);
            } else if (stage == 1) {
                store_natural(v, true);
            } else if (stage == 2) {
                coarse_freq_decide(v);
            } else if (stage == 3) {
                // correlation in time = conjugate product in frequency; then IFFT through conj(FFT(conj(.)))
                if (tid < T) {
#pragma unroll
                    for (int r = 0; r < 16; r++) v[r] = cmul(v[r], geo.prs_fft_ref_conj[fft_out_bin<NFFT>(tid, r)]);
                }
                store_natural(v, true);
            } else {
                fine_time_decide(v, cp, sp, np);
            }
        }
    }

    // steps 6-11 of RunCoarseFreqSync (:416-470) on the spectrum of the third transform
    __device__ __forceinline__ void coarse_freq_decide(const float2 (&v)[16]) {
        // step 6: magnitude in dB, fft-shifted (:911-920)
        float* resp = geo.freq_response + size_t(stream) * NFFT;
        float* mag = reinterpret_cast<float*>(nat);
        __syncthreads();  // every thread is past its exchange-2 reads before `nat` (exchange 1) is overwritten
        if (tid < T) {
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int k = fft_out_bin<NFFT>(tid, r);
                const float m = 20.0f * log10f(sqrtf(v[r].x * v[r].x + v[r].y * v[r].y));
                const int i = (k + NFFT / 2) % NFFT;
                mag[i] = m;
                resp[i] = m;
            }
        }
        __syncthreads();
        // step 7: peak inside the allowed window, first maximum wins
        const int Mh = NFFT / 2;
        int max_off = int(st.cfg.sync_max_coarse_freq_correction_norm * float(NFFT));
        max_off = max(0, min(max_off, Mh));
        const int lo = -max_off + Mh;
        const int hi = min(max_off + Mh, NFFT - 1);  // fft_index == NFFT is skipped by the reference
        const ArgMax peak = block_argmax(hi - lo + 1, [&](int i) { return mag[lo + i]; });
        if (tid == 0) {
            const int max_index = peak.index + lo - Mh;
            // step 8: magnitude-weighted centroid of the three bins around the peak
            float pk_mag[3];
            int pk_idx[3];
            for (int k = 0; k < 3; k++) {
                int index = max_index - 1 + k;
                index = max(-max_off, min(index, max_off));
                int fi = index + Mh;
                if (fi >= NFFT) fi = NFFT - 1;
                pk_mag[k] = powf(10.0f, mag[fi] / 20.0f);
                pk_idx[k] = fi - Mh;
            }
            float peak_sum = 0.0f, lerp = 0.0f;
            for (int k = 0; k < 3; k++) peak_sum += pk_mag[k];
            for (int k = 0; k < 3; k++) lerp += float(pk_idx[k]) * pk_mag[k] / peak_sum;
            const float predicted = -lerp / float(NFFT);
            const float error = predicted - st.freq_coarse;
            // steps 9-11: fast / slow update and the counter-adjustment of the fine offset
            const float large_thresh = 1.5f / float(NFFT);
            const bool fast = (fabsf(error) > large_thresh) || !st.is_found_coarse;
            const float beta = fast ? 1.0f : st.cfg.sync_coarse_freq_slow_beta;
            const float delta = beta * error;
            st.freq_coarse += delta;
            st.is_found_coarse = 1;
            update_fine_thread0(-delta);
            st.state = DAB_OFDM_RUNNING_FINE_TIME_SYNC;
        }
        __syncthreads();
    }

    // the impulse-response half of RunFineTimeSync (:497-547) on the spectrum of the fifth transform
    __device__ __forceinline__ void fine_time_decide(const float2 (&v)[16], int cp, int sp, int np) {
        float* resp = geo.impulse_response + size_t(stream) * NFFT;
        float* imp = reinterpret_cast<float*>(nat);
        float partial = 0.0f;
        __syncthreads();  // see coarse_freq_decide
        if (tid < T) {
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int k = fft_out_bin<NFFT>(tid, r);
                const float a = 20.0f * log10f(sqrtf(v[r].x * v[r].x + v[r].y * v[r].y));
                imp[k] = a;
                resp[k] = a;
                partial += a;
            }
        }
        // mean of the impulse response
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) partial += __shfl_xor_sync(0xFFFFFFFFu, partial, d);
        if ((tid & 31) == 0) red[16 + (tid >> 5)] = make_float2(partial, 0.0f);
        __syncthreads();
        // peak weighted by its distance from the expected position (start value = unweighted sample 0, :503)
        const float decay = 1.0f - st.cfg.sync_impulse_peak_distance_probability;
        const ArgMax peak = block_argmax(NFFT, [&](int i) {
            const float norm_dist = float(abs(cp - i)) / float(sp);
            return (1.0f - decay * norm_dist) * imp[i];
        });
        if (tid == 0) {
            float avg = 0.0f;
            for (int w = 0; w < THREADS / 32; w++) avg += red[16 + w].x;
            avg /= float(NFFT);
            float max_value = imp[0];
            int max_index = 0;
            if (peak.value > max_value) { max_value = peak.value; max_index = peak.index; }
            if ((max_value - avg) < st.cfg.sync_impulse_peak_threshold_db) {
                reset_thread0();
            } else {
                const int offset = max_index - cp;
                st.frame_start = st.corr_base + np + offset;
                st.corr_length = 0;
                st.fine_time_offset = offset;
                st.state = DAB_OFDM_READING_SYMBOLS;
                // from here to the end of the frame nothing changes the frequency offsets (the previous frame's fine update came
                // before this synchronisation): the frame kernel may start on the symbols as they arrive
                st.frame_freq = st.freq_coarse + st.freq_fine;
                st.symbols_done = 0;
                st.frame_slot = choose_slot_thread0();
                st.syncs_in_call++;
                st.pending_info.frame_start = st.frame_start;
                st.pending_info.fine_time_offset = offset;
                st.pending_info.coarse_offset = st.freq_coarse;
                st.pending_info.fine_offset_used = st.freq_fine;
            }
        }
        __syncthreads();
    }

    // Soft-bit ring slot for the frame whose PRS was just synchronised.  The j-th synchronisation of call c takes slot (c + j) mod R
    // unless a frame completed earlier in this call still sits there: streams that complete one frame per call (the batched steady
    // state) then all use the same slot in the same call, and their frames leave the device as one strided copy.
    __device__ int choose_slot_thread0() {
        const int R = geo.ring_slots;
        int cand = int((geo.call_index + uint32_t(st.syncs_in_call)) % uint32_t(R));
        for (int k = 0; k < R; k++) {
            bool busy = false;
            for (int f = 0; f < st.frames_in_call; f++) busy = busy || (geo.frame_slots[size_t(stream) * geo.slots + f] == cand);
            if (!busy) break;
            cand = (cand + 1) % R;
        }
        return cand;
    }

    // ---- the fine-frequency half of CoordinatorThread (ofdm_demodulator.cpp:606-618, 632) for the frame that just completed
    __device__ void finish_pipeline() {
        // the per-symbol phase errors arrive with one parallel load; the sum keeps the reference's symbol order
        const float* pe = geo.phase_err + size_t(stream) * geo.n_symbols;
        for (int s = tid; s < geo.n_symbols; s += THREADS) l1buf[s] = pe[s];
        __syncthreads();
        if (tid == 0) {
            float total = 0.0f;
            for (int s = 0; s < geo.n_symbols; s++) total += l1buf[s];
            const float avg = total / float(geo.n_symbols);
            const float two_pi = 3.14159265358979323846f * 2.0f;
            const float fine_error = (1.0f / float(NFFT)) * avg / two_pi;  // CalculateFineFrequencyError :821-823
            update_fine_thread0(-st.cfg.sync_fine_freq_update_beta * fine_error);
            st.pending_info.fine_offset_after = st.freq_fine;
            st.total_frames_read++;
            geo.infos[size_t(stream) * geo.slots + (st.frames_in_call - 1)] = st.pending_info;
            st.pipeline_pending = 0;
        }
        __syncthreads();
    }

    // Hands symbols [st.symbols_done, s_end) of the frame being received to the frame kernel launched after this pass, split into
    // up to n_chunks work items, and lets those items own the UpdateSignalAverage windows inside the samples they read.
    __device__ void dispatch_thread0(int pass, int s_end) {
        const int s_lo = st.symbols_done;
        const int n_new = s_end - s_lo;
        const int SP = geo.symbol_period;
        int parts = (n_new + geo.syms_per_chunk - 1) / geo.syms_per_chunk;
        parts = max(1, min(parts, geo.n_chunks));
        // window ownership: windows [w_lo, w_hi] of this call lie inside the samples the items load
        const int K = st.cfg.signal_l1_nb_samples;
        const int L = K * st.cfg.signal_l1_nb_decimate;
        const int64_t N = st.call_end - st.call_begin;
        int64_t n_windows = 0;
        if (K > 0 && L > 0 && N >= K) n_windows = (N - K + L - 1) / L;
        // the frame kernel sums at most l1_per_symbol windows per symbol (two per sub-warp) of at most L1_PREFIX + 1 samples
        const bool can_own = geo.frame_owns_l1 && st.avg_pending && n_windows > 0 && K - 1 <= L1_PREFIX && st.own_n < CTRL_MAX_OWNED &&
                             n_windows <= int64_t(geo.l1_windows_stride) && (SP + K - 1) / L + 1 <= geo.l1_per_symbol;
        int64_t own_lo = -1, own_hi = -2;
        FrameDesc d;
        d.src = src;
        d.mask = geo.mask;
        d.limit = geo.limit;
        d.start = st.frame_start;
        d.freq = st.frame_freq;
        d.valid = 1;
        d.bits = geo.bits + (size_t(stream) * geo.ring_slots + st.frame_slot) * geo.frame_bits;
        d.phase_err = geo.phase_err + size_t(stream) * geo.n_symbols;
        d.fft_tap = geo.fft_tap ? geo.fft_tap + size_t(stream) * geo.n_symbols * NFFT : nullptr;
        d.vec_tap = geo.vec_tap ? geo.vec_tap + size_t(stream) * (geo.n_symbols - 1) * geo.n_carriers : nullptr;
        d.l1_origin = st.call_begin;
        d.l1_k = K;
        d.l1_step = L;
        FrameDesc* out = geo.descs + (size_t(pass) * geo.n_streams + stream) * geo.n_chunks;
        int b = s_lo;
        for (int c = 0; c < parts; c++) {
            const int e = b + n_new / parts + (c < n_new % parts ? 1 : 0);
            d.s_begin = b;
            d.s_end = e;
            d.l1_out = nullptr;
            d.l1_w_lo = 0;
            d.l1_w_hi = -1;
            if (can_own) {
                // first item of the dispatch: windows that begin inside its first loaded symbol (and inside this call); later items:
                // windows that end after their reference symbol (the previous item takes those ending in it)
                const int64_t first_loaded = st.frame_start + int64_t(max(b - 1, 0)) * SP;
                const int64_t lo_abs = (c == 0) ? max(first_loaded, st.call_begin) : st.frame_start + int64_t(b) * SP - K + 1;
                const int64_t hi_abs = st.frame_start + int64_t(e) * SP - K;   // last admissible window start
                const int64_t rel_lo = max(lo_abs - st.call_begin, int64_t(0));
                const int64_t w_lo = (rel_lo + L - 1) / L;
                int64_t w_hi = (hi_abs >= st.call_begin) ? (hi_abs - st.call_begin) / L : -1;
                w_hi = min(w_hi, n_windows - 1);
                if (w_hi >= w_lo) {
                    d.l1_out = geo.l1_windows + size_t(stream) * geo.l1_windows_stride;
                    d.l1_w_lo = int(w_lo);
                    d.l1_w_hi = int(w_hi);
                    if (own_lo < 0) own_lo = w_lo;
                    own_hi = w_hi;
                }
            }
            out[c] = d;
            b = e;
        }
        if (own_hi >= own_lo && own_lo >= 0) {
            st.own_lo[st.own_n] = int(own_lo);
            st.own_hi[st.own_n] = int(own_hi);
            st.own_n++;
        }
        st.symbols_done = s_end;
    }

    // ---- ReadSymbols (ofdm_demodulator.cpp:550-577): returns true when the pass has to stop for the frame kernel
    __device__ bool read_symbols_thread0(int pass, bool& dispatched) {
        const int S = geo.n_symbols;
        const int64_t frame_cap = int64_t(S) * geo.symbol_period + geo.null_period;
        const int64_t frame_end = st.frame_start + frame_cap;
        const int64_t take = min(st.call_end - st.consumed, frame_end - st.consumed);
        st.consumed += take;
        const bool may_dispatch = pass < geo.frame_passes;
        if (st.consumed != frame_end) {
            // the block ends inside the frame: the symbols that are complete can be demodulated now
            if (geo.eager && may_dispatch) {
                const int avail = int(min(int64_t(S), (st.consumed - st.frame_start) / geo.symbol_period));
                if (avail > st.symbols_done) {
                    dispatch_thread0(pass, avail);
                    dispatched = true;
                }
            }
            return false;
        }
        // the trailing NULL symbol becomes the head of the next correlation buffer (:557-562)
        st.corr_base = frame_end - geo.null_period;
        st.corr_length = uint32_t(geo.null_period);
        st.corr_explicit_len = 0;
        // the frame is complete (SignalStart, :572): whatever the frame kernel has not seen yet goes out now
        const bool need_kernel = st.symbols_done < S;
        if (need_kernel) {
            dispatch_thread0(pass, S);
            dispatched = true;
        }
        const int f = st.frames_in_call;
        geo.frame_slots[size_t(stream) * geo.slots + f] = st.frame_slot;
        st.pending_info.signal_average = st.l1_average;
        st.pending_info.total_desync = st.total_frames_desync;
        st.pending_info.slot = st.frame_slot;
        st.pending_slot = st.frame_slot;
        st.pipeline_pending = 1;
        st.frames_in_call = f + 1;
        st.state = DAB_OFDM_READING_NULL_AND_PRS;
        return need_kernel;
    }
};

// pass: index of this control pass within the call (0 = first; it also opens the call: OFDM_Demod::Process's entry,
// ofdm_demodulator.cpp:235-243).  Passes below geo.frame_passes are followed by a frame-kernel launch over the items they wrote.
template <int NFFT, int SB>
__global__ void __launch_bounds__(ControlSmem<NFFT>::THREADS, 4)
ofdm_control_kernel(ControlGeom geo, int pass) {
    using G = FftGeom<NFFT>;
    using C = Control<NFFT, SB>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ StreamState st;
    __shared__ int stop_flag, dispatched_flag;
    const int stream = geo.stream0 + blockIdx.x, tid = threadIdx.x;

    float2* tw1 = reinterpret_cast<float2*>(smem_raw);
    float2* tw2 = tw1 + G::TW1_SIZE;
    float2* e1 = tw2 + G::TW2_SIZE;
    float2* e2 = e1 + G::E1_SIZE;
    float2* nat = e1;
    float* l1buf = reinterpret_cast<float*>(e1);
    float2* red = e2 + G::E2_SIZE;

    if (tid == 0) {
        st = geo.states[stream];
        stop_flag = 0;
        dispatched_flag = 0;
        if (pass == 0) {
            const uint64_t n = geo.n_per_stream ? geo.n_per_stream[stream] : geo.n_uniform;
            st.call_begin = st.call_end;
            st.call_end = st.call_begin + int64_t(n);
            st.avg_pending = (n > 0) ? 1 : 0;
            st.frames_in_call = 0;
            st.syncs_in_call = 0;
            st.own_n = 0;
        }
    }
    // the work items this pass may fill start out invalid
    if (pass < geo.frame_passes && tid < geo.n_chunks) geo.descs[(size_t(pass) * geo.n_streams + stream) * geo.n_chunks + tid].valid = 0;
    __syncthreads();
    // nothing to do: no frame waiting for its fine update, no unread samples, the signal average of the call is folded
    if (!st.pipeline_pending && st.consumed >= st.call_end && !st.avg_pending) {
        if (pass == 0 && tid == 0) {
            geo.states[stream] = st;
            geo.frames_in_call[stream] = 0;
        }
        return;
    }

    __shared__ int tw_loaded;  // the twiddle tables are only staged when a synchronisation stage actually runs
    if (tid == 0) tw_loaded = 0;
    const void* src = reinterpret_cast<const unsigned char*>(geo.samples) + size_t(stream) * geo.stream_stride * size_t(SB);
    C ctl{geo, st, stream, tid, tw1, tw2, e1, e2, nat, l1buf, red, src};

    if (st.pipeline_pending) ctl.finish_pipeline();
    __syncthreads();

    // OFDM_Demod::Process main loop (ofdm_demodulator.cpp:245-274)
    while (st.consumed < st.call_end && !stop_flag) {
        switch (st.state) {
        case DAB_OFDM_FINDING_NULL_POWER_DIP:
            ctl.find_null_power_dip();
            break;
        case DAB_OFDM_READING_NULL_AND_PRS:  // ReadNullPRS :349-358
            if (tid == 0) {
                const int64_t cap = int64_t(geo.null_period) + geo.symbol_period;
                const int64_t take = min(st.call_end - st.consumed, cap - int64_t(st.corr_length));
                st.corr_length += uint32_t(take);
                st.consumed += take;
                if (int64_t(st.corr_length) == cap) st.state = DAB_OFDM_RUNNING_COARSE_FREQ_SYNC;
            }
            __syncthreads();
            break;
        case DAB_OFDM_RUNNING_COARSE_FREQ_SYNC:
        case DAB_OFDM_RUNNING_FINE_TIME_SYNC:
            if (!tw_loaded) {
                fft_load_twiddles<NFFT>(tw1, geo.twiddles, tid, C::THREADS);
                __syncthreads();
                if (tid == 0) tw_loaded = 1;
                __syncthreads();
            }
            ctl.run_sync();
            break;
        case DAB_OFDM_READING_SYMBOLS:
            if (tid == 0) {
                bool dispatched = false;
                if (ctl.read_symbols_thread0(pass, dispatched)) stop_flag = 1;  // wait for the frame kernel before touching the next PRS
                if (dispatched) dispatched_flag = 1;
            }
            __syncthreads();
            // a frame whose symbols were all demodulated as they arrived needs no kernel: its fine-frequency update runs right here
            if (st.pipeline_pending && !stop_flag) ctl.finish_pipeline();
            break;
        }
    }
    __syncthreads();
    // the call is over for this stream and every item it dispatched has run: fold UpdateSignalAverage
    if (st.avg_pending && st.consumed >= st.call_end && !dispatched_flag && !st.pipeline_pending) ctl.fold_average();
    if (tid == 0) {
        geo.states[stream] = st;
        geo.frames_in_call[stream] = st.frames_in_call;
    }
}

}  // namespace dabb200
