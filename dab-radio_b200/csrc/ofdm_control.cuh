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
// UpdateSignalAverage (ofdm_demodulator.cpp:934-950) runs first in the reference's Process(); here its value is only needed when a
// stream searches for the NULL symbol or when the call is over, so it is folded lazily: the pass in which a stream finishes its
// block sums the block's window averages and folds them into the running average in window order (a searching stream does it the
// moment FindNullPowerDip needs the thresholds; from pass 1 on every other CTA of an SM folds before its synchronisation transforms
// so that the memory-bound and the issue-bound phases of co-resident CTAs overlap).  Two other homes for the window sums were built and measured -- inside the frame
// kernel while the samples sit in shared memory, and a kernel of their own on a side stream beside the frame kernel -- and dropped:
// profiles/r02_step_probes.md.
#pragma once
#include "ofdm_device.cuh"
#include "ofdm_frame.cuh"
#include "ofdm_frame_v3.cuh"

namespace dabb200 {

constexpr int FRAME_MAX_CHUNKS = 16;  // work items a frame is split into (load balance: a frame is 76 - 153 symbols long)

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
    // --- cursor / current Process() call
    int64_t consumed;        // absolute index of the next unread sample
    int64_t call_begin;
    int64_t call_end;
    int32_t avg_pending;         // UpdateSignalAverage of this call not folded into l1_average yet (see fold_average)
    int32_t pipeline_pending;    // a frame was dispatched; its phase errors still have to update the fine offset
    int32_t pending_slot;        // output slot of the most recently completed frame
    int32_t frames_in_call;      // frames completed during this call
    dab_ofdm_config cfg;
    dab_ofdm_frame_info pending_info;
};

struct ControlGeom {
    int n_symbols, symbol_period, null_period, cyclic_prefix, n_carriers;
    int slots;               // frames a stream can complete per call = soft-bit buffers per stream
    int n_streams;           // streams of the handle (row length of descs)
    int stream0;             // first stream this launch covers (launches are split into pipeline ways, see run_call)
    int n_chunks;            // work items per frame (<= FRAME_MAX_CHUNKS)
    int frame_passes;        // control passes of this call that are followed by a frame-kernel launch
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
    int8_t* bits;            // [n_streams][slots][frame_bits]
    float* phase_err;        // [n_streams][n_symbols]
    const float2* twiddles;  // precomputed FFT twiddles (fft_twiddle_init_kernel)
    float2* fft_tap;         // optional [n_streams][n_symbols * NFFT]
    float2* vec_tap;         // optional [n_streams][(n_symbols-1) * n_carriers]
    const uint64_t* n_per_stream;  // samples of this call per stream, nullptr: n_uniform
    uint64_t n_uniform;
    int n_sm;                      // SMs of the device: CTAs b and b + n_sm share an SM in the first wave (phase staggering)
};

constexpr int CTRL_L1_BATCH = 1024;  // windows whose L1 averages are computed in parallel before they are folded / searched in order

template <int NFFT>
struct ControlSmem {
    using G = FftGeom<NFFT>;
    // four warps whatever the FFT size: the transforms use the first N/16 threads, the L1 window scans (FindNullPowerDip walks whole
    // blocks while a stream is unlocked) use all of them.  With N/16 threads (one warp in Modes II / III) a 1024-stream step spent
    // 0.5 ms in the scans of the unlocked streams; 64 threads (twice the CTAs per SM for the synchronisation transforms) measured
    // 20 - 30 % slower than 128 on every Mode II / III configuration (profiles/r02_modes.md)
    static constexpr int THREADS = 128;
    // `nat` (natural-order spectrum between two transforms) aliases exchange 1 and the L1 window batch aliases the (contiguous) exchanges: both
    // are only alive while no transform is in flight (every hand-over is separated by a CTA barrier)
    static_assert(G::E1_SIZE >= NFFT, "nat must fit into exchange 1");
    static_assert(size_t(G::E1_SIZE + G::E2_SIZE) * sizeof(float2) >= size_t(CTRL_L1_BATCH) * sizeof(float) && CTRL_L1_BATCH >= 2 * THREADS, "L1 batch must fit into the exchanges");
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
    // into out[0 .. count).  One window after the other costs a full memory latency per window (~1 us), which made a single unlocked
    // stream (FindNullPowerDip scans whole blocks) the long pole of a 1024-stream step, and one lane per sample costs ~70 warp
    // instructions per 100-sample window (the fold of a Mode I frame's 393 windows was half of the control pass that ran it:
    // profiles/r02_control_ncu.md).  No barrier inside.
    __device__ void l1_windows(float* out, int64_t first, int step, int K, int count) {
        // the fold is DRAM-latency bound (what counts is bytes in flight per round trip) and, for short windows, bound by the
        // sectors per load instruction: a group of LPW lanes shares a window, each lane reads two samples per 16-byte load, so a
        // request covers whole 32-byte sectors exactly once.  100-sample windows: 8 lanes x 7 loads, 4 windows per warp and
        // round; windows of up to 32 samples (25 x 2 is the setting under which Modes II / III lock on whole-frame blocks):
        // 2 lanes x up to 8 loads, 2 x 16 windows per warp and round -- one 8-byte load per sample and one window per thread made
        // that fold 1.7 ms of a 2.3 ms Mode I step (32 sectors per request, every sector requested four times).
        if (K <= 32) l1_windows_t<2, 2>(out, first, step, K, count);
        else l1_windows_t<8, 1>(out, first, step, K, count);
    }

    // LPW lanes per window, U windows per lane group and round.  The odd sample in front of / behind the aligned pairs is read by
    // lane 0 / lane 1 of the group.  A window that wraps around the stream's ring (one per revolution) takes masked single-sample
    // loads instead.
    template <int LPW, int U>
    __device__ __forceinline__ void l1_windows_t(float* out, int64_t first, int step, int K, int count) {
        constexpr int Q = 8;            // pair loads per lane and window in one batch (Q * LPW pairs)
        constexpr int WPR = 32 / LPW;   // windows per warp and round (x U)
        const int lane = tid & 31, warp = tid >> 5, n_warps = THREADS / 32;
        const int sub = lane & (LPW - 1), grp = lane / LPW;
        const unsigned char* base_ptr = reinterpret_cast<const unsigned char*>(src);
        const uint32_t base_odd = uint32_t(reinterpret_cast<uintptr_t>(base_ptr) / uint32_t(SB)) & 1u;
        for (int w = warp * WPR * U; w < count; w += n_warps * WPR * U) {
            float acc[U];
            float2 edge[U];
            const unsigned char* wp[U];
            int j_first[U], j_end[U];   // aligned pairs j_first, + LPW, ... < j_end of window u are read by this lane
            bool slow[U];
            uint64_t p0[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int ww = w + WPR * u + grp;
                acc[u] = 0.0f;
                edge[u] = make_float2(0.0f, 0.0f);
                j_first[u] = j_end[u] = 0;
                slow[u] = false;
                wp[u] = base_ptr;
                p0[u] = uint64_t(first + int64_t(ww) * step) & geo.mask;
                if (ww >= count) continue;
                if (p0[u] + uint64_t(K) - 1 > geo.mask) { slow[u] = true; continue; }
                const int o = int((uint32_t(p0[u]) ^ base_odd) & 1u);   // 1: the window starts on the second sample of an aligned pair
                wp[u] = base_ptr + (p0[u] - uint64_t(o)) * uint64_t(SB);   // pair j = samples 2j - o, 2j - o + 1 of the window
                j_first[u] = o + sub;
                j_end[u] = (K + o) / 2;
                // requested here, added after the pair loads are on their way (one memory round trip per round, not two)
                if (sub == 0 && o) edge[u] = load_sample_ptr<SB>(wp[u] + SB, geo.fmt);
                if (sub == 1 && ((K + o) & 1)) edge[u] = load_sample_ptr<SB>(wp[u] + size_t(K - 1 + o) * SB, geo.fmt);
            }
            for (int jj = 0; jj < (K + 1) / 2; jj += LPW * Q) {
                float2 a[U][Q], b[U][Q];
#pragma unroll
                for (int u = 0; u < U; u++)
#pragma unroll
                    for (int q = 0; q < Q; q++) {
                        a[u][q] = b[u][q] = make_float2(0.0f, 0.0f);
                        const int j = j_first[u] + jj + LPW * q;
                        if (j < j_end[u]) load_sample_pair<SB>(wp[u] + size_t(j) * (2 * SB), geo.fmt, a[u][q], b[u][q]);
                    }
#pragma unroll
                for (int u = 0; u < U; u++)
#pragma unroll
                    for (int q = 0; q < Q; q++) acc[u] += (fabsf(a[u][q].x) + fabsf(a[u][q].y)) + (fabsf(b[u][q].x) + fabsf(b[u][q].y));
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                acc[u] += fabsf(edge[u].x) + fabsf(edge[u].y);
                if (slow[u]) {
                    for (int i = sub; i < K; i += LPW) {
                        const float2 x = sample(int64_t(p0[u]) + i);
                        acc[u] += fabsf(x.x) + fabsf(x.y);
                    }
                }
#pragma unroll
                for (int d = LPW / 2; d >= 1; d >>= 1) acc[u] += __shfl_xor_sync(0xFFFFFFFFu, acc[u], d);
                const int ww = w + WPR * u + grp;
                if (sub == 0 && ww < count) out[ww] = acc[u] / float(K);
            }
        }
    }

    // ---- UpdateSignalAverage (ofdm_demodulator.cpp:934-950) of the current call, folded into l1_average in window order.  Runs
    // once per call: in the pass in which the stream finishes its block, or the moment FindNullPowerDip needs the average.
    __device__ void fold_average() {
        const int64_t N = st.call_end - st.call_begin;
        const int K = st.cfg.signal_l1_nb_samples;
        if (N >= K && K > 0) {
            const int64_t M = N - K;
            const int L = K * st.cfg.signal_l1_nb_decimate;
            const int64_t n_windows = (L > 0) ? (M + L - 1) / L : 0;
            for (int64_t w0 = 0; w0 < n_windows; w0 += CTRL_L1_BATCH) {
                const int count = int(min(int64_t(CTRL_L1_BATCH), n_windows - w0));
                l1_windows(l1buf, st.call_begin + w0 * L, L, K, count);
                __syncthreads();
                // avg <- beta avg + (1 - beta) x_w over the windows in order (:941-947).  The recurrence is linear: a thread runs it over
                // its own run of m consecutive windows starting from zero, r_t, and the runs are chained as avg <- beta^m avg + r_t.
                // Same value up to float rounding (the signal average is compared at 1e-4), and thread 0 walks 128 partial
                // results instead of up to 1024 windows while the other 127 threads wait.
                const float beta = st.cfg.signal_l1_update_beta;
                const int m = (count + THREADS - 1) / THREADS;
                float* part = l1buf;   // (r_t, beta^len_t) pairs, written over the window averages once every thread has read its run
                {
                    float r = 0.0f, bp = 1.0f;
                    const int lo = min(tid * m, count), hi = min(lo + m, count);
                    for (int w = lo; w < hi; w++) {
                        r = beta * r + (1.0f - beta) * l1buf[w];
                        bp *= beta;
                    }
                    __syncthreads();
                    part[2 * tid] = r;
                    part[2 * tid + 1] = bp;
                }
                __syncthreads();
                if (tid == 0) {
                    float avg = st.l1_average;
                    for (int t = 0; t < THREADS; t++) avg = part[2 * t + 1] * avg + part[2 * t];
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
            // The reference walks the windows in order: the first one below the start threshold arms the detector, the first one
            // after it above the end threshold ends the NULL symbol (:306-326).  Both are "first index where" searches: two block
            // reductions instead of one thread walking up to 1024 windows while the other 127 wait (that walk was 39 % of the stall
            // samples of a control pass with unlocked streams in it, and unlocked streams are its long pole).
            const float start_thresh = st.l1_average * st.cfg.null_l1_thresh_null_start;
            const float end_thresh = st.l1_average * st.cfg.null_l1_thresh_null_end;
            bool started = st.null_start_found != 0;
            int from = 0;
            if (!started) {
                const int a = block_first(count, [&](int i) { return l1buf[i] < start_thresh; });
                if (a != 0x7FFFFFFF) {
                    started = true;
                    from = a + 1;
                }
            }
            int b = 0x7FFFFFFF;
            if (started) b = block_first(count - from, [&](int i) { return l1buf[from + i] > end_thresh; });
            if (tid == 0) {
                if (started) st.null_start_found = 1;
                if (b != 0x7FFFFFFF) {
                    st.null_end_found = 1;
                    nb_read_s = (w0 + from + b) * K + K;
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

    // smallest i < n with pred(i), 0x7FFFFFFF if there is none; the same value in every thread
    template <typename P>
    __device__ int block_first(int n, P pred) {
        int best = 0x7FFFFFFF;
        for (int i = tid; i < n; i += THREADS)
            if (pred(i)) {
                best = i;
                break;
            }
        best = __reduce_min_sync(0xFFFFFFFFu, best);
        if ((tid & 31) == 0) red[tid >> 5] = make_float2(__int_as_float(best), 0.0f);
        __syncthreads();
        int r = __float_as_int(red[0].x);
        for (int w = 1; w < THREADS / 32; w++) r = min(r, __float_as_int(red[w].x));
        __syncthreads();
        return r;
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
    __device__ void run_sync(int state) {
        int stage = 3;
        if (state == DAB_OFDM_RUNNING_COARSE_FREQ_SYNC) {
            const bool coarse = st.cfg.sync_is_coarse_freq_correction != 0;
            __syncthreads();   // everyone has read the configuration before thread 0 touches the state
            if (coarse) {
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
                st.pending_info.frame_start = st.frame_start;
                st.pending_info.fine_time_offset = offset;
            }
        }
        __syncthreads();
    }

    // ---- the fine-frequency half of CoordinatorThread (ofdm_demodulator.cpp:606-618, 632) for the frame dispatched before
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
            geo.infos[size_t(stream) * geo.slots + st.pending_slot] = st.pending_info;
            st.pipeline_pending = 0;
        }
        __syncthreads();
    }

    // ---- ReadSymbols (ofdm_demodulator.cpp:550-577): returns true when a frame was dispatched
    __device__ bool read_symbols_thread0(int pass) {
        const int S = geo.n_symbols;
        const int64_t frame_cap = int64_t(S) * geo.symbol_period + geo.null_period;
        const int64_t frame_end = st.frame_start + frame_cap;
        const int64_t take = min(st.call_end - st.consumed, frame_end - st.consumed);
        st.consumed += take;
        if (st.consumed != frame_end) return false;
        // the trailing NULL symbol becomes the head of the next correlation buffer (:557-562)
        st.corr_base = frame_end - geo.null_period;
        st.corr_length = uint32_t(geo.null_period);
        st.corr_explicit_len = 0;
        // hand the frame to the frame kernel launched after this pass (SignalStart, :572), n_chunks work items of about equal length
        const int slot = st.frames_in_call;
        FrameDesc d;
        d.src = src;
        d.mask = geo.mask;
        d.limit = geo.limit;
        d.start = st.frame_start;
        d.freq = st.freq_coarse + st.freq_fine;
        d.valid = 1;
        d.bits = geo.bits + (size_t(stream) * geo.slots + slot) * geo.frame_bits;
        d.phase_err = geo.phase_err + size_t(stream) * geo.n_symbols;
        d.fft_tap = geo.fft_tap ? geo.fft_tap + size_t(stream) * geo.n_symbols * NFFT : nullptr;
        d.vec_tap = geo.vec_tap ? geo.vec_tap + size_t(stream) * (geo.n_symbols - 1) * geo.n_carriers : nullptr;
        FrameDesc* out = geo.descs + (size_t(pass) * geo.n_streams + stream) * geo.n_chunks;
        int b = 0;
        for (int c = 0; c < geo.n_chunks; c++) {
            d.s_begin = b;
            d.s_end = b + S / geo.n_chunks + (c < S % geo.n_chunks ? 1 : 0);
            b = d.s_end;
            out[c] = d;
        }
        st.pending_info.coarse_offset = st.freq_coarse;
        st.pending_info.fine_offset_used = st.freq_fine;
        st.pending_info.signal_average = st.l1_average;
        st.pending_info.total_desync = st.total_frames_desync;
        st.pending_slot = slot;
        st.pipeline_pending = 1;
        st.frames_in_call = slot + 1;
        st.state = DAB_OFDM_READING_NULL_AND_PRS;
        return true;
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
    __shared__ int stop_flag;
    const int stream = geo.stream0 + blockIdx.x, tid = threadIdx.x;
    // the work items this pass may fill start out invalid
    if (pass < geo.frame_passes && tid < geo.n_chunks) geo.descs[(size_t(pass) * geo.n_streams + stream) * geo.n_chunks + tid].valid = 0;

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
        if (pass == 0) {
            const uint64_t n = geo.n_per_stream ? geo.n_per_stream[stream] : geo.n_uniform;
            st.call_begin = st.call_end;
            st.call_end = st.call_begin + int64_t(n);
            st.avg_pending = (n > 0) ? 1 : 0;
            st.frames_in_call = 0;
        }
    }
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

    // UpdateSignalAverage belongs to the entry of Process() (ofdm_demodulator.cpp:241) and only reads the samples of this call, so
    // where in the call it runs is free.  It is DRAM-latency bound and the synchronisation transforms are issue bound: every other
    // CTA of an SM folds before its transforms instead of after them, so the two kinds of phase meet on the SM instead of all CTAs
    // waiting on memory together (pass 0 stays short: the frame kernel of this way waits for it).
    if (pass >= 1 && st.avg_pending && ((blockIdx.x / geo.n_sm) & 1)) ctl.fold_average();
    if (st.pipeline_pending) ctl.finish_pipeline();
    __syncthreads();

    // OFDM_Demod::Process main loop (ofdm_demodulator.cpp:245-274).  The stream state lives in shared memory and thread 0 mutates
    // it: every thread takes its copy of what steers the iteration, then a barrier, and only then may thread 0 move on -- without
    // it a warp that is late to the loop head can see the NEXT state and walk into a different case (divergent barriers; found by
    // compute-sanitizer synccheck on Mode II, where three of the four warps only wait at the barriers of the transforms).
    for (;;) {
        const bool go = st.consumed < st.call_end && !stop_flag;
        const int state = st.state;
        __syncthreads();
        if (!go) break;
        switch (state) {
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
            ctl.run_sync(state);
            break;
        case DAB_OFDM_READING_SYMBOLS:
            if (tid == 0) {
                if (ctl.read_symbols_thread0(pass)) stop_flag = 1;  // wait for the frame kernel before touching the next PRS
            }
            __syncthreads();
            break;
        }
    }
    __syncthreads();
    // the call is over for this stream: fold UpdateSignalAverage
    if (st.avg_pending && st.consumed >= st.call_end && !st.pipeline_pending) ctl.fold_average();
    if (tid == 0) {
        geo.states[stream] = st;
        geo.frames_in_call[stream] = st.frames_in_call;
    }
}

}  // namespace dabb200
