// Host side of the OFDM demodulator C ABI (include/dab_b200.h): owns the per-stream device state, feeds the stream rings,
// sequences  control -> frame kernel -> control  passes per Process() call and delivers the soft-bit callback.
// Replaces OFDM_Demod's constructor / Process / Reset / getters (reference src/ofdm/ofdm_demodulator.cpp:80-146, 235-289)
// and its coordinator / pipeline threads (ofdm_demodulator_threads.cpp), whose ordering becomes CUDA stream order.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <complex>
#include <map>
#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "ofdm_control.cuh"
#include "stage_layout.h"
#include "ofdm_frame.cuh"
#include "ofdm_frame_v3.cuh"

namespace dabb200 {

__global__ void ofdm_reset_stream_kernel(StreamState* states, int stream) {
    // OFDM_Demod::Reset (ofdm_demodulator.cpp:277-289)
    StreamState& st = states[stream];
    st.state = DAB_OFDM_FINDING_NULL_POWER_DIP;
    st.corr_length = 0;
    st.corr_explicit_len = 0;
    st.total_frames_desync++;
    st.is_found_coarse = 0;
    st.freq_coarse = 0.0f;
    st.freq_fine = 0.0f;
    st.fine_time_offset = 0;
}

// dab_ofdm_rebase_device_streams: every absolute sample index of a stream moves down by delta (the caller's resident buffer is a
// ring it refills behind the demodulator)
__global__ void ofdm_rebase_kernel(StreamState* states, int n_streams, int64_t delta) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    StreamState& st = states[s];
    st.corr_base -= delta;
    st.frame_start -= delta;
    st.consumed -= delta;
    st.call_begin -= delta;
    st.call_end -= delta;
    st.pending_info.frame_start -= delta;
}

struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Ofdm {
    dab_ofdm_params p{};
    int nfft = 0;
    int n_streams = 0;
    int device = 0;
    int format = DAB_IQ_F32;
    int sb = 8;                         // bytes per complex sample in the stream buffers
    SampleFmt fmt{};
    bool debug_taps = false;
    size_t max_block = 0;
    size_t ring_samples = 0;
    size_t max_pitch = 1u << 30;        // cudaDeviceProp::memPitch bound for pitched copies (set at create)
    int slots = 1;                      // frames a stream can complete in one call = soft-bit buffers per stream
    size_t frame_bits = 0;
    int syms_per_chunk = 0;             // DAB_B200_SYMS_PER_CHUNK: target symbols per frame-kernel work item, 0 = by geometry (create_impl)
    int n_chunks = 3;                   // work items per frame
    bool force_generic_kernel = false;  // DAB_B200_GENERIC_FRAME_KERNEL=1: run the generic-geometry frame kernel (tests)
    bool dab_geometry = false;          // the v3 kernel applies
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // pipeline ways: the streams of a handle are split into `ways` contiguous groups, each sequenced on its own CUDA stream, so
    // that the latency-bound control passes of one group overlap the frame kernel of another (DAB_B200_PIPELINE_WAYS, default 4)
    static constexpr int MAX_WAYS = 8;
    int ways = 4;
    int call_ways = 1;                  // ways of the most recent call (n_ways_for)
    uint64_t way_min_samples = uint64_t(1) << 24;   // DAB_B200_WAY_MIN_SAMPLES: samples per way and call below which ways are merged
    cudaStream_t way_stream[MAX_WAYS] = {};
    int n_sm = 148;                     // cudaDevAttrMultiProcessorCount (set at create)
    cudaEvent_t way_done[MAX_WAYS] = {};
    cudaEvent_t counts_ready[MAX_WAYS] = {};
    cudaEvent_t bits_ready[MAX_WAYS] = {};
    cudaEvent_t up_done[MAX_WAYS] = {};
    // host-buffer path: all uploads on one stream and all soft-bit downloads on another, each in way order, so that the PCIe
    // link serves the ways first-in first-out (copies queued on the way streams themselves share the link and all finish last)
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaEvent_t fork_event = nullptr;
    std::vector<dab_ofdm_config> cfgs;  // per-stream configuration as last set by the caller (survives init_states)
    bool ways_pending = false;          // way streams carry work the handle's stream has not been ordered after yet
    // device memory
    DeviceBuffer<unsigned char> ring_iq;
    DeviceBuffer<float2> null_ring, corr_explicit, prs_fft_ref_conj, prs_time_ref_conj, fft_tap, vec_tap, twiddles;
    DeviceBuffer<float> impulse, freq_resp, phase_err, stage_phase_err;
    DeviceBuffer<StreamState> states;
    DeviceBuffer<FrameDesc> descs, stage_descs;
    DeviceBuffer<dab_ofdm_frame_info> infos;
    DeviceBuffer<int32_t> frames_in_call;
    DeviceBuffer<int8_t> bits;
    DeviceBuffer<int16_t> bin_to_pos, bin_to_carrier, bin_to_slot;
    DeviceBuffer<uint16_t> chunk_src;
    int stage_wavefronts[3] = {0, 0, 0};   // store instructions per symbol, wavefronts in position order, wavefronts as laid out
    DeviceBuffer<uint64_t> d_n;
    // external (device resident) streams
    const void* ext_base = nullptr;
    size_t ext_stride = 0, ext_total = 0;
    // host bookkeeping
    std::vector<uint64_t> fed;       // samples handed to each stream so far
    std::vector<uint64_t> n_call;
    PinnedBuffer<uint64_t> h_n;
    PinnedBuffer<int32_t> h_frames;
    PinnedBuffer<dab_ofdm_frame_info> h_infos;
    PinnedBuffer<int8_t> h_bits;
    dab_ofdm_frame_cb cb = nullptr;
    void* cb_user = nullptr;
    uint64_t launches = 0;
    // host snapshot behind the scalar / small-array getters: refreshed at the end of every synchronous process call, read under
    // snap_mtx only -- a GUI thread polling GetState() never waits for a Process() in flight on the reader thread
    static constexpr int SNAP_ARRAYS_MAX_STREAMS = 4;   // response / frame-bit snapshots only for handles this small (the mirror class)
    PinnedBuffer<StreamState> h_states;
    PinnedBuffer<float> h_resp;         // [n_streams][2][nfft]: impulse response, coarse frequency response
    std::mutex snap_mtx;
    bool snap_valid = false;
    std::vector<dab_ofdm_state> snap;
    std::vector<int32_t> snap_pending_slot;
    std::vector<float> snap_resp;
    std::vector<int8_t> snap_bits;      // [n_streams][frame_bits]: last frame delivered through the callback
    std::vector<uint8_t> snap_bits_valid;
    // optional per-kernel event timing (roofline measurement)
    bool timing = false;
    struct TimedLaunch { cudaEvent_t start, stop; int pass; bool is_frame; };
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    dab_ofdm_kernel_times times{};
    std::recursive_mutex mtx;
};

static cudaEvent_t take_event(Ofdm* o) {
    if (!o->event_pool.empty()) {
        cudaEvent_t e = o->event_pool.back();
        o->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ScopedKernelTimer {
    Ofdm* o;
    Ofdm::TimedLaunch t{};
    bool on;
    cudaStream_t st;
    ScopedKernelTimer(Ofdm* o_, cudaStream_t st_, int pass, bool is_frame) : o(o_), on(o_->timing), st(st_) {
        if (!on) return;
        t.start = take_event(o);
        t.stop = take_event(o);
        t.pass = pass;
        t.is_frame = is_frame;
        cudaEventRecord(t.start, st);
    }
    ~ScopedKernelTimer() {
        if (!on) return;
        cudaEventRecord(t.stop, st);
        o->timed.push_back(t);
    }
};

// Orders the handle's stream after everything queued on the way streams.  run_call leaves the ways un-joined when it can, so
// that consecutive calls pipeline (no GPU-wide barrier per call); every entry point that exposes results joins first.
static int join_ways(Ofdm* o) {
    if (!o->ways_pending) return DAB_OK;
    for (int w = 0; w < o->ways; w++) {
        DAB_CUDA_CHECK(cudaEventRecord(o->way_done[w], o->way_stream[w]));
        DAB_CUDA_CHECK(cudaStreamWaitEvent(o->stream, o->way_done[w], 0));
    }
    o->ways_pending = false;
    return DAB_OK;
}

static int collect_times(Ofdm* o) {
    {
        int rc = join_ways(o);
        if (rc != DAB_OK) return rc;
    }
    DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));  // run_call joins the way streams into o->stream
    for (auto& t : o->timed) {
        float ms = 0.0f;
        DAB_CUDA_CHECK(cudaEventElapsedTime(&ms, t.start, t.stop));
        const int p = std::min(t.pass, DAB_OFDM_TIMING_PASSES - 1);
        if (t.is_frame) { o->times.frame_ms[p] += ms; o->times.frame_launches[p]++; }
        else { o->times.control_ms[p] += ms; o->times.control_launches[p]++; }
        o->event_pool.push_back(t.start);
        o->event_pool.push_back(t.stop);
    }
    o->timed.clear();
    return DAB_OK;
}

static void host_fft(std::vector<std::complex<double>>& x, int sign) {
    const size_t n = x.size();
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(x[i], x[j]);
    }
    const double two_pi = 6.283185307179586476925286766559;
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t k = 0; k < len / 2; k++) {
            const double a = double(sign) * two_pi * double(k) / double(len);
            const std::complex<double> w(std::cos(a), std::sin(a));
            for (size_t base = 0; base < n; base += len) {
                const auto u = x[base + k], t = x[base + k + len / 2] * w;
                x[base + k] = u + t;
                x[base + k + len / 2] = u - t;
            }
        }
    }
}

static FrameGeom frame_geom(const Ofdm* o) {
    FrameGeom g;
    g.n_symbols = int(o->p.nb_frame_symbols);
    g.symbol_period = int(o->p.nb_symbol_period);
    g.cyclic_prefix = int(o->p.nb_cyclic_prefix);
    g.n_carriers = int(o->p.nb_data_carriers);
    g.fmt = o->fmt;
    g.bin_to_pos = o->bin_to_pos.ptr;
    g.bin_to_carrier = o->bin_to_carrier.ptr;
    g.bin_to_slot = o->bin_to_slot.ptr;
    g.chunk_src = o->chunk_src.ptr;
    g.twiddles = o->twiddles.ptr;
    return g;
}

template <int NFFT, int SB>
static int launch_frame_t(Ofdm* o, cudaStream_t st, const FrameDesc* d_descs, int n_items) {
    const FrameGeom g = frame_geom(o);
    if (o->dab_geometry && !o->force_generic_kernel) {
        // the four DAB transmission modes, TMA-fed kernel with the separable PLL (ofdm_frame_v3.cuh)
        constexpr size_t smem = FrameV3Smem<NFFT, SB>::TOTAL_BYTES;
        constexpr int GROUPS = FrameV3Smem<NFFT, SB>::GROUPS;
        const int grid = (n_items + GROUPS - 1) / GROUPS;
        if (o->debug_taps) {
            DAB_CUDA_CHECK(cudaFuncSetAttribute(ofdm_frame_v3_kernel<NFFT, SB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            ofdm_frame_v3_kernel<NFFT, SB, true><<<grid, FRAME_CTA_THREADS, smem, st>>>(g, d_descs, n_items);
        } else {
            DAB_CUDA_CHECK(cudaFuncSetAttribute(ofdm_frame_v3_kernel<NFFT, SB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            ofdm_frame_v3_kernel<NFFT, SB, false><<<grid, FRAME_CTA_THREADS, smem, st>>>(g, d_descs, n_items);
        }
        o->launches++;
        DAB_CUDA_CHECK(cudaGetLastError());
        return DAB_OK;
    }
    const size_t smem = FrameSmem<NFFT>::total_bytes(g.n_carriers);
    DAB_CUDA_CHECK(cudaFuncSetAttribute(ofdm_frame_kernel<NFFT, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    constexpr int GROUPS = FrameSmem<NFFT>::GROUPS;
    const int grid = (n_items + GROUPS - 1) / GROUPS;
    ofdm_frame_kernel<NFFT, SB><<<grid, FRAME_CTA_THREADS, smem, st>>>(g, d_descs, n_items);
    o->launches++;
    DAB_CUDA_CHECK(cudaGetLastError());
    return DAB_OK;
}

template <int SB>
static int launch_frame_sb(Ofdm* o, cudaStream_t st, const FrameDesc* d_descs, int n_items) {
    switch (o->nfft) {
    case 2048: return launch_frame_t<2048, SB>(o, st, d_descs, n_items);
    case 1024: return launch_frame_t<1024, SB>(o, st, d_descs, n_items);
    case 512: return launch_frame_t<512, SB>(o, st, d_descs, n_items);
    case 256: return launch_frame_t<256, SB>(o, st, d_descs, n_items);
    }
    return set_error(DAB_ERR_INVALID, "unsupported FFT size %d", o->nfft);
}

static int launch_frame(Ofdm* o, cudaStream_t st, const FrameDesc* d_descs, int n_items) {
    if (n_items <= 0) return DAB_OK;
    switch (o->sb) {
    case 8: return launch_frame_sb<8>(o, st, d_descs, n_items);
    case 2: return launch_frame_sb<2>(o, st, d_descs, n_items);
    case 4: return launch_frame_sb<4>(o, st, d_descs, n_items);
    }
    return set_error(DAB_ERR_INVALID, "unsupported sample size %d", o->sb);
}

static ControlGeom control_geom(const Ofdm* o) {
    ControlGeom g;
    g.n_symbols = int(o->p.nb_frame_symbols);
    g.symbol_period = int(o->p.nb_symbol_period);
    g.null_period = int(o->p.nb_null_period);
    g.cyclic_prefix = int(o->p.nb_cyclic_prefix);
    g.n_carriers = int(o->p.nb_data_carriers);
    g.slots = o->slots;
    g.n_streams = o->n_streams;
    g.stream0 = 0;
    g.n_chunks = o->n_chunks;
    g.frame_passes = 0;
    g.frame_bits = o->frame_bits;
    if (o->ext_base) {
        g.mask = ~uint64_t(0);
        g.limit = o->ext_total;
        g.stream_stride = o->ext_stride;
        g.samples = o->ext_base;
    } else {
        g.mask = uint64_t(o->ring_samples) - 1;
        g.limit = o->ring_samples;
        g.stream_stride = o->ring_samples;
        g.samples = o->ring_iq.ptr;
    }
    g.fmt = o->fmt;
    g.n_sm = o->n_sm;
    g.ring = o->null_ring.ptr;
    g.corr_explicit = o->corr_explicit.ptr;
    g.prs_fft_ref_conj = o->prs_fft_ref_conj.ptr;
    g.prs_time_ref_conj = o->prs_time_ref_conj.ptr;
    g.impulse_response = o->impulse.ptr;
    g.freq_response = o->freq_resp.ptr;
    g.states = o->states.ptr;
    g.descs = o->descs.ptr;
    g.infos = o->infos.ptr;
    g.frames_in_call = o->frames_in_call.ptr;
    g.bits = o->bits.ptr;
    g.phase_err = o->phase_err.ptr;
    g.twiddles = o->twiddles.ptr;
    g.fft_tap = o->debug_taps ? o->fft_tap.ptr : nullptr;
    g.vec_tap = o->debug_taps ? o->vec_tap.ptr : nullptr;
    g.n_per_stream = nullptr;
    g.n_uniform = 0;
    return g;
}

template <int NFFT, int SB>
static int launch_control_t(Ofdm* o, cudaStream_t st, const ControlGeom& g, int count, int pass) {
    const size_t smem = ControlSmem<NFFT>::bytes();
    DAB_CUDA_CHECK(cudaFuncSetAttribute(ofdm_control_kernel<NFFT, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    ofdm_control_kernel<NFFT, SB><<<count, ControlSmem<NFFT>::THREADS, smem, st>>>(g, pass);
    o->launches++;
    DAB_CUDA_CHECK(cudaGetLastError());
    return DAB_OK;
}

template <int SB>
static int launch_control_sb(Ofdm* o, cudaStream_t st, const ControlGeom& g, int count, int pass) {
    switch (o->nfft) {
    case 2048: return launch_control_t<2048, SB>(o, st, g, count, pass);
    case 1024: return launch_control_t<1024, SB>(o, st, g, count, pass);
    case 512: return launch_control_t<512, SB>(o, st, g, count, pass);
    case 256: return launch_control_t<256, SB>(o, st, g, count, pass);
    }
    return set_error(DAB_ERR_INVALID, "unsupported FFT size %d", o->nfft);
}

static int launch_control(Ofdm* o, cudaStream_t st, const ControlGeom& g, int count, int pass) {
    switch (o->sb) {
    case 8: return launch_control_sb<8>(o, st, g, count, pass);
    case 2: return launch_control_sb<2>(o, st, g, count, pass);
    case 4: return launch_control_sb<4>(o, st, g, count, pass);
    }
    return set_error(DAB_ERR_INVALID, "unsupported sample size %d", o->sb);
}

static size_t sample_bytes(const Ofdm* o) { return size_t(o->sb); }

// frame completions possible inside one call of n samples: consecutive frame ends are at least frame_cap - cp apart
static int passes_for(const Ofdm* o, uint64_t n_max) {
    if (n_max == 0) return 0;
    const uint64_t frame_cap = o->p.nb_frame_symbols * o->p.nb_symbol_period + o->p.nb_null_period;
    return 1 + int((n_max - 1) / (frame_cap - o->p.nb_cyclic_prefix));
}

struct WayRange {
    int lo, hi;
    cudaStream_t st;
};
// Pipeline ways of a call: splitting pays when every way still fills the GPU for a while; a call with little work per stream
// (4096-sample blocks, the short frames of Modes II / III) is launch bound, and every extra way multiplies its launches.
static int n_ways_for(const Ofdm* o, uint64_t n_max) {
    if (o->n_streams < 64 * o->ways) return 1;
    const uint64_t work = uint64_t(o->n_streams) * n_max;   // samples in this call
    if (work >= o->way_min_samples * uint64_t(o->ways)) return o->ways;
    if (work * 4 >= o->way_min_samples * 5 && o->ways >= 2) return 2;   // 1024 Mode III frames (25 M samples): two ways measured 4 - 16 % faster than one
    return 1;
}
static int n_ways(const Ofdm* o) { return o->call_ways; }   // of the most recent call
static WayRange way_range(const Ofdm* o, int w, int ways) {
    WayRange r;
    r.lo = int(int64_t(o->n_streams) * w / ways);
    r.hi = int(int64_t(o->n_streams) * (w + 1) / ways);
    r.st = (ways > 1) ? o->way_stream[w] : o->stream;
    return r;
}

// (control -> frame)* -> control  for the streams of one way, in order on the way's CUDA stream.  Control pass 0 opens the call
// (OFDM_Demod::Process's entry); the frame kernel after pass p runs the work items pass p wrote; the pass in which a stream
// finishes its block folds the block's UpdateSignalAverage windows into the running average.
static int issue_way_kernels(Ofdm* o, const WayRange& r, bool uniform, uint64_t n_uniform, int passes) {
    const int count = r.hi - r.lo;
    cudaStream_t st = r.st;
    ControlGeom g = control_geom(o);
    g.stream0 = r.lo;
    g.frame_passes = passes;
    g.n_per_stream = uniform ? nullptr : o->d_n.ptr;
    g.n_uniform = n_uniform;
    for (int p = 0; p <= passes; p++) {
        int rc;
        {
            ScopedKernelTimer timer(o, st, p, false);
            rc = launch_control(o, st, g, count, p);
        }
        if (rc != DAB_OK) return rc;
        if (p < passes) {
            ScopedKernelTimer timer(o, st, p, true);
            rc = launch_frame(o, st, g.descs + (size_t(p) * size_t(o->n_streams) + size_t(r.lo)) * size_t(o->n_chunks), count * o->n_chunks);
            if (rc != DAB_OK) return rc;
        }
    }
    return DAB_OK;
}

// Host blocks of the streams [r.lo, r.hi) -> their rings.  When the caller's blocks are rows of one array (equal length, equal
// ring position, constant pointer stride: what a batched front end hands over) the whole way goes up as one pitched copy (two
// when the ring wraps) instead of one copy per stream: per-copy driver and DMA set-up cost is what keeps 1024 x 1.5 MB copies
// at 48 GB/s and 1024 x 0.4 MB copies far lower, against 55 GB/s for one large copy (profiles/r01e_pcie.md).
static int upload_blocks(Ofdm* o, const WayRange& r, const void* const* iq, size_t sb, cudaStream_t up) {
    const int count = r.hi - r.lo;
    bool strided = count >= 2 && iq[r.lo] != nullptr && o->n_call[size_t(r.lo)] > 0;
    ptrdiff_t pitch = 0;
    if (strided) {
        const uint64_t n0 = o->n_call[size_t(r.lo)], f0 = o->fed[size_t(r.lo)];
        pitch = static_cast<const unsigned char*>(iq[r.lo + 1]) - static_cast<const unsigned char*>(iq[r.lo]);
        if (pitch < ptrdiff_t(n0 * sb) || size_t(pitch) > o->max_pitch) strided = false;
        for (int s = r.lo; strided && s < r.hi; s++) {
            if (o->n_call[size_t(s)] != n0 || o->fed[size_t(s)] != f0) strided = false;
            else if (static_cast<const unsigned char*>(iq[s]) != static_cast<const unsigned char*>(iq[r.lo]) + ptrdiff_t(s - r.lo) * pitch) strided = false;
        }
    }
    if (strided) {
        const size_t n = size_t(o->n_call[size_t(r.lo)]);
        const size_t pos = size_t(o->fed[size_t(r.lo)] & (o->ring_samples - 1));
        const size_t first = std::min(n, o->ring_samples - pos);
        unsigned char* base = o->ring_iq.ptr + size_t(r.lo) * o->ring_samples * sb;
        const size_t dpitch = o->ring_samples * sb;
        DAB_CUDA_CHECK(cudaMemcpy2DAsync(base + pos * sb, dpitch, iq[r.lo], size_t(pitch), first * sb, size_t(count), cudaMemcpyHostToDevice, up));
        if (first < n)
            DAB_CUDA_CHECK(cudaMemcpy2DAsync(base, dpitch, static_cast<const unsigned char*>(iq[r.lo]) + first * sb, size_t(pitch), (n - first) * sb, size_t(count),
                                             cudaMemcpyHostToDevice, up));
        return DAB_OK;
    }
    for (int s = r.lo; s < r.hi; s++) {
        const size_t n = size_t(o->n_call[size_t(s)]);
        if (n == 0) continue;
        const size_t pos = size_t(o->fed[size_t(s)] & (o->ring_samples - 1));
        const size_t first = std::min(n, o->ring_samples - pos);
        unsigned char* base = o->ring_iq.ptr + size_t(s) * o->ring_samples * sb;
        DAB_CUDA_CHECK(cudaMemcpyAsync(base + pos * sb, iq[s], first * sb, cudaMemcpyHostToDevice, up));
        if (first < n)
            DAB_CUDA_CHECK(cudaMemcpyAsync(base, static_cast<const unsigned char*>(iq[s]) + first * sb, (n - first) * sb, cudaMemcpyHostToDevice, up));
    }
    return DAB_OK;
}

// One Process() call for every stream.  iq == nullptr: the n_call[s] new samples are already visible at [fed[s], fed[s] +
// n_call[s]) (attached device streams).  Otherwise iq[s] is the caller's host block, copied into the stream ring first.
//
// The streams are split into pipeline ways.  Every way runs  upload -> (control -> frame)* -> control -> download of the frame
// counts  in order on its own CUDA stream (per-stream ordering is the reference's real-time order, ofdm_control.cuh); different
// ways only share the GPU and the PCIe link, so one way's control passes (latency bound) hide behind another way's frame kernel,
// and uploads, kernels and downloads of different ways overlap.
static int run_call(Ofdm* o, bool uniform, uint64_t n_uniform, const void* const* iq, bool snapshot) {
    NvtxRange nvtx_call("dab_ofdm call");
    uint64_t n_max = 0;
    if (uniform) {
        n_max = n_uniform;
    } else {
        for (int s = 0; s < o->n_streams; s++) n_max = std::max(n_max, o->n_call[size_t(s)]);
        DAB_CUDA_CHECK(o->h_n.reserve(size_t(o->n_streams)));
        memcpy(o->h_n.ptr, o->n_call.data(), sizeof(uint64_t) * size_t(o->n_streams));
        DAB_CUDA_CHECK(cudaMemcpyAsync(o->d_n.ptr, o->h_n.ptr, sizeof(uint64_t) * size_t(o->n_streams), cudaMemcpyHostToDevice, o->stream));
    }
    const int passes = passes_for(o, n_max);
    if (passes > o->slots) return set_error(DAB_ERR_CAPACITY, "call of %llu samples exceeds max_block_samples", (unsigned long long)n_max);
    const int ways = n_ways_for(o, n_max);
    if (ways != o->call_ways) {   // the stream -> way mapping changes: order everything queued under the old mapping first
        int rc = join_ways(o);
        if (rc != DAB_OK) return rc;
        o->call_ways = ways;
    }
    const size_t sb = sample_bytes(o), ns = size_t(o->n_streams), slots = size_t(o->slots);
    if (o->cb || snapshot) {
        DAB_CUDA_CHECK(o->h_frames.reserve(ns));
        DAB_CUDA_CHECK(o->h_infos.reserve(ns * slots));
    }
    if (snapshot) DAB_CUDA_CHECK(o->h_states.reserve(ns));
    if (ways > 1) DAB_CUDA_CHECK(cudaEventRecord(o->fork_event, o->stream));
    for (int w = 0; w < ways; w++) {
        NvtxRange nvtx_way("dab_ofdm way");
        const WayRange r = way_range(o, w, ways);
        if (ways > 1) DAB_CUDA_CHECK(cudaStreamWaitEvent(r.st, o->fork_event, 0));
        if (iq) {
            // the caller's span is only valid during the call (ofdm_demodulator.cpp:235): copy into the stream ring now
            cudaStream_t up = (ways > 1) ? o->up_stream : r.st;
            if (ways > 1 && w == 0) DAB_CUDA_CHECK(cudaStreamWaitEvent(up, o->fork_event, 0));
            int rc_up = upload_blocks(o, r, iq, sb, up);
            if (rc_up != DAB_OK) return rc_up;
            if (ways > 1) {
                DAB_CUDA_CHECK(cudaEventRecord(o->up_done[w], up));
                DAB_CUDA_CHECK(cudaStreamWaitEvent(r.st, o->up_done[w], 0));
            }
        }
        int rc = issue_way_kernels(o, r, uniform, n_uniform, passes);
        if (rc != DAB_OK) return rc;
        const size_t cnt = size_t(r.hi - r.lo);
        if (o->cb) {
            DAB_CUDA_CHECK(cudaMemcpyAsync(o->h_frames.ptr + r.lo, o->frames_in_call.ptr + r.lo, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, r.st));
            DAB_CUDA_CHECK(cudaMemcpyAsync(o->h_infos.ptr + size_t(r.lo) * slots, o->infos.ptr + size_t(r.lo) * slots, cnt * slots * sizeof(dab_ofdm_frame_info),
                                           cudaMemcpyDeviceToHost, r.st));
        }
        if (snapshot) {
            DAB_CUDA_CHECK(cudaMemcpyAsync(o->h_states.ptr + r.lo, o->states.ptr + r.lo, cnt * sizeof(StreamState), cudaMemcpyDeviceToHost, r.st));
            if (o->n_streams <= Ofdm::SNAP_ARRAYS_MAX_STREAMS) {
                const size_t nfft = size_t(o->nfft);
                DAB_CUDA_CHECK(o->h_resp.reserve(ns * 2 * nfft));
                for (int s = r.lo; s < r.hi; s++) {
                    DAB_CUDA_CHECK(cudaMemcpyAsync(o->h_resp.ptr + size_t(s) * 2 * nfft, o->impulse.ptr + size_t(s) * nfft, nfft * sizeof(float), cudaMemcpyDeviceToHost, r.st));
                    DAB_CUDA_CHECK(cudaMemcpyAsync(o->h_resp.ptr + size_t(s) * 2 * nfft + nfft, o->freq_resp.ptr + size_t(s) * nfft, nfft * sizeof(float), cudaMemcpyDeviceToHost, r.st));
                }
            }
        }
        DAB_CUDA_CHECK(cudaEventRecord(o->counts_ready[w], r.st));
    }
    if (ways > 1) {
        o->ways_pending = true;
        // per-stream sample counts live in one device buffer that the next call overwrites from the handle's stream
        if (!uniform) {
            int rc = join_ways(o);
            if (rc != DAB_OK) return rc;
        }
    }
    if (o->timed.size() > 4096) {
        int rc = collect_times(o);
        if (rc != DAB_OK) return rc;
    }
    for (int s = 0; s < o->n_streams; s++) o->fed[size_t(s)] += uniform ? n_uniform : o->n_call[size_t(s)];
    return DAB_OK;
}

static int run_callbacks(Ofdm* o, int w, int ways, size_t k) {
    const WayRange r = way_range(o, w, ways);
    const size_t slots = size_t(o->slots), fb = o->frame_bits;
    DAB_CUDA_CHECK(cudaEventSynchronize(o->bits_ready[w]));
    const bool keep_bits = o->n_streams <= Ofdm::SNAP_ARRAYS_MAX_STREAMS;
    for (int s = r.lo; s < r.hi; s++)
        for (int f = 0; f < o->h_frames.ptr[s]; f++, k++) {
            if (keep_bits) {  // GetFrameDataBits() between calls is served from this copy
                std::lock_guard<std::mutex> snap_lock(o->snap_mtx);
                o->snap_bits.resize(size_t(o->n_streams) * fb);
                o->snap_bits_valid.resize(size_t(o->n_streams), 0);
                memcpy(o->snap_bits.data() + size_t(s) * fb, o->h_bits.ptr + k * fb, fb);
                o->snap_bits_valid[size_t(s)] = 1;
            }
            o->cb(o->cb_user, s, o->h_bits.ptr + k * fb, fb, &o->h_infos.ptr[size_t(s) * slots + size_t(f)]);
        }
    return DAB_OK;
}

// soft-bit callback delivery (CoordinatorThread's Notify, ofdm_demodulator.cpp:635), in stream then frame order.  Way by way:
// as soon as a way's frame counts are on the host its soft bits are fetched (consecutive streams as one strided copy), and its
// callbacks run while the later ways are still uploading / computing.
static int deliver(Ofdm* o) {
    if (!o->cb) return DAB_OK;
    const int ways = n_ways(o);
    const size_t slots = size_t(o->slots), fb = o->frame_bits;
    DAB_CUDA_CHECK(o->h_bits.reserve(size_t(o->n_streams) * slots * fb));
    std::vector<size_t> first_k(size_t(ways) + 1, 0);
    for (int w = 0; w < ways; w++) {
        const WayRange r = way_range(o, w, ways);
        DAB_CUDA_CHECK(cudaEventSynchronize(o->counts_ready[w]));
        size_t k = first_k[size_t(w)];
        // device source of frame (s, f) is (s * slots + f) * fb, host destination (callback order: stream, then frame) is k * fb.
        // Consecutive streams that completed the same number of frames c form a run whose frame f goes down as ONE pitched copy
        // (rows of fb bytes, source pitch slots * fb, destination pitch c * fb): one copy per way in the steady state instead of
        // one per stream -- 1024 separate 230 KB copies cost 6.7 ms of set-up per step, more than the transfer itself.
        cudaStream_t down = (ways > 1) ? o->down_stream : r.st;  // counts_ready[w] (synchronised above) follows the way's kernels
        for (int s0 = r.lo; s0 < r.hi;) {
            const int c = o->h_frames.ptr[s0];
            int s1 = s0 + 1;
            while (s1 < r.hi && o->h_frames.ptr[s1] == c) s1++;
            if (c > 0) {
                const size_t rows = size_t(s1 - s0);
                for (int f = 0; f < c; f++) {
                    const int8_t* src = o->bits.ptr + (size_t(s0) * slots + size_t(f)) * fb;
                    int8_t* dst = o->h_bits.ptr + (k + size_t(f)) * fb;
                    if (size_t(c) == slots && c == 1) {
                        DAB_CUDA_CHECK(cudaMemcpyAsync(dst, src, rows * fb, cudaMemcpyDeviceToHost, down));
                    } else {
                        DAB_CUDA_CHECK(cudaMemcpy2DAsync(dst, size_t(c) * fb, src, slots * fb, fb, rows, cudaMemcpyDeviceToHost, down));
                    }
                }
                k += rows * size_t(c);
            }
            s0 = s1;
        }
        first_k[size_t(w) + 1] = k;
        DAB_CUDA_CHECK(cudaEventRecord(o->bits_ready[w], down));
        // the previous way's soft bits arrived while this way was uploading / computing: hand them over now
        if (w > 0) {
            int rc2 = run_callbacks(o, w - 1, ways, first_k[size_t(w) - 1]);
            if (rc2 != DAB_OK) return rc2;
        }
    }
    {
        int rc2 = run_callbacks(o, ways - 1, ways, first_k[size_t(ways) - 1]);
        if (rc2 != DAB_OK) return rc2;
    }
    return join_ways(o);
}

static void fill_state(dab_ofdm_state* out, const StreamState& st) {
    out->state = st.state;
    out->fine_time_offset = st.fine_time_offset;
    out->total_frames_read = st.total_frames_read;
    out->total_frames_desync = st.total_frames_desync;
    out->signal_average = st.l1_average;
    out->fine_frequency_offset = st.freq_fine;
    out->coarse_frequency_offset = st.freq_coarse;
    out->reserved = 0;
}

// after a synchronous call: the states (and, for small handles, the sync responses) copied by run_call become the getter snapshot
static void publish_snapshot(Ofdm* o) {
    std::lock_guard<std::mutex> snap_lock(o->snap_mtx);
    o->snap.resize(size_t(o->n_streams));
    o->snap_pending_slot.resize(size_t(o->n_streams));
    for (int s = 0; s < o->n_streams; s++) {
        fill_state(&o->snap[size_t(s)], o->h_states.ptr[s]);
        o->snap_pending_slot[size_t(s)] = o->h_states.ptr[s].pending_slot;
    }
    if (o->n_streams <= Ofdm::SNAP_ARRAYS_MAX_STREAMS && o->h_resp.ptr)
        o->snap_resp.assign(o->h_resp.ptr, o->h_resp.ptr + size_t(o->n_streams) * 2 * size_t(o->nfft));
    o->snap_valid = true;
}

static void invalidate_snapshot(Ofdm* o) {
    std::lock_guard<std::mutex> snap_lock(o->snap_mtx);
    o->snap_valid = false;
    std::fill(o->snap_bits_valid.begin(), o->snap_bits_valid.end(), uint8_t(0));
}

static int init_states(Ofdm* o) {
    // the configuration is the caller's (OFDM_Demod::GetConfig() is mutable state of the object, not of a run): it survives a
    // re-initialisation of the stream states, e.g. attaching device-resident streams after dab_ofdm_set_config
    if (o->cfgs.size() != size_t(o->n_streams)) {
        o->cfgs.resize(size_t(o->n_streams));
        for (auto& c : o->cfgs) dab_ofdm_default_config(&c);
    }
    std::vector<StreamState> init(size_t(o->n_streams));
    for (size_t s = 0; s < init.size(); s++) {
        StreamState& st = init[s];
        memset(&st, 0, sizeof(st));
        st.state = DAB_OFDM_FINDING_NULL_POWER_DIP;
        st.cfg = o->cfgs[s];
    }
    DAB_CUDA_CHECK(cudaMemcpy(o->states.ptr, init.data(), init.size() * sizeof(StreamState), cudaMemcpyHostToDevice));
    DAB_CUDA_CHECK(cudaMemset(o->null_ring.ptr, 0, o->null_ring.count * sizeof(float2)));
    DAB_CUDA_CHECK(cudaMemset(o->frames_in_call.ptr, 0, o->frames_in_call.count * sizeof(int32_t)));
    DAB_CUDA_CHECK(cudaMemset(o->descs.ptr, 0, o->descs.count * sizeof(FrameDesc)));
    std::fill(o->fed.begin(), o->fed.end(), 0);
    invalidate_snapshot(o);
    return DAB_OK;
}

// Staging layout of the v3 frame kernel for one carrier map (stage_layout.h): the search runs once per process and map.
static const StageLayout& stage_layout_for(int nfft, int ncarr, const std::vector<int16_t>& b2p, bool optimise) {
    static std::mutex mtx;
    static std::map<std::vector<int16_t>, StageLayout> cache;
    std::vector<int16_t> key(b2p);
    key.push_back(int16_t(optimise ? 1 : 0));
    std::lock_guard<std::mutex> lock(mtx);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    // the kernel's store instructions: register slot r of the lanes of one warp (or of one transform when it has fewer threads)
    const int T = nfft / 16, R3 = nfft / 256, lanes = std::min(T, 32);
    std::vector<std::vector<int>> groups;
    for (int t0 = 0; t0 < T; t0 += lanes)
        for (int r = 0; r < 16; r++) {
            std::vector<int> g;
            bool any = false;
            for (int l = 0; l < lanes; l++) {
                const int t = t0 + l;
                const int bin = (R3 == 1) ? t + 16 * r : (t + T * (r / R3)) + 256 * (r % R3);   // fft_out_bin (ofdm_device.cuh)
                g.push_back(bin);
                any = any || b2p[size_t(bin)] >= 0;
            }
            if (any) groups.push_back(std::move(g));
        }
    return cache.emplace(std::move(key), stage_layout_optimise(groups, b2p, ncarr, optimise, 20000)).first->second;
}

static int create_impl(Ofdm* o, const dab_c32* prs, const int* mapper) {
    const size_t nfft = o->p.nb_fft, ncarr = o->p.nb_data_carriers, ns = size_t(o->n_streams);
    const size_t frame_cap = o->p.nb_frame_symbols * o->p.nb_symbol_period + o->p.nb_null_period;
    o->frame_bits = (o->p.nb_frame_symbols - 1) * ncarr * 2;
    o->slots = std::max(1, passes_for(o, o->max_block));
    o->dab_geometry = (o->nfft == 2048 && DabGeom<2048>::matches(int(o->p.nb_symbol_period), int(o->p.nb_cyclic_prefix), int(ncarr))) ||
                      (o->nfft == 1024 && DabGeom<1024>::matches(int(o->p.nb_symbol_period), int(o->p.nb_cyclic_prefix), int(ncarr))) ||
                      (o->nfft == 512 && DabGeom<512>::matches(int(o->p.nb_symbol_period), int(o->p.nb_cyclic_prefix), int(ncarr))) ||
                      (o->nfft == 256 && DabGeom<256>::matches(int(o->p.nb_symbol_period), int(o->p.nb_cyclic_prefix), int(ncarr)));
    // Work items per frame: pieces of about syms_per_chunk symbols.  An item costs one more transform (its differential reference)
    // and its symbols run one after the other, so long items are cheap in work and long in latency.  Where the frame kernel's grid
    // is several waves of long transforms (Modes I / IV, hundreds of streams) 26-symbol items are the measured optimum; with the
    // short transforms of Modes II / III, or few streams, the launch is one or two waves and its duration is the length of an item:
    // 13 symbols (Mode III, 1024 streams: 0.48 -> 0.40 ms per frame period; 256 streams: -10 %), 8 for a handful of streams.
    if (o->syms_per_chunk <= 0) o->syms_per_chunk = (o->n_streams < 64) ? 8 : (o->nfft >= 1024 && o->n_streams >= 512) ? 26 : 13;
    o->n_chunks = std::max(1, std::min(FRAME_MAX_CHUNKS, int((o->p.nb_frame_symbols + size_t(o->syms_per_chunk) - 1) / size_t(o->syms_per_chunk))));
    size_t need = frame_cap + o->p.nb_null_period + o->p.nb_symbol_period + o->max_block + 1024;
    o->ring_samples = 1;
    while (o->ring_samples < need) o->ring_samples <<= 1;

    // constant tables: conj(PRS) and conj(IFFT(conj(PRS[i]) PRS[i+1])) (ofdm_demodulator.cpp:128-140), bin -> soft-bit position
    std::vector<float2> ref_conj(nfft), time_conj(nfft);
    std::vector<std::complex<double>> rel(nfft);
    for (size_t i = 0; i < nfft; i++) ref_conj[i] = make_float2(prs[i].re, -prs[i].im);
    for (size_t i = 0; i + 1 < nfft; i++) {
        const float ar = prs[i].re, ai = -prs[i].im, br = prs[i + 1].re, bi = prs[i + 1].im;
        rel[i] = std::complex<double>(double(ar * br - ai * bi), double(ar * bi + ai * br));
    }
    rel[nfft - 1] = 0.0;
    host_fft(rel, +1);
    for (size_t i = 0; i < nfft; i++) time_conj[i] = make_float2(float(rel[i].real()), -float(rel[i].imag()));
    std::vector<int16_t> b2p(nfft, int16_t(-1)), b2c(nfft, int16_t(-1));
    std::vector<int> inverse(ncarr, -1);
    for (size_t i = 0; i < ncarr; i++) {
        if (mapper[i] < 0 || size_t(mapper[i]) >= ncarr) return set_error(DAB_ERR_INVALID, "carrier_mapper[%zu] = %d out of range", i, mapper[i]);
        inverse[size_t(mapper[i])] = int(i);
    }
    const int half = int(ncarr / 2);
    for (int c = 0; c < 2 * half; c++) {  // CalculateDQPSK carrier order: -M..-1, +1..+M (ofdm_demodulator.cpp:853-864)
        const int k = (c < half) ? (c - half) : (c - half + 1);
        const size_t bin = size_t((int(nfft) + k) % int(nfft));
        b2c[bin] = int16_t(c);
        b2p[bin] = int16_t(inverse[size_t(c)]);
    }

    DAB_CUDA_CHECK(cudaStreamCreateWithFlags(&o->own_stream, cudaStreamNonBlocking));
    o->stream = o->own_stream;
    {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, o->device) == cudaSuccess && v > 0) o->n_sm = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxPitch, o->device) == cudaSuccess && v > 0) o->max_pitch = size_t(v);
    }
    for (int w = 0; w < Ofdm::MAX_WAYS; w++) {
        DAB_CUDA_CHECK(cudaStreamCreateWithFlags(&o->way_stream[w], cudaStreamNonBlocking));
        DAB_CUDA_CHECK(cudaEventCreateWithFlags(&o->way_done[w], cudaEventDisableTiming));
        DAB_CUDA_CHECK(cudaEventCreateWithFlags(&o->counts_ready[w], cudaEventDisableTiming));
        DAB_CUDA_CHECK(cudaEventCreateWithFlags(&o->bits_ready[w], cudaEventDisableTiming));
        DAB_CUDA_CHECK(cudaEventCreateWithFlags(&o->up_done[w], cudaEventDisableTiming));
    }
    DAB_CUDA_CHECK(cudaStreamCreateWithFlags(&o->up_stream, cudaStreamNonBlocking));
    DAB_CUDA_CHECK(cudaStreamCreateWithFlags(&o->down_stream, cudaStreamNonBlocking));
    DAB_CUDA_CHECK(cudaEventCreateWithFlags(&o->fork_event, cudaEventDisableTiming));
    DAB_CUDA_CHECK(o->ring_iq.reserve(ns * o->ring_samples * sample_bytes(o)));
    DAB_CUDA_CHECK(cudaMemset(o->ring_iq.ptr, 0, ns * o->ring_samples * sample_bytes(o)));
    DAB_CUDA_CHECK(o->null_ring.reserve(ns * o->p.nb_null_period));
    DAB_CUDA_CHECK(o->corr_explicit.reserve(ns * o->p.nb_null_period));
    DAB_CUDA_CHECK(o->prs_fft_ref_conj.reserve(nfft));
    DAB_CUDA_CHECK(o->prs_time_ref_conj.reserve(nfft));
    DAB_CUDA_CHECK(o->impulse.reserve(ns * nfft));
    DAB_CUDA_CHECK(o->freq_resp.reserve(ns * nfft));
    DAB_CUDA_CHECK(cudaMemset(o->impulse.ptr, 0, ns * nfft * sizeof(float)));
    DAB_CUDA_CHECK(cudaMemset(o->freq_resp.ptr, 0, ns * nfft * sizeof(float)));
    DAB_CUDA_CHECK(o->states.reserve(ns));
    DAB_CUDA_CHECK(o->descs.reserve(ns * size_t(o->slots) * size_t(o->n_chunks)));
    DAB_CUDA_CHECK(o->infos.reserve(ns * size_t(o->slots)));
    DAB_CUDA_CHECK(cudaMemset(o->infos.ptr, 0, ns * size_t(o->slots) * sizeof(dab_ofdm_frame_info)));
    DAB_CUDA_CHECK(o->frames_in_call.reserve(ns));
    DAB_CUDA_CHECK(o->bits.reserve(ns * size_t(o->slots) * o->frame_bits));
    DAB_CUDA_CHECK(cudaMemset(o->bits.ptr, 0, ns * size_t(o->slots) * o->frame_bits));
    DAB_CUDA_CHECK(o->phase_err.reserve(ns * o->p.nb_frame_symbols));
    DAB_CUDA_CHECK(cudaMemset(o->phase_err.ptr, 0, ns * o->p.nb_frame_symbols * sizeof(float)));
    DAB_CUDA_CHECK(o->bin_to_pos.reserve(nfft));
    DAB_CUDA_CHECK(o->bin_to_carrier.reserve(nfft));
    DAB_CUDA_CHECK(o->d_n.reserve(ns));
    {
        const int n_tw = int(nfft) + 16 * int(nfft / 256);  // TW1_SIZE + TW2_SIZE
        DAB_CUDA_CHECK(o->twiddles.reserve(size_t(n_tw)));
        const int blocks = (n_tw + 127) / 128;
        switch (o->nfft) {
        case 2048: fft_twiddle_init_kernel<2048><<<blocks, 128>>>(o->twiddles.ptr); break;
        case 1024: fft_twiddle_init_kernel<1024><<<blocks, 128>>>(o->twiddles.ptr); break;
        case 512: fft_twiddle_init_kernel<512><<<blocks, 128>>>(o->twiddles.ptr); break;
        case 256: fft_twiddle_init_kernel<256><<<blocks, 128>>>(o->twiddles.ptr); break;
        }
        DAB_CUDA_CHECK(cudaGetLastError());
        DAB_CUDA_CHECK(cudaDeviceSynchronize());
    }
    if (o->debug_taps) {
        DAB_CUDA_CHECK(o->fft_tap.reserve(ns * o->p.nb_frame_symbols * nfft));
        DAB_CUDA_CHECK(o->vec_tap.reserve(ns * (o->p.nb_frame_symbols - 1) * ncarr));
        DAB_CUDA_CHECK(cudaMemset(o->fft_tap.ptr, 0, o->fft_tap.count * sizeof(float2)));
        DAB_CUDA_CHECK(cudaMemset(o->vec_tap.ptr, 0, o->vec_tap.count * sizeof(float2)));
    }
    DAB_CUDA_CHECK(cudaMemcpy(o->prs_fft_ref_conj.ptr, ref_conj.data(), nfft * sizeof(float2), cudaMemcpyHostToDevice));
    DAB_CUDA_CHECK(cudaMemcpy(o->prs_time_ref_conj.ptr, time_conj.data(), nfft * sizeof(float2), cudaMemcpyHostToDevice));
    DAB_CUDA_CHECK(cudaMemcpy(o->bin_to_pos.ptr, b2p.data(), nfft * sizeof(int16_t), cudaMemcpyHostToDevice));
    DAB_CUDA_CHECK(cudaMemcpy(o->bin_to_carrier.ptr, b2c.data(), nfft * sizeof(int16_t), cudaMemcpyHostToDevice));
    {
        const StageLayout& lay = stage_layout_for(int(nfft), int(ncarr), b2p, o->dab_geometry && ncarr % 64 == 0);
        DAB_CUDA_CHECK(o->bin_to_slot.reserve(nfft));
        DAB_CUDA_CHECK(o->chunk_src.reserve(lay.chunk_src.size()));
        DAB_CUDA_CHECK(cudaMemcpy(o->bin_to_slot.ptr, lay.bin_to_slot.data(), nfft * sizeof(int16_t), cudaMemcpyHostToDevice));
        DAB_CUDA_CHECK(cudaMemcpy(o->chunk_src.ptr, lay.chunk_src.data(), lay.chunk_src.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        o->stage_wavefronts[0] = lay.store_instructions;
        o->stage_wavefronts[1] = lay.wavefronts_before;
        o->stage_wavefronts[2] = lay.wavefronts_after;
    }
    o->fed.assign(ns, 0);
    o->n_call.assign(ns, 0);
    return init_states(o);
}

}  // namespace dabb200

using namespace dabb200;

extern "C" {

void dab_ofdm_default_config(dab_ofdm_config* c) {
    if (!c) return;
    c->signal_l1_update_beta = 0.95f;
    c->signal_l1_nb_samples = 100;
    c->signal_l1_nb_decimate = 5;
    c->null_l1_thresh_null_start = 0.35f;
    c->null_l1_thresh_null_end = 0.75f;
    c->sync_fine_freq_update_beta = 0.9f;
    c->sync_is_coarse_freq_correction = 1;
    c->sync_max_coarse_freq_correction_norm = 0.5f;
    c->sync_coarse_freq_slow_beta = 0.1f;
    c->sync_impulse_peak_threshold_db = 20.0f;
    c->sync_impulse_peak_distance_probability = 0.15f;
}

static bool sample_format(int format, int* sb, SampleFmt* f) {
    // examples/app_helpers/app_iq_readers.h:20-35: BIAS / MAX_AMPLITUDE per integer type (signed: 0 / max, unsigned: max/2 + 0.5 twice);
    // a signed component is read as unsigned ^ sign bit, which raises the bias by 2^(bits-1)
    f->flip = 0;
    f->prmt = 0x3210;
    f->bias = 0.0f;
    f->scale = 1.0f;
    switch (format) {
    case DAB_IQ_F32: *sb = 8; return true;
    case DAB_IQ_U8: *sb = 2; f->bias = 127.5f; f->scale = 1.0f / 127.5f; return true;
    case DAB_IQ_S8: *sb = 2; f->flip = 0x80u; f->bias = 128.0f; f->scale = 1.0f / 127.0f; return true;
    case DAB_IQ_S16LE: case DAB_IQ_S16BE: *sb = 4; f->flip = 0x8000u; f->bias = 32768.0f; f->scale = 1.0f / 32767.0f; break;
    case DAB_IQ_U16LE: case DAB_IQ_U16BE: *sb = 4; f->bias = 32767.5f; f->scale = 1.0f / 32767.5f; break;
    default: return false;
    }
    if (format == DAB_IQ_S16BE || format == DAB_IQ_U16BE) f->prmt = 0x2301;
    return true;
}

size_t dab_iq_format_bytes(int format) {
    int sb = 0;
    SampleFmt f;
    return sample_format(format, &sb, &f) ? size_t(sb) : 0;
}

dab_ofdm* dab_ofdm_create(const dab_ofdm_params* params, const dab_c32* prs_fft_ref, const int* carrier_mapper, const dab_ofdm_options* options,
                          int* status) {
    auto fail = [&](int rc) -> dab_ofdm* { if (status) *status = rc; return nullptr; };
    if (!params || !prs_fft_ref || !carrier_mapper || !options) return fail(set_error(DAB_ERR_INVALID, "null argument"));
    if (options->n_streams < 1) return fail(set_error(DAB_ERR_INVALID, "n_streams must be >= 1"));
    const size_t nfft = params->nb_fft;
    if (nfft != 256 && nfft != 512 && nfft != 1024 && nfft != 2048) return fail(set_error(DAB_ERR_INVALID, "nb_fft %zu not in {256, 512, 1024, 2048}", nfft));
    if (params->nb_symbol_period != nfft + params->nb_cyclic_prefix || params->nb_cyclic_prefix > nfft / 4 || params->nb_cyclic_prefix == 0)
        return fail(set_error(DAB_ERR_INVALID, "symbol period / cyclic prefix not supported (need 0 < cp <= nfft/4)"));
    if (params->nb_data_carriers >= nfft || params->nb_data_carriers % 8 != 0 || params->nb_frame_symbols < 2 ||
        params->nb_null_period < params->nb_cyclic_prefix)
        return fail(set_error(DAB_ERR_INVALID, "unsupported OFDM geometry"));
    int sb = 0;
    SampleFmt fmt;
    if (!sample_format(options->sample_format, &sb, &fmt)) return fail(set_error(DAB_ERR_INVALID, "unknown sample_format %d", options->sample_format));
    DeviceGuard guard;
    int rc = select_device(options->device);
    if (rc != DAB_OK) return fail(rc);
    auto* o = new Ofdm();
    o->p = *params;
    o->nfft = int(nfft);
    o->n_streams = options->n_streams;
    o->device = options->device;
    o->format = options->sample_format;
    o->sb = sb;
    o->fmt = fmt;
    o->debug_taps = options->keep_debug_taps != 0;
    o->max_block = options->max_block_samples ? options->max_block_samples : 262144;
    if (const char* e = getenv("DAB_B200_GENERIC_FRAME_KERNEL")) o->force_generic_kernel = (e[0] == '1');
    if (const char* e = getenv("DAB_B200_PIPELINE_WAYS")) { const int w = atoi(e); if (w >= 1 && w <= Ofdm::MAX_WAYS) o->ways = w; }
    if (const char* e = getenv("DAB_B200_SYMS_PER_CHUNK")) { const int c = atoi(e); if (c >= 1 && c <= 1024) o->syms_per_chunk = c; }
    if (const char* e = getenv("DAB_B200_WAY_MIN_SAMPLES")) { const long long v = atoll(e); if (v >= 0) o->way_min_samples = uint64_t(v); }
    rc = create_impl(o, prs_fft_ref, carrier_mapper);
    if (rc != DAB_OK) { delete o; return fail(rc); }
    if (status) *status = DAB_OK;
    return reinterpret_cast<dab_ofdm*>(o);
}

void dab_ofdm_destroy(dab_ofdm* h) {
    auto* o = reinterpret_cast<Ofdm*>(h);
    if (!o) return;
    DeviceGuard guard;
    cudaSetDevice(o->device);
    cudaStreamSynchronize(o->stream);
    for (auto& t : o->timed) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
    for (auto e : o->event_pool) cudaEventDestroy(e);
    for (int w = 0; w < Ofdm::MAX_WAYS; w++) {
        if (o->way_stream[w]) { cudaStreamSynchronize(o->way_stream[w]); cudaStreamDestroy(o->way_stream[w]); }
        if (o->way_done[w]) cudaEventDestroy(o->way_done[w]);
        if (o->counts_ready[w]) cudaEventDestroy(o->counts_ready[w]);
        if (o->bits_ready[w]) cudaEventDestroy(o->bits_ready[w]);
        if (o->up_done[w]) cudaEventDestroy(o->up_done[w]);
    }
    if (o->up_stream) { cudaStreamSynchronize(o->up_stream); cudaStreamDestroy(o->up_stream); }
    if (o->down_stream) { cudaStreamSynchronize(o->down_stream); cudaStreamDestroy(o->down_stream); }
    if (o->fork_event) cudaEventDestroy(o->fork_event);
    if (o->own_stream) cudaStreamDestroy(o->own_stream);
    delete o;
}

// The handle's lock is recursive: the frame callback runs inside a process call and may call back into the same handle from the
// same thread (see include/dab_b200.h).  The caller's current CUDA device is restored when the call returns.
#define OFDM_HANDLE(h)                                              \
    auto* o = reinterpret_cast<Ofdm*>(h);                           \
    if (!o) return set_error(DAB_ERR_INVALID, "null handle");       \
    std::lock_guard<std::recursive_mutex> lock(o->mtx);             \
    DeviceGuard device_guard;                                       \
    DAB_CUDA_CHECK(cudaSetDevice(o->device))

int dab_ofdm_set_cuda_stream(dab_ofdm* h, void* cuda_stream) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));
    o->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : o->own_stream;
    return DAB_OK;
}

void dab_ofdm_count_frames_cb(void* user, int stream, const int8_t* bits, size_t n_bits, const dab_ofdm_frame_info* info) {
    auto* c = static_cast<dab_ofdm_frame_counter*>(user);
    if (!c) return;
    uint64_t sum = uint64_t(stream) + uint64_t(info ? info->frame_start : 0);
    const size_t edge = n_bits < 64 ? n_bits : 64;
    for (size_t i = 0; i < edge; i++) sum = sum * 31u + uint8_t(bits[i]) + uint8_t(bits[n_bits - 1 - i]);
    c->frames++;
    c->bits += n_bits;
    c->checksum += sum;
}

int dab_ofdm_set_frame_callback(dab_ofdm* h, dab_ofdm_frame_cb cb, void* user) {
    OFDM_HANDLE(h);
    o->cb = cb;
    o->cb_user = user;
    return DAB_OK;
}

int dab_ofdm_set_config(dab_ofdm* h, int stream, const dab_ofdm_config* cfg) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    if (!cfg || stream < -1 || stream >= o->n_streams) return set_error(DAB_ERR_INVALID, "bad stream / config");
    const int lo = (stream < 0) ? 0 : stream, hi = (stream < 0) ? o->n_streams : stream + 1;
    for (int s = lo; s < hi; s++) o->cfgs[size_t(s)] = *cfg;
    // one strided copy into the cfg member of every selected stream state
    DAB_CUDA_CHECK(cudaMemcpy2DAsync(reinterpret_cast<char*>(o->states.ptr + lo) + offsetof(StreamState, cfg), sizeof(StreamState), &o->cfgs[size_t(lo)],
                                     sizeof(dab_ofdm_config), sizeof(dab_ofdm_config), size_t(hi - lo), cudaMemcpyHostToDevice, o->stream));
    DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));
    return DAB_OK;
}

int dab_ofdm_get_config(dab_ofdm* h, int stream, dab_ofdm_config* cfg) {
    OFDM_HANDLE(h);
    if (!cfg || stream < 0 || stream >= o->n_streams) return set_error(DAB_ERR_INVALID, "bad stream / config");
    *cfg = o->cfgs[size_t(stream)];   // the device copies are only ever written from this host copy
    return DAB_OK;
}

static int ingest_and_run(Ofdm* o, const void* const* iq, const size_t* n, int format) {
    if (o->ext_base) return set_error(DAB_ERR_INVALID, "handle is attached to device-resident streams; use dab_ofdm_advance");
    if (format >= 0 && format != o->format) return set_error(DAB_ERR_INVALID, "handle was created with sample_format = %d", o->format);
    for (int s = 0; s < o->n_streams; s++) {
        const size_t ns = n[s];
        if (ns > o->max_block) return set_error(DAB_ERR_CAPACITY, "stream %d: block of %zu samples exceeds max_block_samples %zu", s, ns, o->max_block);
        if (ns > 0 && !iq[s]) return set_error(DAB_ERR_INVALID, "stream %d: null sample pointer", s);
        o->n_call[size_t(s)] = ns;
    }
    int rc = run_call(o, false, 0, iq, true);
    if (rc != DAB_OK) return rc;
    // the caller's span (possibly pageable memory, staged by the driver) must have been consumed before the call returns:
    // deliver() waits for every way's frame counts, which follow the way's upload; without a callback wait here
    if (o->cb) {
        rc = deliver(o);
        if (rc != DAB_OK) return rc;
    } else {
        const int ways = n_ways(o);
        for (int w = 0; w < ways; w++) DAB_CUDA_CHECK(cudaEventSynchronize(o->counts_ready[w]));
    }
    publish_snapshot(o);
    return DAB_OK;
}

int dab_ofdm_process_batch(dab_ofdm* h, const dab_c32* const* iq, const size_t* n) {
    OFDM_HANDLE(h);
    if (!iq || !n) return set_error(DAB_ERR_INVALID, "null argument");
    return ingest_and_run(o, reinterpret_cast<const void* const*>(iq), n, DAB_IQ_F32);
}

int dab_ofdm_process_batch_u8(dab_ofdm* h, const uint8_t* const* iq_u8, const size_t* n) {
    OFDM_HANDLE(h);
    if (!iq_u8 || !n) return set_error(DAB_ERR_INVALID, "null argument");
    return ingest_and_run(o, reinterpret_cast<const void* const*>(iq_u8), n, DAB_IQ_U8);
}

int dab_ofdm_process_batch_raw(dab_ofdm* h, const void* const* iq, const size_t* n) {
    OFDM_HANDLE(h);
    if (!iq || !n) return set_error(DAB_ERR_INVALID, "null argument");
    return ingest_and_run(o, iq, n, -1);
}

int dab_ofdm_process(dab_ofdm* h, int stream, const dab_c32* iq, size_t n) {
    OFDM_HANDLE(h);
    if (stream < 0 || stream >= o->n_streams) return set_error(DAB_ERR_INVALID, "stream %d out of range", stream);
    std::vector<const void*> ptrs(size_t(o->n_streams), nullptr);
    std::vector<size_t> ns(size_t(o->n_streams), 0);
    ptrs[size_t(stream)] = iq;
    ns[size_t(stream)] = n;
    return ingest_and_run(o, ptrs.data(), ns.data(), DAB_IQ_F32);
}

static int attach_impl(dab_ofdm* h, const void* d_iq, size_t stride_samples, size_t total_samples, bool raw) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    if (!d_iq || stride_samples < total_samples) return set_error(DAB_ERR_INVALID, "bad device stream geometry");
    if (!raw && o->format != DAB_IQ_F32) return set_error(DAB_ERR_INVALID, "handle was created with sample_format = %d: use dab_ofdm_attach_device_streams_raw", o->format);
    DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));
    o->ext_base = d_iq;
    o->ext_stride = stride_samples;
    o->ext_total = total_samples;
    return init_states(o);
}

int dab_ofdm_attach_device_streams(dab_ofdm* h, const dab_c32* d_iq, size_t stride_samples, size_t total_samples) {
    return attach_impl(h, d_iq, stride_samples, total_samples, false);
}

int dab_ofdm_attach_device_streams_raw(dab_ofdm* h, const void* d_iq, size_t stride_samples, size_t total_samples) {
    return attach_impl(h, d_iq, stride_samples, total_samples, true);
}

int dab_ofdm_rebase_device_streams(dab_ofdm* h, size_t delta_samples) {
    OFDM_HANDLE(h);
    if (!o->ext_base) return set_error(DAB_ERR_INVALID, "no device-resident streams attached");
    for (int s = 0; s < o->n_streams; s++)
        if (o->fed[size_t(s)] < delta_samples) return set_error(DAB_ERR_INVALID, "stream %d has not advanced %zu samples yet", s, delta_samples);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    // the frame being received and the NULL + PRS window must lie at or after the new origin: the demodulator never looks further
    // back than one frame + one NULL symbol + one symbol behind its cursor
    const size_t keep = o->p.nb_frame_symbols * o->p.nb_symbol_period + 2 * o->p.nb_null_period + o->p.nb_symbol_period;
    for (int s = 0; s < o->n_streams; s++)
        if (o->fed[size_t(s)] - delta_samples < keep)
            return set_error(DAB_ERR_INVALID, "stream %d: fewer than %zu samples would remain behind the cursor after rebasing", s, keep);
    ofdm_rebase_kernel<<<(o->n_streams + 127) / 128, 128, 0, o->stream>>>(o->states.ptr, o->n_streams, int64_t(delta_samples));
    o->launches++;
    DAB_CUDA_CHECK(cudaGetLastError());
    for (int s = 0; s < o->n_streams; s++) o->fed[size_t(s)] -= delta_samples;
    invalidate_snapshot(o);
    return DAB_OK;
}

static int advance_impl(Ofdm* o, const size_t* n, size_t n_uniform) {
    if (!o->ext_base) return set_error(DAB_ERR_INVALID, "no device-resident streams attached");
    for (int s = 0; s < o->n_streams; s++) {
        const size_t ns = n ? n[s] : n_uniform;
        if (ns > o->max_block) return set_error(DAB_ERR_CAPACITY, "block of %zu samples exceeds max_block_samples %zu", ns, o->max_block);
        if (o->fed[size_t(s)] + ns > o->ext_total) return set_error(DAB_ERR_CAPACITY, "stream %d: advance past the end of the attached buffer", s);
        o->n_call[size_t(s)] = ns;
    }
    invalidate_snapshot(o);   // asynchronous: the getters synchronise on demand
    int rc = run_call(o, n == nullptr, n_uniform, nullptr, false);
    if (rc != DAB_OK) return rc;
    if (n != nullptr) DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));  // the pinned staging copy of n[] is reused by the next call
    return deliver(o);
}

int dab_ofdm_advance(dab_ofdm* h, const size_t* n) {
    OFDM_HANDLE(h);
    if (!n) return set_error(DAB_ERR_INVALID, "null argument");
    return advance_impl(o, n, 0);
}

int dab_ofdm_advance_uniform(dab_ofdm* h, size_t n) {
    OFDM_HANDLE(h);
    return advance_impl(o, nullptr, n);
}

int dab_ofdm_device_bits(dab_ofdm* h, const int8_t** d_bits, size_t* n_bits, int* slots_per_stream, const int32_t** d_frames_in_call) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    if (d_bits) *d_bits = o->bits.ptr;
    if (n_bits) *n_bits = o->frame_bits;
    if (slots_per_stream) *slots_per_stream = o->slots;
    if (d_frames_in_call) *d_frames_in_call = o->frames_in_call.ptr;
    return DAB_OK;
}

int dab_ofdm_reset(dab_ofdm* h, int stream) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    if (stream < 0 || stream >= o->n_streams) return set_error(DAB_ERR_INVALID, "stream %d out of range", stream);
    invalidate_snapshot(o);
    ofdm_reset_stream_kernel<<<1, 1, 0, o->stream>>>(o->states.ptr, stream);
    o->launches++;
    DAB_CUDA_CHECK(cudaGetLastError());
    return DAB_OK;
}

int dab_ofdm_get_state(dab_ofdm* h, int stream, dab_ofdm_state* out) {
    auto* o0 = reinterpret_cast<Ofdm*>(h);
    if (!o0) return set_error(DAB_ERR_INVALID, "null handle");
    if (!out || stream < 0 || stream >= o0->n_streams) return set_error(DAB_ERR_INVALID, "bad stream / output");
    {   // fast path: the snapshot of the last synchronous call -- no GPU work, and no wait for a call in flight on another thread
        std::lock_guard<std::mutex> snap_lock(o0->snap_mtx);
        if (o0->snap_valid) {
            *out = o0->snap[size_t(stream)];
            return DAB_OK;
        }
    }
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    StreamState st;
    DAB_CUDA_CHECK(cudaMemcpyAsync(&st, o->states.ptr + stream, sizeof(st), cudaMemcpyDeviceToHost, o->stream));
    DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));
    fill_state(out, st);
    return DAB_OK;
}

int dab_ofdm_join(dab_ofdm* h) {
    OFDM_HANDLE(h);
    return join_ways(o);
}

int dab_ofdm_sync(dab_ofdm* h) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));
    return DAB_OK;
}

size_t dab_ofdm_frame_bits(const dab_ofdm* h) {
    auto* o = reinterpret_cast<const Ofdm*>(h);
    return o ? o->frame_bits : 0;
}

int dab_ofdm_get_params(const dab_ofdm* h, dab_ofdm_params* out) {
    auto* o = reinterpret_cast<const Ofdm*>(h);
    if (!o || !out) return set_error(DAB_ERR_INVALID, "null argument");
    *out = o->p;
    return DAB_OK;
}

static int copy_out(Ofdm* o, void* dst, const void* d_src, size_t bytes) {
    DAB_CUDA_CHECK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, o->stream));
    DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));
    return DAB_OK;
}

// which = 0: impulse response, 1: coarse frequency response
static int get_response(dab_ofdm* h, int stream, float* out, size_t nb_fft, int which) {
    auto* o0 = reinterpret_cast<Ofdm*>(h);
    if (!o0) return set_error(DAB_ERR_INVALID, "null handle");
    if (!out || stream < 0 || stream >= o0->n_streams || nb_fft != size_t(o0->nfft)) return set_error(DAB_ERR_INVALID, "bad argument");
    {
        std::lock_guard<std::mutex> snap_lock(o0->snap_mtx);
        if (o0->snap_valid && o0->snap_resp.size() == size_t(o0->n_streams) * 2 * nb_fft) {
            memcpy(out, o0->snap_resp.data() + (size_t(stream) * 2 + size_t(which)) * nb_fft, nb_fft * sizeof(float));
            return DAB_OK;
        }
    }
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    return copy_out(o, out, (which ? o->freq_resp.ptr : o->impulse.ptr) + size_t(stream) * nb_fft, nb_fft * sizeof(float));
}

int dab_ofdm_get_impulse_response(dab_ofdm* h, int stream, float* out, size_t nb_fft) { return get_response(h, stream, out, nb_fft, 0); }

int dab_ofdm_get_coarse_frequency_response(dab_ofdm* h, int stream, float* out, size_t nb_fft) { return get_response(h, stream, out, nb_fft, 1); }

int dab_ofdm_get_frame_data_bits(dab_ofdm* h, int stream, int8_t* out, size_t n_bits) {
    auto* o0 = reinterpret_cast<Ofdm*>(h);
    if (!o0) return set_error(DAB_ERR_INVALID, "null handle");
    if (!out || stream < 0 || stream >= o0->n_streams || n_bits != o0->frame_bits) return set_error(DAB_ERR_INVALID, "bad argument");
    {
        std::lock_guard<std::mutex> snap_lock(o0->snap_mtx);
        if (size_t(stream) < o0->snap_bits_valid.size() && o0->snap_bits_valid[size_t(stream)]) {
            memcpy(out, o0->snap_bits.data() + size_t(stream) * n_bits, n_bits);
            return DAB_OK;
        }
    }
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    StreamState st;
    int rc = copy_out(o, &st, o->states.ptr + stream, sizeof(st));
    if (rc != DAB_OK) return rc;
    return copy_out(o, out, o->bits.ptr + (size_t(stream) * size_t(o->slots) + size_t(st.pending_slot)) * o->frame_bits, n_bits);
}

int dab_ofdm_get_correlation_time_buffer(dab_ofdm* h, int stream, dab_c32* out, size_t n) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    const size_t cap = o->p.nb_null_period + o->p.nb_symbol_period;
    if (!out || stream < 0 || stream >= o->n_streams || n != cap) return set_error(DAB_ERR_INVALID, "bad argument (n must be nb_null_period + nb_symbol_period)");
    if (o->format != DAB_IQ_F32) return set_error(DAB_ERR_INVALID, "not available with a raw sample_format");
    StreamState st;
    int rc = copy_out(o, &st, o->states.ptr + stream, sizeof(st));
    if (rc != DAB_OK) return rc;
    memset(out, 0, cap * sizeof(dab_c32));
    const size_t filled = std::min<size_t>(cap, st.corr_length);
    const size_t expl = std::min<size_t>(filled, st.corr_explicit_len);
    if (expl) {
        rc = copy_out(o, out, o->corr_explicit.ptr + size_t(stream) * o->p.nb_null_period, expl * sizeof(float2));
        if (rc != DAB_OK) return rc;
    }
    for (size_t i = expl; i < filled;) {  // stream-backed part, split where the ring wraps
        const uint64_t abs_index = uint64_t(st.corr_base + int64_t(i));
        size_t run = filled - i;
        const float2* src;
        if (o->ext_base) {
            src = reinterpret_cast<const float2*>(o->ext_base) + size_t(stream) * o->ext_stride + abs_index;
        } else {
            const uint64_t pos = abs_index & (o->ring_samples - 1);
            run = std::min<size_t>(run, o->ring_samples - pos);
            src = reinterpret_cast<const float2*>(o->ring_iq.ptr) + size_t(stream) * o->ring_samples + pos;
        }
        rc = copy_out(o, out + i, src, run * sizeof(float2));
        if (rc != DAB_OK) return rc;
        i += run;
    }
    return DAB_OK;
}

int dab_ofdm_get_frame_fft(dab_ofdm* h, int stream, dab_c32* out, size_t n) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    const size_t want = o->p.nb_frame_symbols * size_t(o->nfft);
    if (!o->debug_taps) return set_error(DAB_ERR_INVALID, "handle was created without keep_debug_taps");
    if (!out || stream < 0 || stream >= o->n_streams || n != want) return set_error(DAB_ERR_INVALID, "bad argument (n must be nb_frame_symbols * nb_fft)");
    return copy_out(o, out, o->fft_tap.ptr + size_t(stream) * want, want * sizeof(float2));
}

int dab_ofdm_get_frame_data_vec(dab_ofdm* h, int stream, dab_c32* out, size_t n) {
    OFDM_HANDLE(h);
    { int jrc = join_ways(o); if (jrc != DAB_OK) return jrc; }
    const size_t want = (o->p.nb_frame_symbols - 1) * o->p.nb_data_carriers;
    if (!o->debug_taps) return set_error(DAB_ERR_INVALID, "handle was created without keep_debug_taps");
    if (!out || stream < 0 || stream >= o->n_streams || n != want) return set_error(DAB_ERR_INVALID, "bad argument (n must be (nb_frame_symbols-1) * nb_data_carriers)");
    return copy_out(o, out, o->vec_tap.ptr + size_t(stream) * want, want * sizeof(float2));
}

int dab_ofdm_set_kernel_timing(dab_ofdm* h, int enable) {
    OFDM_HANDLE(h);
    int rc = collect_times(o);
    if (rc != DAB_OK) return rc;
    o->timing = enable != 0;
    memset(&o->times, 0, sizeof(o->times));
    return DAB_OK;
}

int dab_ofdm_get_kernel_times(dab_ofdm* h, dab_ofdm_kernel_times* out) {
    OFDM_HANDLE(h);
    if (!out) return set_error(DAB_ERR_INVALID, "null argument");
    int rc = collect_times(o);
    if (rc != DAB_OK) return rc;
    *out = o->times;
    return DAB_OK;
}

uint64_t dab_ofdm_kernel_launches(const dab_ofdm* h) {
    auto* o = reinterpret_cast<const Ofdm*>(h);
    return o ? o->launches : 0;
}

int dab_ofdm_demod_frames_device(dab_ofdm* h, const dab_c32* d_frames, size_t frame_stride, int n_frames, const float* freq_offset, int8_t* d_bits,
                                 float* d_phase_error) {
    OFDM_HANDLE(h);
    if (!d_frames || !freq_offset || !d_bits || n_frames < 0) return set_error(DAB_ERR_INVALID, "null argument");
    if (o->format != DAB_IQ_F32) return set_error(DAB_ERR_INVALID, "stage-level entry takes complex float frames");
    if (n_frames == 0) return DAB_OK;
    const int S = int(o->p.nb_frame_symbols), parts = o->n_chunks;
    std::vector<FrameDesc> descs(static_cast<size_t>(n_frames) * size_t(parts));
    for (int f = 0; f < n_frames; f++) {
        FrameDesc d;
        memset(&d, 0, sizeof(d));
        d.src = reinterpret_cast<const float2*>(d_frames) + size_t(f) * frame_stride;
        d.mask = ~uint64_t(0);
        d.limit = o->p.nb_frame_symbols * o->p.nb_symbol_period;
        d.start = 0;
        d.freq = freq_offset[f];
        d.valid = 1;
        d.bits = d_bits + size_t(f) * o->frame_bits;
        d.phase_err = d_phase_error ? d_phase_error + size_t(f) * o->p.nb_frame_symbols : nullptr;
        int b = 0;
        for (int c = 0; c < parts; c++) {
            d.s_begin = b;
            d.s_end = b + S / parts + (c < S % parts ? 1 : 0);
            b = d.s_end;
            descs[size_t(f) * size_t(parts) + size_t(c)] = d;
        }
    }
    DAB_CUDA_CHECK(o->stage_descs.reserve(descs.size()));
    DAB_CUDA_CHECK(cudaMemcpyAsync(o->stage_descs.ptr, descs.data(), descs.size() * sizeof(FrameDesc), cudaMemcpyHostToDevice, o->stream));
    int rc = launch_frame(o, o->stream, o->stage_descs.ptr, int(descs.size()));
    // `descs` is a local: the upload must have left it before we return
    DAB_CUDA_CHECK(cudaStreamSynchronize(o->stream));
    return rc;
}

}  // extern "C"
