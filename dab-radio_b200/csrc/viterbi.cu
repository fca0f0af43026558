// DAB Viterbi decoder for sm_100a: K = 7, rate 1/4 mother code with ETSI puncturing.
//
// Replaces DAB_Viterbi_Decoder (reference src/dab/algorithms/dab_viterbi_decoder.cpp:103-181) on top of
// ViterbiDecoder_AVX_u16<7,4> (vendor/viterbi_decoder/include/viterbi/x86/viterbi_decoder_avx_u16.h:47-170) and
// ViterbiDecoder_Core::chainback (viterbi_decoder_core.h:214-236).  Output bytes and the u64 path error are bit-exact
// with that decoder: saturating u16 path metrics, tie -> predecessor 1, renormalise when metric[0] >= 60455.
//
// Mapping: one warp runs one 64-state trellis (one FIC group or one sub-channel CIF).
//   * lane s owns butterfly s: it reads old[s] and old[s+32] and produces new[2s] (low half) and new[2s+1] (high half)
//     of one packed u16x2 register.  The two predecessor metrics arrive with two SHFLs of the packed register.
//   * branch metric sum_r |t_r - sym_r| is one VABSDIFF4.ACC on the packed int8 symbols; add-compare-select is
//     2x VIADD.16x2 + one VIMNMX.U16x2 whose predicate outputs are the two decision bits.
//   * the packed adds wrap; the reference saturates.  A warp vote after every step tells whether any metric is within one
//     step of 65535, and only then the next step runs the explicit saturating 32-bit path, so results are identical.
//   * depuncturing is fused into the symbol load: every 32 steps lane i fetches the <= 4 kept symbols of step t0+i by
//     walking the cyclic count table arithmetically; the step loop broadcasts them with one SHFL.
//   * survivor decisions (two ballots per step) stay in shared memory when the trellis fits the per-warp window,
//     otherwise they stream to a global scratch in 256-byte rows and are paged back per window for the traceback.
//   * traceback walks the 8-bit register of ViterbiTracebackBuffer<7> (state = reg >> 2) from shared memory.
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace dabb200 {

constexpr int VIT_WARPS_PER_CTA = 4;
constexpr uint32_t VIT_MAX_ERROR = 1016;        // (127 - -127) * 4, dab_viterbi_decoder.cpp:31
constexpr uint32_t VIT_NON_START = 5080;        // 5 * max_error, :32-36
constexpr uint32_t VIT_RENORM = 60455;          // 65535 - 5080, :37
constexpr uint32_t VIT_NEAR_SAT = 65535 - 1020; // a metric below this cannot saturate in the next step (e, 1016-e <= 1020)

// device-side form of one update() call, with the depuncture walk pre-digested on the host
struct DevSegment {
    uint32_t first_step;   // trellis step at which the segment starts
    uint32_t n_steps;      // n_out / 4
    uint32_t soft_start;   // punctured symbols consumed by the previous segments
    uint32_t period_syms;  // kept symbols per full cycle of the code
    uint32_t code_len;
    uint32_t counts;       // 8 x 4-bit kept count (1..4) per group
    uint32_t prefix_lo;    // 4 x 8-bit: kept symbols before group r within a cycle, r = 0..3
    uint32_t prefix_hi;    // r = 4..7
};

struct DevSchedule {
    DevSegment seg[DAB_VIT_MAX_SEGMENTS];
    uint32_t n_seg;
    uint32_t total_steps;
    uint32_t n_out_bits;
    uint32_t soft_symbols;
    uint32_t start_state;
    uint32_t end_state;
    uint32_t pad[2];
};

__device__ __forceinline__ uint32_t vabsdiff4_sum(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("vabsdiff4.u32.s32.s32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u));
    return d;
}

__device__ __forceinline__ int parity7(uint32_t v) { return __popc(v & 0x7Fu) & 1; }

struct StepOut {
    uint32_t pack;
    bool d0, d1;
};

// One add-compare-select for butterfly `lane`.  m0m0 / m1m1: predecessor metrics duplicated into both halves.
template <bool SATURATING>
__device__ __forceinline__ StepOut acs(uint32_t m0m0, uint32_t m1m1, uint32_t e, uint32_t inv) {
    StepOut o;
    if (!SATURATING) {
        const uint32_t a = __vadd2(m0m0, e | (inv << 16));   // lo: m0 + e      hi: m0 + (1016 - e)
        const uint32_t b = __vadd2(m1m1, inv | (e << 16));   // lo: m1 + (1016-e) hi: m1 + e
        bool p_hi, p_lo;
        o.pack = __vibmin_u16x2(b, a, &p_hi, &p_lo);         // pred = (b <= a): tie keeps the m1 branch, decision bit 1
        o.d0 = p_lo;
        o.d1 = p_hi;
    } else {
        const uint32_t m0 = m0m0 & 0xFFFFu, m1 = m1m1 & 0xFFFFu;
        const uint32_t a_lo = min(m0 + e, 65535u), b_lo = min(m1 + inv, 65535u);
        const uint32_t a_hi = min(m0 + inv, 65535u), b_hi = min(m1 + e, 65535u);
        o.d0 = b_lo <= a_lo;
        o.d1 = b_hi <= a_hi;
        o.pack = min(a_lo, b_lo) | (min(a_hi, b_hi) << 16);
    }
    return o;
}

// decisions of one step: bit s of w0 = decision of state 2s, bit s of w1 = decision of state 2s+1
__device__ __forceinline__ uint32_t decision_bit(uint2 w, uint32_t state) {
    const uint32_t word = (state & 1u) ? w.y : w.x;
    return (word >> (state >> 1)) & 1u;
}

__global__ void __launch_bounds__(VIT_WARPS_PER_CTA * 32)
viterbi_kernel(const int8_t* __restrict__ soft, size_t soft_bytes, const dab_vit_job* __restrict__ jobs, int n_jobs,
               const DevSchedule* __restrict__ schedules, int n_schedules, uint8_t* __restrict__ out, size_t out_bytes,
               uint64_t* __restrict__ path_error, int32_t* __restrict__ job_status, uint2* __restrict__ scratch,
               uint32_t scratch_steps_per_job, uint32_t window_steps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job_index = blockIdx.x * VIT_WARPS_PER_CTA + warp;
    if (job_index >= n_jobs) return;

    // per-warp carve-up: [decision window][schedule copy]
    const size_t per_warp = size_t(window_steps) * sizeof(uint2) + sizeof(DevSchedule);
    uint2* win = reinterpret_cast<uint2*>(smem_raw + size_t(warp) * per_warp);
    DevSchedule* sch = reinterpret_cast<DevSchedule*>(smem_raw + size_t(warp) * per_warp + size_t(window_steps) * sizeof(uint2));

    const dab_vit_job job = jobs[job_index];
    int status = DAB_OK;
    if (job.schedule >= uint32_t(n_schedules)) status = DAB_ERR_INVALID;
    if (status == DAB_OK) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&schedules[job.schedule]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sch);
        for (int i = lane; i < int(sizeof(DevSchedule) / 4); i += 32) dst[i] = src[i];
        __syncwarp();
        if (sch->soft_symbols > job.n_soft || job.soft_offset + sch->soft_symbols > soft_bytes) status = DAB_ERR_UNDERRUN;
        else if (sch->n_out_bits + 6u > sch->total_steps) status = DAB_ERR_TRACEBACK;
        else if (job.out_offset + sch->n_out_bits / 8u > out_bytes) status = DAB_ERR_CAPACITY;
    }
    if (status != DAB_OK) {
        if (lane == 0) {
            if (job_status) job_status[job_index] = status;
            if (path_error) path_error[job_index] = 0;
        }
        return;
    }

    const uint32_t total_steps = sch->total_steps;
    const bool long_mode = total_steps > window_steps;
    uint2* spill = long_mode ? (scratch + size_t(job_index) * scratch_steps_per_job) : nullptr;
    const int8_t* my_soft = soft + job.soft_offset;

    // branch table of ViterbiBranchTable<7,4> (viterbi_branch_table.h:44-52) for butterfly `lane`, one byte per polynomial
    uint32_t table4 = 0;
    {
        const uint32_t G[4] = {109, 79, 83, 109};
#pragma unroll
        for (int r = 0; r < 4; r++) table4 |= (parity7((uint32_t(lane) << 1) & G[r]) ? 0x7Fu : 0x81u) << (8 * r);
    }
    // ViterbiDecoder_Core::reset (viterbi_decoder_core.h:202-211)
    const uint32_t start_state = sch->start_state & 63u;
    uint32_t pack = VIT_NON_START | (VIT_NON_START << 16);
    if ((start_state >> 1) == uint32_t(lane)) pack = (start_state & 1u) ? (pack & 0x0000FFFFu) : (pack & 0xFFFF0000u);
    uint64_t renorm_acc = 0;
    bool near_sat = false;

    const uint32_t sel = (lane & 1) ? 0x3232u : 0x1010u;  // pick half (lane & 1) of the fetched pair and duplicate it
    const int src_a = lane >> 1, src_b = 16 + (lane >> 1);

    for (uint32_t t0 = 0; t0 < total_steps; t0 += 32) {
        // ---- fused depuncture (dab_viterbi_decoder.cpp:131-181): lane i fetches the kept symbols of step t0 + i
        uint32_t my_syms = 0;
        {
            const uint32_t t = t0 + uint32_t(lane);
            if (t < total_steps) {
                int k = 0;
#pragma unroll
                for (int i = 1; i < DAB_VIT_MAX_SEGMENTS; i++)
                    if (i < int(sch->n_seg) && t >= sch->seg[i].first_step) k = i;
                const DevSegment& sg = sch->seg[k];
                const uint32_t g = t - sg.first_step;
                const uint32_t q = g / sg.code_len, r = g - q * sg.code_len;
                const uint32_t prefix = ((r < 4 ? sg.prefix_lo : sg.prefix_hi) >> (8 * (r & 3))) & 0xFFu;
                const uint32_t cnt = (sg.counts >> (4 * r)) & 0xFu;
                const uint32_t idx = sg.soft_start + q * sg.period_syms + prefix;
#pragma unroll
                for (uint32_t j = 0; j < 4; j++)
                    if (j < cnt) my_syms |= uint32_t(uint8_t(my_soft[idx + j])) << (8 * j);
            }
        }
        const uint32_t n_here = min(32u, total_steps - t0);
        for (uint32_t i = 0; i < n_here; i++) {
            const uint32_t sym4 = __shfl_sync(0xFFFFFFFFu, my_syms, int(i));
            const uint32_t e = vabsdiff4_sum(table4, sym4);          // adds_epu16 never saturates: e <= 1020
            const uint32_t inv = (e > VIT_MAX_ERROR) ? 0u : (VIT_MAX_ERROR - e);  // subs_epu16(max_error, e)
            const uint32_t va = __shfl_sync(0xFFFFFFFFu, pack, src_a);
            const uint32_t vb = __shfl_sync(0xFFFFFFFFu, pack, src_b);
            const uint32_t m0m0 = __byte_perm(va, 0, sel), m1m1 = __byte_perm(vb, 0, sel);
            StepOut o = near_sat ? acs<true>(m0m0, m1m1, e, inv) : acs<false>(m0m0, m1m1, e, inv);
            pack = o.pack;
            const uint32_t w0 = __ballot_sync(0xFFFFFFFFu, o.d0);
            const uint32_t w1 = __ballot_sync(0xFFFFFFFFu, o.d1);
            if (lane == 0) win[long_mode ? i : (t0 + i)] = make_uint2(w0, w1);
            // lane 0 votes on the renormalisation test of metric[0]; every other lane on "could saturate next step".
            // (new[1] <= new[0] + 1020, so lane 0's high half is covered by the renormalisation test.)
            const uint32_t lo = pack & 0xFFFFu, hi = pack >> 16;
            const bool pred = (lane == 0) ? (lo >= VIT_RENORM) : (max(lo, hi) >= VIT_NEAR_SAT);
            const uint32_t vote = __ballot_sync(0xFFFFFFFFu, pred);
            near_sat = (vote >> 1) != 0u;
            if (vote & 1u) {  // renormalise (viterbi_decoder_avx_u16.h:138-170)
                uint32_t mn = min(lo, hi);
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
                pack -= mn | (mn << 16);
                renorm_acc += mn;
            }
        }
        if (long_mode) {
            __syncwarp();
            if (uint32_t(lane) < n_here) spill[t0 + uint32_t(lane)] = win[lane];
            __syncwarp();
        }
    }
    __syncwarp();

    // ---- DAB_Viterbi_Decoder::chainback (dab_viterbi_decoder.cpp:124-129): error = sum of renormalisations + metric[0]
    if (lane == 0) {
        if (path_error) path_error[job_index] = renorm_acc + uint64_t(pack & 0xFFFFu);
        if (job_status) job_status[job_index] = DAB_OK;
    }

    // ---- ViterbiDecoder_Core::chainback (viterbi_decoder_core.h:214-236) with ViterbiTracebackBuffer<7>
    const uint32_t n_bits = sch->n_out_bits;
    uint8_t* my_out = out + job.out_offset;
    uint32_t reg = (sch->end_state & 63u) << 2;
    if (!long_mode) {
        if (lane == 0) {
            for (int32_t j = int32_t(n_bits) - 1; j >= 0; j--) {
                const uint32_t bit = decision_bit(win[uint32_t(j) + 6u], reg >> 2);
                reg = (reg >> 1) | (bit << 7);
                if ((j & 7) == 0) my_out[j >> 3] = uint8_t(reg);
            }
        }
    } else {
        // page the decisions back window by window, newest first
        int64_t hi_step = int64_t(n_bits) + 5;  // decision index of bit n_bits-1
        while (hi_step >= 6) {
            const int64_t lo_step = max(int64_t(6), hi_step - int64_t(window_steps) + 1);
            for (int64_t s = lo_step + lane; s <= hi_step; s += 32) win[s - lo_step] = spill[s];
            __syncwarp();
            if (lane == 0) {
                for (int64_t s = hi_step; s >= lo_step; s--) {
                    const uint32_t bit = decision_bit(win[s - lo_step], reg >> 2);
                    reg = (reg >> 1) | (bit << 7);
                    const int64_t j = s - 6;
                    if ((j & 7) == 0) my_out[j >> 3] = uint8_t(reg);
                }
            }
            reg = __shfl_sync(0xFFFFFFFFu, reg, 0);
            __syncwarp();
            hi_step = lo_step - 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side

struct Viterbi {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::vector<DevSchedule> schedules;
    bool schedules_dirty = false;
    DeviceBuffer<DevSchedule> d_schedules;
    DeviceBuffer<dab_vit_job> d_jobs;
    DeviceBuffer<int8_t> d_soft;
    DeviceBuffer<uint8_t> d_out;
    DeviceBuffer<uint64_t> d_error;
    DeviceBuffer<int32_t> d_status;
    DeviceBuffer<uint2> d_scratch;
    int max_smem_optin = 0;
    int oneshot_slot = -1;  // schedule slot reused by dab_viterbi_decode_one
    uint64_t launches = 0;
    std::mutex mtx;
};

static int digest_schedule(const dab_vit_schedule* s, DevSchedule* d) {
    if (!s || s->n_seg == 0 || s->n_seg > DAB_VIT_MAX_SEGMENTS) return set_error(DAB_ERR_INVALID, "schedule needs 1..%d segments", DAB_VIT_MAX_SEGMENTS);
    memset(d, 0, sizeof(*d));
    uint32_t step = 0, soft = 0;
    for (uint32_t i = 0; i < s->n_seg; i++) {
        const dab_vit_segment& sg = s->seg[i];
        if (sg.code_len < 1 || sg.code_len > 8) return set_error(DAB_ERR_INVALID, "segment %u: code_len %u outside 1..8", i, sg.code_len);
        if (sg.n_out % 4 != 0) return set_error(DAB_ERR_INVALID, "segment %u: requested_output_symbols %u is not a multiple of the code rate", i, sg.n_out);
        DevSegment& o = d->seg[i];
        o.first_step = step;
        o.n_steps = sg.n_out / 4;
        o.soft_start = soft;
        o.code_len = sg.code_len;
        uint32_t prefix = 0;
        for (uint32_t r = 0; r < sg.code_len; r++) {
            if (sg.counts[r] > 4) return set_error(DAB_ERR_INVALID, "segment %u: puncture count %u > 4", i, unsigned(sg.counts[r]));
            o.counts |= uint32_t(sg.counts[r]) << (4 * r);
            if (r < 4) o.prefix_lo |= prefix << (8 * r); else o.prefix_hi |= prefix << (8 * (r - 4));
            prefix += sg.counts[r];
        }
        o.period_syms = prefix;
        const uint32_t full = o.n_steps / sg.code_len, rem = o.n_steps % sg.code_len;
        uint32_t used = full * prefix;
        for (uint32_t r = 0; r < rem; r++) used += sg.counts[r];
        soft += used;
        step += o.n_steps;
    }
    d->n_seg = s->n_seg;
    d->total_steps = step;
    d->n_out_bits = s->n_out_bytes * 8u;
    d->soft_symbols = soft;
    d->start_state = s->start_state;
    d->end_state = s->end_state;
    if (d->n_out_bits + 6u > step) return set_error(DAB_ERR_TRACEBACK, "chainback of %u bits needs %u trellis steps, schedule has %u", d->n_out_bits, d->n_out_bits + 6u, step);
    return DAB_OK;
}

static int upload_schedules(Viterbi* v) {
    if (!v->schedules_dirty) return DAB_OK;
    DAB_CUDA_CHECK(v->d_schedules.reserve(std::max<size_t>(v->schedules.size(), 16)));
    DAB_CUDA_CHECK(cudaMemcpyAsync(v->d_schedules.ptr, v->schedules.data(), v->schedules.size() * sizeof(DevSchedule), cudaMemcpyHostToDevice, v->stream));
    // the host vector may be reallocated by a later add_schedule; make the copy complete first
    DAB_CUDA_CHECK(cudaStreamSynchronize(v->stream));
    v->schedules_dirty = false;
    return DAB_OK;
}

// window (steps kept in shared memory per warp) for a launch whose longest trellis has max_steps steps
static uint32_t pick_window(const Viterbi* v, uint32_t max_steps, size_t* smem_bytes) {
    const size_t budget = size_t(v->max_smem_optin) - 1024;
    const size_t per_warp_budget = budget / VIT_WARPS_PER_CTA - sizeof(DevSchedule);
    uint32_t cap = uint32_t(per_warp_budget / sizeof(uint2)) & ~31u;
    uint32_t want = (max_steps + 31u) & ~31u;
    uint32_t window = std::min(cap, std::max(want, 32u));
    *smem_bytes = size_t(VIT_WARPS_PER_CTA) * (size_t(window) * sizeof(uint2) + sizeof(DevSchedule));
    return window;
}

static int launch(Viterbi* v, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* d_jobs, int n_jobs, uint32_t max_steps,
                  uint8_t* d_out, size_t out_bytes, uint64_t* d_error, int32_t* d_status) {
    if (n_jobs <= 0) return DAB_OK;
    int rc = upload_schedules(v);
    if (rc != DAB_OK) return rc;
    size_t smem = 0;
    const uint32_t window = pick_window(v, max_steps, &smem);
    uint32_t scratch_steps = 0;
    if (max_steps > window) {
        scratch_steps = (max_steps + 31u) & ~31u;
        DAB_CUDA_CHECK(v->d_scratch.reserve(size_t(scratch_steps) * size_t(n_jobs)));
    }
    DAB_CUDA_CHECK(cudaFuncSetAttribute(viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v->max_smem_optin));
    const int grid = (n_jobs + VIT_WARPS_PER_CTA - 1) / VIT_WARPS_PER_CTA;
    viterbi_kernel<<<grid, VIT_WARPS_PER_CTA * 32, smem, v->stream>>>(d_soft, soft_bytes, d_jobs, n_jobs, v->d_schedules.ptr,
                                                                     int(v->schedules.size()), d_out, out_bytes, d_error, d_status,
                                                                     v->d_scratch.ptr, scratch_steps, window);
    v->launches++;
    DAB_CUDA_CHECK(cudaGetLastError());
    return DAB_OK;
}

static uint32_t max_steps_of(const Viterbi* v, const dab_vit_job* jobs, int n_jobs, int* bad) {
    uint32_t m = 0;
    for (int i = 0; i < n_jobs; i++) {
        if (jobs[i].schedule >= v->schedules.size()) { *bad = i; continue; }
        m = std::max(m, v->schedules[jobs[i].schedule].total_steps);
    }
    return m;
}

}  // namespace dabb200

using namespace dabb200;

extern "C" {

dab_viterbi* dab_viterbi_create(int device, int* status) {
    int rc = select_device(device);
    if (rc != DAB_OK) { if (status) *status = rc; return nullptr; }
    auto* v = new Viterbi();
    v->device = device;
    if (cudaStreamCreateWithFlags(&v->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaDeviceGetAttribute(&v->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) {
        if (status) *status = set_error(DAB_ERR_CUDA, "viterbi create: %s", cudaGetErrorString(cudaGetLastError()));
        delete v;
        return nullptr;
    }
    v->stream = v->own_stream;
    if (status) *status = DAB_OK;
    return reinterpret_cast<dab_viterbi*>(v);
}

void dab_viterbi_destroy(dab_viterbi* h) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return;
    cudaSetDevice(v->device);
    cudaStreamSynchronize(v->stream);
    if (v->own_stream) cudaStreamDestroy(v->own_stream);
    delete v;
}

int dab_viterbi_set_cuda_stream(dab_viterbi* h, void* cuda_stream) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    v->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : v->own_stream;
    return DAB_OK;
}

int dab_viterbi_add_schedule(dab_viterbi* h, const dab_vit_schedule* s) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    DevSchedule d;
    int rc = digest_schedule(s, &d);
    if (rc != DAB_OK) return rc;
    std::lock_guard<std::mutex> lock(v->mtx);
    if (v->schedules.size() >= DAB_VIT_MAX_SCHEDULES) return set_error(DAB_ERR_CAPACITY, "more than %d schedules", DAB_VIT_MAX_SCHEDULES);
    v->schedules.push_back(d);
    v->schedules_dirty = true;
    return int(v->schedules.size()) - 1;
}

int64_t dab_viterbi_schedule_soft_symbols(const dab_vit_schedule* s) {
    DevSchedule d;
    int rc = digest_schedule(s, &d);
    if (rc != DAB_OK && rc != DAB_ERR_TRACEBACK) return rc;
    return int64_t(d.soft_symbols);
}

int dab_viterbi_decode_jobs_device(dab_viterbi* h, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* d_jobs, int n_jobs,
                                   uint32_t max_steps, uint8_t* d_out, size_t out_bytes, uint64_t* d_path_error, int32_t* d_job_status) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    if (n_jobs < 0 || (n_jobs > 0 && (!d_soft || !d_jobs || !d_out))) return set_error(DAB_ERR_INVALID, "null buffer");
    std::lock_guard<std::mutex> lock(v->mtx);
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    return launch(v, d_soft, soft_bytes, d_jobs, n_jobs, max_steps, d_out, out_bytes, d_path_error, d_job_status);
}

int dab_viterbi_decode_batch_device(dab_viterbi* h, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* jobs, int n_jobs,
                                    uint8_t* d_out, size_t out_bytes, uint64_t* d_path_error, int32_t* d_job_status) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    if (n_jobs < 0 || (n_jobs > 0 && (!d_soft || !jobs || !d_out))) return set_error(DAB_ERR_INVALID, "null buffer");
    if (n_jobs == 0) return DAB_OK;
    std::lock_guard<std::mutex> lock(v->mtx);
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    int bad = -1;
    const uint32_t max_steps = max_steps_of(v, jobs, n_jobs, &bad);
    if (bad >= 0) return set_error(DAB_ERR_INVALID, "job %d names unknown schedule %u", bad, jobs[bad].schedule);
    DAB_CUDA_CHECK(v->d_jobs.reserve(size_t(n_jobs)));
    DAB_CUDA_CHECK(cudaMemcpyAsync(v->d_jobs.ptr, jobs, size_t(n_jobs) * sizeof(dab_vit_job), cudaMemcpyHostToDevice, v->stream));
    int rc = launch(v, d_soft, soft_bytes, v->d_jobs.ptr, n_jobs, max_steps, d_out, out_bytes, d_path_error, d_job_status);
    // `jobs` is caller memory: make sure the staged copy has left it before returning
    DAB_CUDA_CHECK(cudaStreamSynchronize(v->stream));
    return rc;
}

int dab_viterbi_decode_batch(dab_viterbi* h, const int8_t* soft, size_t soft_bytes, const dab_vit_job* jobs, int n_jobs, uint8_t* out,
                             size_t out_bytes, uint64_t* path_error, int32_t* job_status) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    if (n_jobs < 0 || (n_jobs > 0 && (!soft || !jobs || !out))) return set_error(DAB_ERR_INVALID, "null buffer");
    if (n_jobs == 0) return DAB_OK;
    std::lock_guard<std::mutex> lock(v->mtx);
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    int bad = -1;
    const uint32_t max_steps = max_steps_of(v, jobs, n_jobs, &bad);
    if (bad >= 0) return set_error(DAB_ERR_INVALID, "job %d names unknown schedule %u", bad, jobs[bad].schedule);
    DAB_CUDA_CHECK(v->d_jobs.reserve(size_t(n_jobs)));
    DAB_CUDA_CHECK(v->d_soft.reserve(soft_bytes));
    DAB_CUDA_CHECK(v->d_out.reserve(out_bytes));
    DAB_CUDA_CHECK(v->d_error.reserve(size_t(n_jobs)));
    DAB_CUDA_CHECK(v->d_status.reserve(size_t(n_jobs)));
    DAB_CUDA_CHECK(cudaMemcpyAsync(v->d_jobs.ptr, jobs, size_t(n_jobs) * sizeof(dab_vit_job), cudaMemcpyHostToDevice, v->stream));
    DAB_CUDA_CHECK(cudaMemcpyAsync(v->d_soft.ptr, soft, soft_bytes, cudaMemcpyHostToDevice, v->stream));
    DAB_CUDA_CHECK(cudaMemsetAsync(v->d_out.ptr, 0, out_bytes, v->stream));
    int rc = launch(v, v->d_soft.ptr, soft_bytes, v->d_jobs.ptr, n_jobs, max_steps, v->d_out.ptr, out_bytes, v->d_error.ptr, v->d_status.ptr);
    if (rc != DAB_OK) return rc;
    DAB_CUDA_CHECK(cudaMemcpyAsync(out, v->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, v->stream));
    if (path_error) DAB_CUDA_CHECK(cudaMemcpyAsync(path_error, v->d_error.ptr, size_t(n_jobs) * sizeof(uint64_t), cudaMemcpyDeviceToHost, v->stream));
    std::vector<int32_t> st(static_cast<size_t>(n_jobs));
    DAB_CUDA_CHECK(cudaMemcpyAsync(st.data(), v->d_status.ptr, size_t(n_jobs) * sizeof(int32_t), cudaMemcpyDeviceToHost, v->stream));
    DAB_CUDA_CHECK(cudaStreamSynchronize(v->stream));
    int first_bad = DAB_OK;
    for (int i = 0; i < n_jobs; i++) {
        if (job_status) job_status[i] = st[size_t(i)];
        if (st[size_t(i)] != DAB_OK && first_bad == DAB_OK) first_bad = set_error(st[size_t(i)], "job %d failed with status %d", i, st[size_t(i)]);
    }
    return first_bad;
}

int dab_viterbi_decode_one(dab_viterbi* h, const dab_vit_schedule* s, const int8_t* soft, size_t n_soft, uint8_t* out, uint64_t* path_error) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    if (!s || !soft || !out) return set_error(DAB_ERR_INVALID, "null argument");
    DevSchedule d;
    int rc = digest_schedule(s, &d);
    if (rc != DAB_OK) return rc;
    {
        std::lock_guard<std::mutex> lock(v->mtx);
        if (v->oneshot_slot < 0) {
            if (v->schedules.size() >= DAB_VIT_MAX_SCHEDULES) return set_error(DAB_ERR_CAPACITY, "no schedule slot left");
            v->schedules.push_back(d);
            v->oneshot_slot = int(v->schedules.size()) - 1;
        } else {
            v->schedules[size_t(v->oneshot_slot)] = d;
        }
        v->schedules_dirty = true;
    }
    dab_vit_job job;
    job.schedule = uint32_t(v->oneshot_slot);
    job.n_soft = uint32_t(n_soft);
    job.soft_offset = 0;
    job.out_offset = 0;
    int32_t st = 0;
    rc = dab_viterbi_decode_batch(h, soft, n_soft, &job, 1, out, s->n_out_bytes, path_error, &st);
    return rc;
}

int dab_viterbi_sync(dab_viterbi* h) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    DAB_CUDA_CHECK(cudaStreamSynchronize(v->stream));
    return DAB_OK;
}

uint64_t dab_viterbi_kernel_launches(const dab_viterbi* h) {
    auto* v = reinterpret_cast<const Viterbi*>(h);
    return v ? v->launches : 0;
}

}  // extern "C"
