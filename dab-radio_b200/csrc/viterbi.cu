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
#include <memory>
#include <mutex>
#include <vector>

#include "viterbi_core.cuh"
#include "viterbi_lanes.cuh"

namespace dabb200 {

// the plain batch decoder's view of a job: punctured symbols contiguous at soft + soft_offset, bytes to out + out_offset
struct PlainView {
    const int8_t* soft;
    uint8_t* out;
    __device__ __forceinline__ uint32_t fetch(uint32_t idx) const { return uint32_t(uint8_t(soft[idx])); }
    __device__ __forceinline__ const int8_t* soft_base() const { return soft; }
    __device__ __forceinline__ void store(uint32_t byte, uint32_t value) { out[byte] = uint8_t(value); }
};

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
    uint2* spill = (sch->total_steps > window_steps) ? (scratch + size_t(job_index) * scratch_steps_per_job) : nullptr;
    PlainView view{soft + job.soft_offset, out + job.out_offset};
    const uint64_t err = viterbi_trellis(sch, view, win, spill, window_steps, lane);
    if (lane == 0) {
        if (path_error) path_error[job_index] = err;
        if (job_status) job_status[job_index] = DAB_OK;
    }
}

// The bulk form: one trellis per thread (viterbi_lanes.cuh).  Warp w runs group groups[w]: the 32 jobs order[32 g + lane]
// (order == nullptr: job 32 g + lane).  The host orders the jobs by schedule so that the 32 trellises of a group walk the same
// puncturing schedule in lock step, and launches the longest groups first (vitl_plan).  Decision rows of warp w start at row
// warp_row[w] of the scratch.
__global__ void __maxnreg__(VITL_MAX_REGS)
viterbi_lanes_kernel(const int8_t* __restrict__ soft, size_t soft_bytes, const dab_vit_job* __restrict__ jobs, int n_jobs,
                     const int32_t* __restrict__ order, const int32_t* __restrict__ groups, const unsigned long long* __restrict__ warp_row,
                     const DevSchedule* __restrict__ schedules, int n_schedules, uint8_t* __restrict__ out, size_t out_bytes,
                     uint64_t* __restrict__ path_error, int32_t* __restrict__ job_status, uint2* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (VITL_THREADS / 32) + (threadIdx.x >> 5);
    const uint32_t rows = uint32_t(warp_row[w + 1] - warp_row[w]);
    const long long slot = (long long)(groups[w]) * 32 + lane;
    int job_index = -1;
    if (slot < n_jobs) job_index = order ? order[slot] : int(slot);
    int status = DAB_OK;
    dab_vit_job job{};
    const DevSchedule* sch = schedules;
    if (job_index >= 0) {
        job = jobs[job_index];
        if (job.schedule >= uint32_t(n_schedules)) status = DAB_ERR_INVALID;
        else {
            sch = schedules + job.schedule;
            if (sch->soft_symbols > job.n_soft || job.soft_offset + sch->soft_symbols > soft_bytes) status = DAB_ERR_UNDERRUN;
            else if (sch->n_out_bits + 6u > sch->total_steps) status = DAB_ERR_TRACEBACK;
            else if (job.out_offset + sch->n_out_bits / 8u > out_bytes) status = DAB_ERR_CAPACITY;
            else if (sch->total_steps > rows) status = DAB_ERR_CAPACITY;
        }
    }
    const bool active = job_index >= 0 && status == DAB_OK;
    PlainView view{soft + job.soft_offset, out + job.out_offset};
    const uint64_t err = viterbi_lane_trellis(sch, view, scratch + size_t(warp_row[w]) * 32u, lane, active);
    if (job_index >= 0) {
        if (path_error) path_error[job_index] = active ? err : 0;
        if (job_status) job_status[job_index] = status;
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
    // job orders / warp plans of the bulk form: slot 0 is rebuilt by every call that brings its own job list, the others belong
    // to dab_viterbi_prepare_jobs
    struct Prepared {
        bool used = false;
        int n_jobs = 0;
        uint32_t max_steps = 0;
        bool has_order = false;
        int n_warps = 0;
        unsigned long long rows = 0;
        DeviceBuffer<dab_vit_job> d_jobs;
        DeviceBuffer<int32_t> d_order, d_groups;
        DeviceBuffer<unsigned long long> d_warp_row;
    };
    std::vector<std::unique_ptr<Prepared>> prepared;
    int max_smem_optin = 0;
    int oneshot_slot = -1;  // schedule slot reused by dab_viterbi_decode_one
    uint64_t launches = 0;
    std::recursive_mutex mtx;   // recursive: dab_viterbi_decode_one holds it across its schedule slot write and the decode
};

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

static int launch_warps(Viterbi* v, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* d_jobs, int n_jobs, uint32_t max_steps,
                        uint8_t* d_out, size_t out_bytes, uint64_t* d_error, int32_t* d_status) {
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


// Order the jobs (when the host can see them) and plan the warps of the bulk form into `p`.
static int plan_lanes(Viterbi* v, Viterbi::Prepared* p, const dab_vit_job* host_jobs, int n_jobs, uint32_t max_steps) {
    const int n_groups = (n_jobs + 31) / 32;
    std::vector<uint32_t> cost(static_cast<size_t>(n_groups), max_steps);
    p->has_order = false;
    if (host_jobs) {
        auto steps_of = [&](int32_t j) { return host_jobs[j].schedule < v->schedules.size() ? v->schedules[host_jobs[j].schedule].total_steps : 0u; };
        bool mixed = false;
        for (int i = 1; i < n_jobs && !mixed; i++) mixed = host_jobs[i].schedule != host_jobs[0].schedule;
        std::vector<int32_t> order(static_cast<size_t>(n_jobs));
        for (int i = 0; i < n_jobs; i++) order[size_t(i)] = i;
        if (mixed) {
            // longest schedules first, equal schedules together (stable: neighbouring jobs stay neighbours)
            std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
                const uint32_t sa = steps_of(a), sb = steps_of(b);
                return sa != sb ? sa > sb : host_jobs[a].schedule < host_jobs[b].schedule;
            });
            DAB_CUDA_CHECK(p->d_order.reserve(size_t(n_jobs)));
            DAB_CUDA_CHECK(cudaMemcpyAsync(p->d_order.ptr, order.data(), size_t(n_jobs) * sizeof(int32_t), cudaMemcpyHostToDevice, v->stream));
            p->has_order = true;
        }
        for (int g = 0; g < n_groups; g++) {
            uint32_t c = 0;
            for (int i = 32 * g; i < std::min(n_jobs, 32 * g + 32); i++) c = std::max(c, steps_of(order[size_t(i)]));
            cost[size_t(g)] = c;
        }
    }
    VitlPlan plan;
    vitl_plan(cost, plan);
    p->n_jobs = n_jobs;
    p->max_steps = max_steps;
    p->n_warps = plan.n_warps();
    p->rows = plan.rows();
    DAB_CUDA_CHECK(p->d_groups.reserve(std::max<size_t>(plan.groups.size(), 1)));
    DAB_CUDA_CHECK(p->d_warp_row.reserve(plan.warp_row.size()));
    DAB_CUDA_CHECK(cudaMemcpyAsync(p->d_groups.ptr, plan.groups.data(), plan.groups.size() * sizeof(int32_t), cudaMemcpyHostToDevice, v->stream));
    DAB_CUDA_CHECK(cudaMemcpyAsync(p->d_warp_row.ptr, plan.warp_row.data(), plan.warp_row.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, v->stream));
    DAB_CUDA_CHECK(cudaStreamSynchronize(v->stream));   // the sources are local vectors
    return DAB_OK;
}

static int launch_lanes(Viterbi* v, const Viterbi::Prepared* p, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* d_jobs,
                        uint8_t* d_out, size_t out_bytes, uint64_t* d_error, int32_t* d_status) {
    DAB_CUDA_CHECK(v->d_scratch.reserve(size_t(p->rows) * 32u));
    viterbi_lanes_kernel<<<p->n_warps, VITL_THREADS, 0, v->stream>>>(d_soft, soft_bytes, d_jobs, p->n_jobs, p->has_order ? p->d_order.ptr : nullptr,
                                                                    p->d_groups.ptr, p->d_warp_row.ptr, v->d_schedules.ptr,
                                                                    int(v->schedules.size()), d_out, out_bytes, d_error, d_status, v->d_scratch.ptr);
    v->launches++;
    DAB_CUDA_CHECK(cudaGetLastError());
    return DAB_OK;
}

static int launch(Viterbi* v, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* d_jobs, const dab_vit_job* host_jobs, int n_jobs,
                  uint32_t max_steps, uint8_t* d_out, size_t out_bytes, uint64_t* d_error, int32_t* d_status) {
    if (n_jobs <= 0) return DAB_OK;
    int rc = upload_schedules(v);
    if (rc != DAB_OK) return rc;
    if (vitl_use_lanes(n_jobs)) {
        if (v->prepared.empty()) v->prepared.emplace_back(new Viterbi::Prepared());
        Viterbi::Prepared* p = v->prepared[0].get();
        // a device-resident job list (host_jobs == nullptr) of the same shape as last time keeps its plan
        if (host_jobs || !p->used || p->n_jobs != n_jobs || p->max_steps != max_steps || p->has_order) {
            rc = plan_lanes(v, p, host_jobs, n_jobs, max_steps);
            if (rc != DAB_OK) return rc;
            p->used = true;
        }
        return launch_lanes(v, p, d_soft, soft_bytes, d_jobs, d_out, out_bytes, d_error, d_status);
    }
    return launch_warps(v, d_soft, soft_bytes, d_jobs, n_jobs, max_steps, d_out, out_bytes, d_error, d_status);
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
    std::lock_guard<std::recursive_mutex> lock(v->mtx);
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
    std::lock_guard<std::recursive_mutex> lock(v->mtx);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    return launch(v, d_soft, soft_bytes, d_jobs, nullptr, n_jobs, max_steps, d_out, out_bytes, d_path_error, d_job_status);
}

int dab_viterbi_decode_batch_device(dab_viterbi* h, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* jobs, int n_jobs,
                                    uint8_t* d_out, size_t out_bytes, uint64_t* d_path_error, int32_t* d_job_status) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    if (n_jobs < 0 || (n_jobs > 0 && (!d_soft || !jobs || !d_out))) return set_error(DAB_ERR_INVALID, "null buffer");
    if (n_jobs == 0) return DAB_OK;
    std::lock_guard<std::recursive_mutex> lock(v->mtx);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    int bad = -1;
    const uint32_t max_steps = max_steps_of(v, jobs, n_jobs, &bad);
    if (bad >= 0) return set_error(DAB_ERR_INVALID, "job %d names unknown schedule %u", bad, jobs[bad].schedule);
    DAB_CUDA_CHECK(v->d_jobs.reserve(size_t(n_jobs)));
    DAB_CUDA_CHECK(cudaMemcpyAsync(v->d_jobs.ptr, jobs, size_t(n_jobs) * sizeof(dab_vit_job), cudaMemcpyHostToDevice, v->stream));
    int rc = launch(v, d_soft, soft_bytes, v->d_jobs.ptr, jobs, n_jobs, max_steps, d_out, out_bytes, d_path_error, d_job_status);
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
    std::lock_guard<std::recursive_mutex> lock(v->mtx);
    DeviceGuard device_guard;
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
    int rc = launch(v, v->d_soft.ptr, soft_bytes, v->d_jobs.ptr, jobs, n_jobs, max_steps, v->d_out.ptr, out_bytes, v->d_error.ptr, v->d_status.ptr);
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
    // the one-shot schedule slot is shared by every decode_one on this handle: hold the lock until the decode that reads it is done
    std::lock_guard<std::recursive_mutex> lock(v->mtx);
    {
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

int dab_viterbi_prepare_jobs(dab_viterbi* h, const dab_vit_job* jobs, int n_jobs) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    if (!jobs || n_jobs <= 0) return set_error(DAB_ERR_INVALID, "empty job list");
    std::lock_guard<std::recursive_mutex> lock(v->mtx);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    int bad = -1;
    const uint32_t max_steps = max_steps_of(v, jobs, n_jobs, &bad);
    if (bad >= 0) return set_error(DAB_ERR_INVALID, "job %d names unknown schedule %u", bad, jobs[bad].schedule);
    int rc = upload_schedules(v);
    if (rc != DAB_OK) return rc;
    if (v->prepared.empty()) v->prepared.emplace_back(new Viterbi::Prepared());   // slot 0 stays with the ad-hoc calls
    size_t id = 1;
    while (id < v->prepared.size() && v->prepared[id]->used) id++;
    if (id == v->prepared.size()) v->prepared.emplace_back(new Viterbi::Prepared());
    Viterbi::Prepared* p = v->prepared[id].get();
    DAB_CUDA_CHECK(p->d_jobs.reserve(size_t(n_jobs)));
    DAB_CUDA_CHECK(cudaMemcpyAsync(p->d_jobs.ptr, jobs, size_t(n_jobs) * sizeof(dab_vit_job), cudaMemcpyHostToDevice, v->stream));
    rc = plan_lanes(v, p, jobs, n_jobs, max_steps);   // synchronises: `jobs` is free again on return
    if (rc != DAB_OK) return rc;
    p->used = true;
    return int(id);
}

int dab_viterbi_decode_prepared(dab_viterbi* h, int plan, const int8_t* d_soft, size_t soft_bytes, uint8_t* d_out, size_t out_bytes,
                                uint64_t* d_path_error, int32_t* d_job_status) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    if (!d_soft || !d_out) return set_error(DAB_ERR_INVALID, "null buffer");
    std::lock_guard<std::recursive_mutex> lock(v->mtx);
    if (plan < 1 || size_t(plan) >= v->prepared.size() || !v->prepared[size_t(plan)]->used) return set_error(DAB_ERR_INVALID, "unknown job plan %d", plan);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    Viterbi::Prepared* p = v->prepared[size_t(plan)].get();
    int rc = upload_schedules(v);
    if (rc != DAB_OK) return rc;
    if (vitl_use_lanes(p->n_jobs)) return launch_lanes(v, p, d_soft, soft_bytes, p->d_jobs.ptr, d_out, out_bytes, d_path_error, d_job_status);
    return launch_warps(v, d_soft, soft_bytes, p->d_jobs.ptr, p->n_jobs, p->max_steps, d_out, out_bytes, d_path_error, d_job_status);
}

int dab_viterbi_release_jobs(dab_viterbi* h, int plan) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    std::lock_guard<std::recursive_mutex> lock(v->mtx);
    if (plan < 1 || size_t(plan) >= v->prepared.size() || !v->prepared[size_t(plan)]->used) return set_error(DAB_ERR_INVALID, "unknown job plan %d", plan);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    DAB_CUDA_CHECK(cudaStreamSynchronize(v->stream));
    v->prepared[size_t(plan)].reset(new Viterbi::Prepared());
    return DAB_OK;
}

int dab_viterbi_sync(dab_viterbi* h) {
    auto* v = reinterpret_cast<Viterbi*>(h);
    if (!v) return set_error(DAB_ERR_INVALID, "null handle");
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(v->device));
    DAB_CUDA_CHECK(cudaStreamSynchronize(v->stream));
    return DAB_OK;
}

uint64_t dab_viterbi_kernel_launches(const dab_viterbi* h) {
    auto* v = reinterpret_cast<const Viterbi*>(h);
    return v ? v->launches : 0;
}

}  // extern "C"
