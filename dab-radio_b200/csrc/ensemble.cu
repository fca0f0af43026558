// Ensemble decoder for sm_100a: one OFDM frame of soft bits per stream -> FIBs and sub-channel bytes, batched over streams.
//
// Replaces, per stream (reference file:line):
//   BasicRadio::Process          src/basic_radio/basic_radio.cpp:42-63       split of the frame into FIC + MSC
//   BasicFICRunner::Process      src/basic_radio/basic_fic_runner.cpp:36-51  one FIB group per CIF
//   FIC_Decoder::DecodeFIBGroup  src/dab/fic/fic_decoder.cpp:53-116          Viterbi, energy dispersal, CRC16 per FIB
//   MSC_Decoder::DecodeCIF       src/dab/msc/msc_decoder.cpp:46-170          slice, CIF de-interleave, EEP/UEP Viterbi, dispersal
//   CIF_Deinterleaver            src/dab/msc/cif_deinterleaver.cpp:9-70      16-CIF time de-interleaver
//   AdditiveScrambler            src/dab/algorithms/additive_scrambler.h:10-35
//   CRC_Calculator<uint16_t>     src/dab/algorithms/crc.h:24-33 with fic_decoder.cpp:20-33
//   protection tables            src/dab/constants/subchannel_protection_tables.h:21-154
//
// Data flow per decode call (all in HBM, four launches):
//   1. ens_push_kernel      each CIF of the new frame is written into a per-stream ring as 16 planes (plane r = the soft bits
//                           with index = r mod 16), so that what one output CIF needs from an older CIF is contiguous.
//   2. ens_deint_kernel     output CIF n = plane r of CIF n - (15 - offset[r]) for r = 0..15, re-interleaved through shared
//                           memory into natural order.  De-interleaving is independent of the sub-channel layout, so the whole
//                           CIF is done at once and the reference's per-sub-channel rings (one per MSC_Decoder) are replaced
//                           by one ring per stream + a per-sub-channel "CIFs consumed" counter that gates the output.
//   3. ens_viterbi_kernel   one warp per trellis (FIB group or sub-channel CIF): the decoder core of viterbi_core.cuh with the
//                           energy-dispersal XOR fused into the traceback's byte store and the FIB CRC16 behind it.
//   4. ens_commit_kernel    advances the per-stream CIF count and the per-sub-channel counters.
#include <algorithm>
#include <mutex>
#include <type_traits>
#include <vector>

#include "viterbi_core.cuh"
#include "viterbi_lanes.cuh"

namespace dabb200 {

constexpr int ENS_DEINT_DEPTH = 16;                   // cif_deinterleaver.cpp:8
__constant__ int c_cif_offsets[ENS_DEINT_DEPTH] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};  // :9-11, ETSI table 21
constexpr int ENS_CHUNK = 256;                        // plane bytes per CTA in the push / de-interleave kernels

struct SubDesc {
    int32_t start_cu;   // start_address (capacity units of 64 soft bits)
    int32_t length_cu;
    int32_t schedule;   // index into the schedule table, -1 = slot unused
    int32_t overflow;   // start + length beyond the CIF: never consumed, never decoded (msc_decoder.cpp:49-54)
};

struct EnsGeom {
    int n_streams, nb_cifs, max_subs;
    int nb_fic_bits, nb_fib_cif_bits, nb_cif_bits, nb_fibs_per_cif;
    int fic_enabled;        // FIB group size is the one the reference decodes (fic_decoder.cpp:68-75)
    int plane_len;          // nb_cif_bits / 16
    int plane_pitch;        // plane_len rounded up to 16
    int ring_rows;          // 16 + nb_cifs, so a frame's CIFs can all be pushed before any is de-interleaved
    int cif_pitch;          // de-interleaved CIF pitch (nb_cif_bits rounded up to 16)
    int fib_group_bytes;    // nb_fib_cif_bits / 24
    int msc_cif_bytes;      // nb_cif_bits / 8
    int aligned16;          // the caller's frame rows allow 16-byte loads
};

__device__ __forceinline__ bool stream_has_frame(const int32_t* frames_in_call, int slot, int s) {
    return frames_in_call == nullptr || frames_in_call[s] > slot;
}

// ---- 1. push: CIF c of the new frame -> ring row (cif_count + c) % ring_rows, 16 planes
__global__ void __launch_bounds__(ENS_CHUNK)
ens_push_kernel(EnsGeom g, const int8_t* __restrict__ bits, size_t stream_stride, const int32_t* __restrict__ frames_in_call, int slot,
                const unsigned long long* __restrict__ cif_count, int8_t* __restrict__ ring) {
    __shared__ __align__(16) uint8_t tile[16][ENS_CHUNK];
    const int s = blockIdx.z, c = blockIdx.y, t = threadIdx.x;
    if (!stream_has_frame(frames_in_call, slot, s)) return;
    const int p0 = blockIdx.x * ENS_CHUNK, pos = p0 + t;
    const int8_t* src = bits + size_t(s) * stream_stride + size_t(g.nb_fic_bits) + size_t(c) * size_t(g.nb_cif_bits);
    if (pos < g.plane_len) {
        uint32_t w[4];
        if (g.aligned16) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + 16 * size_t(pos));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                w[i] = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) w[i] |= uint32_t(uint8_t(src[16 * size_t(pos) + 4 * i + b])) << (8 * b);
            }
        }
#pragma unroll
        for (int r = 0; r < 16; r++) tile[r][t] = uint8_t(w[r >> 2] >> (8 * (r & 3)));
    }
    __syncthreads();
    const int row = int((cif_count[s] + uint64_t(c)) % uint64_t(g.ring_rows));
    const int plane = t >> 4, seg = (t & 15) * 16;
    if (p0 + seg < g.plane_len) {   // plane_pitch is a multiple of 16: the tail of the last segment lands in the padding
        int8_t* dst = ring + ((size_t(s) * g.ring_rows + row) * 16 + plane) * size_t(g.plane_pitch) + p0 + seg;
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&tile[plane][seg]);
    }
}

// ---- 2. de-interleave: CIF_Deinterleaver::Deinterleave (cif_deinterleaver.cpp:37-70) for the whole CIF
__global__ void __launch_bounds__(ENS_CHUNK)
ens_deint_kernel(EnsGeom g, const int32_t* __restrict__ frames_in_call, int slot, const unsigned long long* __restrict__ cif_count,
                 const int8_t* __restrict__ ring, int8_t* __restrict__ deint) {
    __shared__ __align__(16) uint8_t tile[16][ENS_CHUNK];
    const int s = blockIdx.z, c = blockIdx.y, t = threadIdx.x;
    if (!stream_has_frame(frames_in_call, slot, s)) return;
    const int p0 = blockIdx.x * ENS_CHUNK;
    const long long newest = (long long)(cif_count[s]) + c;   // absolute index of the CIF being reconstructed
    {
        const int plane = t >> 4, seg = (t & 15) * 16;
        const long long src_cif = newest - (15 - c_cif_offsets[plane]);   // BUFFER_LOOKUP[15 - offset] (:49-66)
        uint4 v = make_uint4(0, 0, 0, 0);
        if (src_cif >= 0 && p0 + seg < g.plane_len) {
            const int row = int(src_cif % g.ring_rows);
            v = *reinterpret_cast<const uint4*>(ring + ((size_t(s) * g.ring_rows + row) * 16 + plane) * size_t(g.plane_pitch) + p0 + seg);
        }
        *reinterpret_cast<uint4*>(&tile[plane][seg]) = v;
    }
    __syncthreads();
    const int pos = p0 + t;
    if (pos < g.plane_len) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int r = 0; r < 16; r++) w[r >> 2] |= uint32_t(tile[r][t]) << (8 * (r & 3));
        int8_t* dst = deint + (size_t(s) * g.nb_cifs + c) * size_t(g.cif_pitch) + 16 * size_t(pos);
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- 3. Viterbi + energy dispersal + FIB CRC
struct EnsView {
    const int8_t* soft;
    uint8_t* out;
    const uint8_t* prbs;   // AdditiveScrambler byte sequence for syncword 0xFFFF
    __device__ __forceinline__ uint32_t fetch(uint32_t idx) const { return uint32_t(uint8_t(soft[idx])); }
    __device__ __forceinline__ const int8_t* soft_base() const { return soft; }
    __device__ __forceinline__ void store(uint32_t byte, uint32_t value) { out[byte] = uint8_t(value ^ prbs[byte]); }
};

// CRC_Calculator<uint16_t>::Process (crc.h:24-33) with G = 0x1021, initial value 0xFFFF, final xor 0xFFFF (fic_decoder.cpp:20-33)
__device__ __forceinline__ uint32_t crc16_fib(const uint8_t* x, int n) {
    uint32_t crc = 0xFFFFu;
    for (int i = 0; i < n; i++) {
        crc ^= uint32_t(x[i]) << 8;
#pragma unroll
        for (int j = 0; j < 8; j++) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) : (crc << 1);
        crc &= 0xFFFFu;
    }
    return crc ^ 0xFFFFu;
}

struct EnsOut {
    uint8_t* fib_bytes;
    uint8_t* fib_valid;
    unsigned long long* fic_error;
    uint8_t* msc_bytes;
    int32_t* msc_nbytes;
    unsigned long long* msc_error;
};

__global__ void __launch_bounds__(VIT_WARPS_PER_CTA * 32)
ens_viterbi_kernel(EnsGeom g, const int8_t* __restrict__ bits, size_t stream_stride, const int32_t* __restrict__ frames_in_call, int slot,
                   const int8_t* __restrict__ deint, const SubDesc* __restrict__ subs, const int32_t* __restrict__ stored,
                   const DevSchedule* __restrict__ schedules, const uint8_t* __restrict__ prbs, EnsOut o, int jobs_per_cif,
                   const int32_t* __restrict__ long_rank, uint2* __restrict__ scratch, uint32_t scratch_steps, uint32_t window_steps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long job = (long long)(blockIdx.x) * VIT_WARPS_PER_CTA + warp;
    const long long total = (long long)(jobs_per_cif) * g.nb_cifs * g.n_streams;
    if (job >= total) return;
    // stream fastest, then CIF, then slot: neighbouring warps run the same kind of trellis
    const int s = int(job % g.n_streams);
    const int c = int((job / g.n_streams) % g.nb_cifs);
    const int slot_k = int(job / (size_t(g.n_streams) * g.nb_cifs));
    if (!stream_has_frame(frames_in_call, slot, s)) return;

    const size_t per_warp = size_t(window_steps) * sizeof(uint2) + sizeof(DevSchedule);
    uint2* win = reinterpret_cast<uint2*>(smem_raw + size_t(warp) * per_warp);
    DevSchedule* sch = reinterpret_cast<DevSchedule*>(smem_raw + size_t(warp) * per_warp + size_t(window_steps) * sizeof(uint2));

    const bool is_fic = g.fic_enabled && slot_k == 0;
    const int k = slot_k - (g.fic_enabled ? 1 : 0);
    int sched_index = 0;   // schedule 0 is the FIC schedule
    const int8_t* soft;
    uint8_t* out;
    SubDesc sd{};
    if (is_fic) {
        soft = bits + size_t(s) * stream_stride + size_t(c) * size_t(g.nb_fib_cif_bits);
        out = o.fib_bytes + (size_t(s) * g.nb_cifs + c) * size_t(g.fib_group_bytes);
    } else {
        if (k >= g.max_subs) return;
        sd = subs[size_t(s) * g.max_subs + k];
        if (sd.schedule < 0) return;
        int32_t* nbytes = o.msc_nbytes + (size_t(s) * g.nb_cifs + c) * size_t(g.max_subs) + k;
        if (sd.overflow) { if (lane == 0) *nbytes = -1; return; }
        // CIF_Deinterleaver::Deinterleave refuses until 16 CIFs were consumed (cif_deinterleaver.cpp:40-43)
        if (stored[size_t(s) * g.max_subs + k] + c + 1 < ENS_DEINT_DEPTH) { if (lane == 0) *nbytes = 0; return; }
        sched_index = sd.schedule;
        soft = deint + (size_t(s) * g.nb_cifs + c) * size_t(g.cif_pitch) + size_t(sd.start_cu) * 64u;
        out = o.msc_bytes + (size_t(s) * g.nb_cifs + c) * size_t(g.msc_cif_bytes) + size_t(sd.start_cu) * 8u;
    }
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&schedules[sched_index]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sch);
        for (int i = lane; i < int(sizeof(DevSchedule) / 4); i += 32) dst[i] = src[i];
        __syncwarp();
    }
    uint2* spill = nullptr;
    if (sch->total_steps > window_steps) {
        const int rank = long_rank[k];   // only sub-channel slots can be longer than the window
        spill = scratch + ((size_t(rank) * g.nb_cifs + c) * size_t(g.n_streams) + s) * size_t(scratch_steps);
    }
    EnsView view{soft, out, prbs};
    const uint64_t err = viterbi_trellis(sch, view, win, spill, window_steps, lane);
    const uint32_t n_out_bytes = sch->n_out_bits / 8u;
    if (is_fic) {
        __syncwarp();
        const int fib_bytes = g.fib_group_bytes / g.nb_fibs_per_cif;
        if (lane < g.nb_fibs_per_cif) {
            const uint8_t* fib = out + lane * fib_bytes;
            const uint32_t rx = (uint32_t(fib[fib_bytes - 2]) << 8) | fib[fib_bytes - 1];
            o.fib_valid[(size_t(s) * g.nb_cifs + c) * size_t(g.nb_fibs_per_cif) + lane] = uint8_t(rx == crc16_fib(fib, fib_bytes - 2));
        }
        if (lane == 0) o.fic_error[size_t(s) * g.nb_cifs + c] = err;
    } else if (lane == 0) {
        o.msc_nbytes[(size_t(s) * g.nb_cifs + c) * size_t(g.max_subs) + k] = int32_t(n_out_bytes);
        o.msc_error[(size_t(s) * g.nb_cifs + c) * size_t(g.max_subs) + k] = err;
    }
}

// The same stage with one trellis per THREAD (viterbi_lanes.cuh), used when a call carries thousands of trellises.  A group
// is 32 consecutive (cif, stream) pairs of one slot (0 = FIC when enabled, then the sub-channel slots), stream fastest -- streams
// tuned to the same ensemble walk the same schedule in lock step; group id = slot * groups_per_slot + chunk.  Warp w runs group
// groups[w], longest slots first (vitl_plan: the short FIC groups are launched last and fill the gaps), with its decision rows
// starting at row warp_row[w] of the scratch.
__global__ void __maxnreg__(VITL_MAX_REGS)
ens_viterbi_lanes_kernel(EnsGeom g, const int8_t* __restrict__ bits, size_t stream_stride, const int32_t* __restrict__ frames_in_call, int slot,
                         const int8_t* __restrict__ deint, const SubDesc* __restrict__ subs, const int32_t* __restrict__ stored,
                         const DevSchedule* __restrict__ schedules, const uint8_t* __restrict__ prbs, EnsOut o, int groups_per_slot,
                         const int32_t* __restrict__ groups, const unsigned long long* __restrict__ warp_row, uint2* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t rows = uint32_t(warp_row[w + 1] - warp_row[w]);
    const long long per_slot = (long long)(g.nb_cifs) * g.n_streams;
    const int gid = groups[w];
    const int slot_k = gid / groups_per_slot;
    const long long j = (long long)(gid - slot_k * groups_per_slot) * 32 + lane;
    const int s = int(j % g.n_streams);
    const int c = int(j / g.n_streams);
    const bool is_fic = g.fic_enabled && slot_k == 0;
    const int k = slot_k - (g.fic_enabled ? 1 : 0);

    bool active = j < per_slot && stream_has_frame(frames_in_call, slot, s);
    const DevSchedule* sch = schedules;   // schedule 0 is the FIC schedule
    const int8_t* soft = bits;
    uint8_t* out = o.fib_bytes;
    if (active) {
        if (is_fic) {
            soft = bits + size_t(s) * stream_stride + size_t(c) * size_t(g.nb_fib_cif_bits);
            out = o.fib_bytes + (size_t(s) * g.nb_cifs + c) * size_t(g.fib_group_bytes);
        } else {
            active = false;
            if (k < g.max_subs) {
                const SubDesc sd = subs[size_t(s) * g.max_subs + k];
                if (sd.schedule >= 0) {
                    int32_t* nbytes = o.msc_nbytes + (size_t(s) * g.nb_cifs + c) * size_t(g.max_subs) + k;
                    if (sd.overflow) *nbytes = -1;
                    // CIF_Deinterleaver::Deinterleave refuses until 16 CIFs were consumed (cif_deinterleaver.cpp:40-43)
                    else if (stored[size_t(s) * g.max_subs + k] + c + 1 < ENS_DEINT_DEPTH) *nbytes = 0;
                    else if (schedules[sd.schedule].total_steps <= rows) {
                        active = true;
                        sch = schedules + sd.schedule;
                        soft = deint + (size_t(s) * g.nb_cifs + c) * size_t(g.cif_pitch) + size_t(sd.start_cu) * 64u;
                        out = o.msc_bytes + (size_t(s) * g.nb_cifs + c) * size_t(g.msc_cif_bytes) + size_t(sd.start_cu) * 8u;
                    }
                }
            }
        }
    }
    EnsView view{soft, out, prbs};
    const uint64_t err = viterbi_lane_trellis(sch, view, scratch + size_t(warp_row[w]) * 32u, lane, active);
    if (!active) return;
    if (is_fic) {
        const int fib_bytes = g.fib_group_bytes / g.nb_fibs_per_cif;
        for (int f = 0; f < g.nb_fibs_per_cif; f++) {
            const uint8_t* fib = out + f * fib_bytes;
            const uint32_t rx = (uint32_t(fib[fib_bytes - 2]) << 8) | fib[fib_bytes - 1];
            o.fib_valid[(size_t(s) * g.nb_cifs + c) * size_t(g.nb_fibs_per_cif) + f] = uint8_t(rx == crc16_fib(fib, fib_bytes - 2));
        }
        o.fic_error[size_t(s) * g.nb_cifs + c] = err;
    } else {
        o.msc_nbytes[(size_t(s) * g.nb_cifs + c) * size_t(g.max_subs) + k] = int32_t(sch->n_out_bits / 8u);
        o.msc_error[(size_t(s) * g.nb_cifs + c) * size_t(g.max_subs) + k] = err;
    }
}

// ---- 4. commit: CIF_Deinterleaver::Consume's counters (cif_deinterleaver.cpp:30-34) for every live sub-channel
__global__ void ens_commit_kernel(EnsGeom g, const int32_t* __restrict__ frames_in_call, int slot, unsigned long long* __restrict__ cif_count,
                                  const SubDesc* __restrict__ subs, int32_t* __restrict__ stored, int32_t* __restrict__ decoded) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.n_streams * g.max_subs) return;
    const int s = idx / g.max_subs, k = idx % g.max_subs;
    const bool has = stream_has_frame(frames_in_call, slot, s);
    if (k == 0) {
        decoded[s] = has ? 1 : 0;
        if (has) cif_count[s] += uint64_t(g.nb_cifs);
    }
    if (!has) return;
    const SubDesc sd = subs[idx];
    if (sd.schedule >= 0 && !sd.overflow) stored[idx] = min(ENS_DEINT_DEPTH, stored[idx] + g.nb_cifs);
}

// ------------------------------------------------------------------------------------------------ host side

// UEP_PROTECTION_TABLE (subchannel_protection_tables.h:21-86): {Lx[4], PIx[4]} per row; the reference's row order is kept
static const uint8_t UEP_ROWS[64][8] = {
    {3, 4, 17, 0, 5, 3, 2, 0},      {3, 3, 18, 0, 11, 6, 5, 0},     {3, 4, 14, 3, 15, 9, 6, 8},     {3, 4, 14, 3, 22, 13, 8, 13},
    {3, 5, 13, 3, 24, 17, 12, 17},  {4, 3, 26, 3, 5, 4, 2, 3},      {3, 4, 26, 3, 9, 6, 4, 6},      {3, 4, 26, 3, 15, 10, 6, 9},
    {3, 4, 26, 3, 24, 14, 8, 15},   {3, 5, 25, 3, 24, 18, 13, 18},  {6, 10, 23, 3, 5, 4, 2, 3},     {6, 10, 23, 3, 9, 6, 4, 5},
    {6, 12, 21, 3, 16, 7, 6, 9},    {6, 10, 23, 3, 23, 13, 8, 13},  {6, 9, 31, 2, 5, 3, 2, 3},      {6, 9, 33, 0, 11, 6, 5, 0},
    {6, 12, 27, 3, 16, 8, 6, 9},    {6, 10, 29, 3, 23, 13, 8, 13},  {6, 11, 28, 3, 24, 18, 12, 18}, {6, 10, 41, 3, 6, 3, 2, 3},
    {6, 10, 41, 3, 11, 6, 5, 6},    {6, 11, 40, 3, 16, 8, 6, 7},    {6, 10, 41, 3, 23, 13, 8, 13},  {6, 10, 41, 3, 24, 17, 12, 18},
    {7, 9, 53, 3, 5, 4, 2, 4},      {7, 10, 52, 3, 9, 6, 4, 6},     {6, 12, 51, 3, 16, 9, 6, 10},   {6, 10, 53, 3, 22, 12, 9, 12},
    {6, 13, 50, 3, 24, 18, 13, 19}, {14, 17, 50, 3, 5, 4, 2, 5},    {11, 21, 49, 3, 9, 6, 4, 8},    {11, 23, 47, 3, 16, 8, 6, 9},
    {11, 21, 49, 3, 23, 12, 9, 14}, {12, 19, 62, 3, 5, 3, 2, 4},    {11, 21, 61, 3, 11, 6, 5, 7},   {11, 22, 60, 3, 16, 9, 6, 10},
    {11, 21, 61, 3, 22, 12, 9, 14}, {11, 20, 62, 3, 24, 17, 13, 19}, {11, 19, 87, 3, 5, 4, 2, 4},   {11, 23, 83, 3, 11, 6, 5, 9},
    {11, 24, 82, 3, 16, 8, 6, 11},  {11, 21, 85, 3, 22, 11, 9, 13}, {11, 22, 84, 3, 24, 18, 12, 19}, {11, 20, 110, 3, 6, 4, 2, 5},
    {11, 22, 108, 3, 10, 6, 4, 9},  {11, 24, 106, 3, 16, 10, 6, 11}, {11, 20, 110, 3, 22, 13, 9, 13}, {11, 21, 109, 3, 24, 20, 13, 24},
    {12, 22, 131, 3, 8, 6, 2, 6},   {12, 26, 127, 3, 12, 8, 4, 11}, {11, 20, 134, 3, 16, 10, 7, 9}, {11, 22, 132, 3, 24, 16, 10, 15},
    {11, 24, 130, 3, 24, 20, 12, 20}, {11, 24, 154, 3, 6, 5, 2, 5}, {11, 24, 154, 3, 12, 9, 5, 10}, {11, 27, 151, 3, 16, 10, 7, 10},
    {11, 22, 156, 3, 24, 14, 10, 13}, {11, 26, 152, 3, 24, 19, 14, 18}, {11, 26, 200, 3, 8, 5, 2, 6}, {11, 25, 201, 3, 13, 9, 5, 10},
    {11, 26, 200, 3, 24, 17, 9, 17}, {11, 27, 247, 3, 8, 6, 2, 7},  {11, 24, 250, 3, 16, 9, 7, 10}, {12, 28, 245, 3, 24, 20, 14, 23},
};
// EEP_PROTECTION_TABLE_TYPE_A / _B (:121-140) as {capacity unit multiple, m0, b0, m1, b1, PI0, PI1}; Lx = m * n + b
static const int EEP_ROWS_A[4][7] = {{12, 6, -3, 0, 3, 24, 23}, {8, 2, -3, 4, 3, 14, 13}, {6, 6, -3, 0, 3, 8, 7}, {4, 4, -3, 2, 3, 3, 2}};
static const int EEP_ROW_2A_SPECIAL[7] = {8, 0, 5, 0, 1, 13, 12};   // :129-130, used when the sub-channel has 8 CU (:145-149)
static const int EEP_ROWS_B[4][7] = {{27, 24, -3, 0, 3, 10, 9}, {21, 24, -3, 0, 3, 6, 5}, {18, 24, -3, 0, 3, 4, 3}, {15, 24, -3, 0, 3, 2, 1}};

static void add_segment(dab_vit_schedule* sch, uint32_t* soft_left, int pi, uint32_t n_out) {
    dab_vit_segment sg{};
    sg.code_len = uint32_t(dab_get_puncture_code(pi, sg.counts));
    sg.n_out = n_out;
    // depuncture_symbols returns an empty result when the input runs out inside the segment (dab_viterbi_decoder.cpp:158-162,
    // the assert is compiled out in release builds): the whole update() then decodes nothing and consumes nothing.
    uint32_t need = 0;
    for (uint32_t g = 0; g < n_out / 4u; g++) need += sg.counts[g % sg.code_len];
    if (need > *soft_left || n_out == 0) return;
    *soft_left -= need;
    sch->seg[sch->n_seg++] = sg;
}

static int subchannel_schedule(const dab_subchannel* sub, dab_vit_schedule* sch, uint32_t* n_soft_out) {
    if (!sub || !sch) return set_error(DAB_ERR_INVALID, "null argument");
    if (sub->length <= 0 || sub->start_address < 0) return set_error(DAB_ERR_INVALID, "sub-channel %d: bad address %d / length %d", sub->id, sub->start_address, sub->length);
    memset(sch, 0, sizeof(*sch));
    const uint32_t n_soft = uint32_t(sub->length) * 64u;
    uint32_t left = n_soft;
    if (!sub->is_uep) {
        if (sub->eep_prot_level < 0 || sub->eep_prot_level > 3) return set_error(DAB_ERR_INVALID, "sub-channel %d: EEP level %d outside 0..3", sub->id, sub->eep_prot_level);
        const int* row = sub->eep_type_b ? EEP_ROWS_B[sub->eep_prot_level] : (sub->length == 8 ? EEP_ROW_2A_SPECIAL : EEP_ROWS_A[sub->eep_prot_level]);
        const int n = sub->length / row[0];
        for (int i = 0; i < 2; i++) {
            const int lx = row[1 + 2 * i] * n + row[2 + 2 * i];
            if (lx > 0) add_segment(sch, &left, row[5 + i], uint32_t(128 * lx));
        }
    } else {
        if (sub->uep_prot_index < 0 || sub->uep_prot_index > 63) return set_error(DAB_ERR_INVALID, "sub-channel %d: UEP index %d outside 0..63", sub->id, sub->uep_prot_index);
        const uint8_t* row = UEP_ROWS[sub->uep_prot_index];
        for (int i = 0; i < 4; i++) add_segment(sch, &left, row[4 + i], 128u * row[i]);
    }
    add_segment(sch, &left, 0, 24);
    uint32_t steps = 0;
    for (uint32_t i = 0; i < sch->n_seg; i++) steps += sch->seg[i].n_out / 4u;
    sch->n_out_bytes = steps >= 6u ? (steps - 6u) / 8u : 0u;   // msc_decoder.cpp:103-108
    if (n_soft_out) *n_soft_out = n_soft;
    return DAB_OK;
}

struct Ensemble {
    int device = 0;
    dab_parameters params{};
    EnsGeom g{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // dab_ensemble_set_decode_stream: everything after the ingest of a frame runs here, so that it overlaps whatever the caller queues
    // on `stream` next (the next frame's upload and demodulation)
    cudaStream_t decode_stream = nullptr;
    cudaEvent_t ev_pushed = nullptr, ev_done = nullptr;
    bool done_pending = false;
    DeviceBuffer<int8_t> d_fic_copy;        // [n_streams][nb_fic_bits]: the FIC soft bits of the ingested frame
    DeviceBuffer<int32_t> d_frames_copy;    // [n_streams]: the caller's frames_in_call at ingest
    int max_smem_optin = 0;
    std::mutex mtx;
    // host tables
    std::vector<std::vector<dab_subchannel>> subs;   // per stream
    std::vector<SubDesc> h_subs;                     // [n_streams][max_subs]
    std::vector<DevSchedule> schedules;              // [0] = FIC
    std::vector<dab_vit_schedule> schedule_keys;     // what each schedule was digested from (de-duplication)
    bool tables_dirty = true;
    int jobs_per_cif = 0;
    uint32_t window_steps = 32, scratch_steps = 0;
    int n_long = 0;
    uint64_t work_trellises = 0, work_steps = 0;
    // device
    DeviceBuffer<int8_t> d_ring, d_deint, d_in;
    DeviceBuffer<SubDesc> d_subs;
    DeviceBuffer<int32_t> d_stored, d_decoded, d_present, d_long_rank, d_msc_nbytes;
    DeviceBuffer<unsigned long long> d_cif_count, d_fic_error, d_msc_error;
    DeviceBuffer<DevSchedule> d_schedules;
    DeviceBuffer<uint8_t> d_prbs, d_fib_bytes, d_fib_valid, d_msc_bytes;
    DeviceBuffer<uint2> d_scratch;
    // one-trellis-per-thread form (ens_viterbi_lanes_kernel): decision rows per slot
    int groups_per_slot = 0;
    VitlPlan lane_plan;
    DeviceBuffer<int32_t> d_groups;
    DeviceBuffer<unsigned long long> d_warp_row;
    DeviceBuffer<uint2> d_lane_scratch;
    uint64_t launches = 0;
};

static int find_or_add_schedule(Ensemble* e, const dab_vit_schedule& key) {
    for (size_t i = 1; i < e->schedule_keys.size(); i++)
        if (memcmp(&e->schedule_keys[i], &key, sizeof(key)) == 0) return int(i);
    DevSchedule d;
    int rc = digest_schedule(&key, &d);
    if (rc != DAB_OK) return rc;
    e->schedule_keys.push_back(key);
    e->schedules.push_back(d);
    return int(e->schedules.size()) - 1;
}

// rebuild the device tables after a sub-channel change; d_stored must already hold the carried-over counters
static int upload_tables(Ensemble* e) {
    const EnsGeom& g = e->g;
    const size_t cap = (size_t(e->max_smem_optin) - 1024) / VIT_WARPS_PER_CTA - sizeof(DevSchedule);
    const uint32_t cap_steps = uint32_t(cap / sizeof(uint2)) & ~31u;
    // per slot k: the longest trellis any stream runs there
    std::vector<uint32_t> slot_steps(size_t(g.max_subs), 0);
    int used_slots = 0;
    e->work_trellises = 0;
    e->work_steps = 0;
    for (int s = 0; s < g.n_streams; s++) {
        for (int k = 0; k < g.max_subs; k++) {
            const SubDesc& sd = e->h_subs[size_t(s) * g.max_subs + k];
            if (sd.schedule < 0) continue;
            used_slots = std::max(used_slots, k + 1);
            if (sd.overflow) continue;
            const uint32_t st = e->schedules[size_t(sd.schedule)].total_steps;
            slot_steps[size_t(k)] = std::max(slot_steps[size_t(k)], st);
            e->work_trellises += uint64_t(g.nb_cifs);
            e->work_steps += uint64_t(g.nb_cifs) * st;
        }
    }
    if (g.fic_enabled) {
        e->work_trellises += uint64_t(g.n_streams) * g.nb_cifs;
        e->work_steps += uint64_t(g.n_streams) * g.nb_cifs * e->schedules[0].total_steps;
    }
    e->jobs_per_cif = (g.fic_enabled ? 1 : 0) + used_slots;
    std::vector<int32_t> long_rank(size_t(g.max_subs), -1);
    uint32_t window = g.fic_enabled ? e->schedules[0].total_steps : 32u, longest = 0;
    e->n_long = 0;
    for (int k = 0; k < g.max_subs; k++) {
        if (slot_steps[size_t(k)] > cap_steps) {
            long_rank[size_t(k)] = e->n_long++;
            longest = std::max(longest, slot_steps[size_t(k)]);
        } else {
            window = std::max(window, slot_steps[size_t(k)]);
        }
    }
    e->window_steps = std::min(cap_steps, std::max((window + 31u) & ~31u, 32u));
    e->scratch_steps = (longest + 31u) & ~31u;
    if (e->n_long > 0) DAB_CUDA_CHECK(e->d_scratch.reserve(size_t(e->n_long) * g.nb_cifs * g.n_streams * e->scratch_steps));
    e->groups_per_slot = int((size_t(g.nb_cifs) * size_t(g.n_streams) + 31) / 32);
    {
        std::vector<uint32_t> cost(size_t(e->jobs_per_cif) * size_t(e->groups_per_slot), 0u);
        for (int sk = 0; sk < e->jobs_per_cif; sk++) {
            const bool fic = g.fic_enabled && sk == 0;
            const uint32_t st = fic ? e->schedules[0].total_steps : slot_steps[size_t(sk - (g.fic_enabled ? 1 : 0))];
            for (int q = 0; q < e->groups_per_slot; q++) cost[size_t(sk) * size_t(e->groups_per_slot) + size_t(q)] = st;
        }
        vitl_plan(cost, e->lane_plan);
    }
    const VitlPlan& lp = e->lane_plan;
    DAB_CUDA_CHECK(e->d_groups.reserve(std::max<size_t>(lp.groups.size(), 1)));
    DAB_CUDA_CHECK(e->d_warp_row.reserve(lp.warp_row.size()));
    DAB_CUDA_CHECK(cudaMemcpyAsync(e->d_groups.ptr, lp.groups.data(), lp.groups.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    DAB_CUDA_CHECK(cudaMemcpyAsync(e->d_warp_row.ptr, lp.warp_row.data(), lp.warp_row.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
    DAB_CUDA_CHECK(e->d_schedules.reserve(std::max<size_t>(e->schedules.size(), 64)));
    DAB_CUDA_CHECK(cudaMemcpyAsync(e->d_schedules.ptr, e->schedules.data(), e->schedules.size() * sizeof(DevSchedule), cudaMemcpyHostToDevice, e->stream));
    DAB_CUDA_CHECK(cudaMemcpyAsync(e->d_subs.ptr, e->h_subs.data(), e->h_subs.size() * sizeof(SubDesc), cudaMemcpyHostToDevice, e->stream));
    DAB_CUDA_CHECK(cudaMemcpyAsync(e->d_long_rank.ptr, long_rank.data(), long_rank.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    DAB_CUDA_CHECK(cudaStreamSynchronize(e->stream));   // the sources are host vectors that may change before the copy would run
    e->tables_dirty = false;
    return DAB_OK;
}

static int sync_streams(Ensemble* e) {
    DAB_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    if (e->decode_stream) DAB_CUDA_CHECK(cudaStreamSynchronize(e->decode_stream));
    return DAB_OK;
}

// One frame per stream: ingest (the only part that reads the caller's buffers) -> de-interleave -> Viterbi (+ descramble, CRC) ->
// commit.  With a decode stream the ingest stays on the handle's stream -- it also stages the FIC soft bits and the caller's
// frames_in_call, so the caller may overwrite both as soon as the work queued on that stream so far has run -- and the rest runs on
// the decode stream; the next ingest waits for the previous decode (it shares the ring counters and the staging buffers with it).
static int decode_device(Ensemble* e, const int8_t* d_bits, size_t stream_stride, const int32_t* d_frames_in_call, int slot) {
    const EnsGeom base = e->g;
    if (e->tables_dirty) {
        int rc = sync_streams(e);
        if (rc != DAB_OK) return rc;
        rc = upload_tables(e);
        if (rc != DAB_OK) return rc;
    }
    EnsGeom g = base;
    g.aligned16 = (reinterpret_cast<uintptr_t>(d_bits) % 16 == 0 && stream_stride % 16 == 0 && g.nb_fic_bits % 16 == 0 && g.nb_cif_bits % 16 == 0) ? 1 : 0;
    const bool has_msc = g.nb_cif_bits > 0 && e->jobs_per_cif > (g.fic_enabled ? 1 : 0);
    const bool split = e->decode_stream != nullptr;
    cudaStream_t first = e->stream, rest = split ? e->decode_stream : e->stream;
    const int8_t* fic_bits = d_bits;
    size_t fic_stride = stream_stride;
    const int32_t* frames = d_frames_in_call;
    if (split) {
        if (e->done_pending) DAB_CUDA_CHECK(cudaStreamWaitEvent(first, e->ev_done, 0));
        if (g.fic_enabled && g.nb_fic_bits > 0) {
            DAB_CUDA_CHECK(e->d_fic_copy.reserve(size_t(g.n_streams) * size_t(g.nb_fic_bits)));
            DAB_CUDA_CHECK(cudaMemcpy2DAsync(e->d_fic_copy.ptr, size_t(g.nb_fic_bits), d_bits, stream_stride, size_t(g.nb_fic_bits), size_t(g.n_streams),
                                             cudaMemcpyDeviceToDevice, first));
            fic_bits = e->d_fic_copy.ptr;
            fic_stride = size_t(g.nb_fic_bits);
        }
        if (d_frames_in_call) {
            DAB_CUDA_CHECK(e->d_frames_copy.reserve(size_t(g.n_streams)));
            DAB_CUDA_CHECK(cudaMemcpyAsync(e->d_frames_copy.ptr, d_frames_in_call, size_t(g.n_streams) * sizeof(int32_t), cudaMemcpyDeviceToDevice, first));
            frames = e->d_frames_copy.ptr;
        }
    }
    const dim3 grid(unsigned((g.plane_len + ENS_CHUNK - 1) / ENS_CHUNK), unsigned(g.nb_cifs), unsigned(g.n_streams));
    if (has_msc) {
        ens_push_kernel<<<grid, ENS_CHUNK, 0, first>>>(g, d_bits, stream_stride, d_frames_in_call, slot, e->d_cif_count.ptr, e->d_ring.ptr);
        e->launches++;
    }
    if (split) {
        DAB_CUDA_CHECK(cudaEventRecord(e->ev_pushed, first));
        DAB_CUDA_CHECK(cudaStreamWaitEvent(rest, e->ev_pushed, 0));
    }
    if (has_msc) {
        ens_deint_kernel<<<grid, ENS_CHUNK, 0, rest>>>(g, frames, slot, e->d_cif_count.ptr, e->d_ring.ptr, e->d_deint.ptr);
        e->launches++;
    }
    const long long total_jobs = (long long)(e->jobs_per_cif) * g.nb_cifs * g.n_streams;
    if (e->jobs_per_cif > 0 && vitl_use_lanes(total_jobs)) {
        DAB_CUDA_CHECK(e->d_lane_scratch.reserve(size_t(e->lane_plan.rows()) * 32u));
        EnsOut o{e->d_fib_bytes.ptr, e->d_fib_valid.ptr, e->d_fic_error.ptr, e->d_msc_bytes.ptr, e->d_msc_nbytes.ptr, e->d_msc_error.ptr};
        ens_viterbi_lanes_kernel<<<unsigned(e->lane_plan.n_warps()), VITL_THREADS, 0, rest>>>(
            g, fic_bits, fic_stride, frames, slot, e->d_deint.ptr, e->d_subs.ptr, e->d_stored.ptr, e->d_schedules.ptr, e->d_prbs.ptr, o,
            e->groups_per_slot, e->d_groups.ptr, e->d_warp_row.ptr, e->d_lane_scratch.ptr);
        e->launches++;
    } else if (e->jobs_per_cif > 0) {
        const size_t smem = size_t(VIT_WARPS_PER_CTA) * (size_t(e->window_steps) * sizeof(uint2) + sizeof(DevSchedule));
        DAB_CUDA_CHECK(cudaFuncSetAttribute(ens_viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
        const long long total = (long long)(e->jobs_per_cif) * g.nb_cifs * g.n_streams;
        const unsigned vgrid = unsigned((total + VIT_WARPS_PER_CTA - 1) / VIT_WARPS_PER_CTA);
        EnsOut o{e->d_fib_bytes.ptr, e->d_fib_valid.ptr, e->d_fic_error.ptr, e->d_msc_bytes.ptr, e->d_msc_nbytes.ptr, e->d_msc_error.ptr};
        ens_viterbi_kernel<<<vgrid, VIT_WARPS_PER_CTA * 32, smem, rest>>>(g, fic_bits, fic_stride, frames, slot, e->d_deint.ptr,
                                                                         e->d_subs.ptr, e->d_stored.ptr, e->d_schedules.ptr, e->d_prbs.ptr, o,
                                                                         e->jobs_per_cif, e->d_long_rank.ptr, e->d_scratch.ptr,
                                                                         e->scratch_steps, e->window_steps);
        e->launches++;
    }
    {
        const int n = g.n_streams * g.max_subs;
        ens_commit_kernel<<<(n + 255) / 256, 256, 0, rest>>>(g, frames, slot, e->d_cif_count.ptr, e->d_subs.ptr, e->d_stored.ptr,
                                                            e->d_decoded.ptr);
        e->launches++;
    }
    if (split) {
        DAB_CUDA_CHECK(cudaEventRecord(e->ev_done, rest));
        e->done_pending = true;
    }
    DAB_CUDA_CHECK(cudaGetLastError());
    return DAB_OK;
}

}  // namespace dabb200

using namespace dabb200;

extern "C" {

// get_dab_parameters (src/dab/constants/dab_parameters.h:26-93)
int dab_get_dab_parameters(int mode, dab_parameters* p) {
    if (!p) return set_error(DAB_ERR_INVALID, "null params");
    // {carriers, symbols incl. PRS, FIC symbols, MSC symbols, FIBs, CIFs, FIBs per CIF}
    static const int T[4][7] = {{1536, 76, 3, 72, 12, 4, 3}, {384, 76, 3, 72, 3, 1, 3}, {192, 153, 8, 144, 4, 1, 4}, {768, 76, 3, 72, 6, 2, 3}};
    if (mode < 1 || mode > 4) return set_error(DAB_ERR_INVALID, "Invalid transmission mode %d", mode);
    const int* t = T[mode - 1];
    p->nb_symbols = t[1] - 1;
    p->nb_frame_bits = t[0] * 2 * p->nb_symbols;
    p->nb_fic_symbols = t[2];
    p->nb_msc_symbols = t[3];
    p->nb_fibs = t[4];
    p->nb_cifs = t[5];
    p->nb_fibs_per_cif = t[6];
    p->nb_sym_bits = p->nb_frame_bits / p->nb_symbols;
    p->nb_fic_bits = p->nb_sym_bits * p->nb_fic_symbols;
    p->nb_msc_bits = p->nb_sym_bits * p->nb_msc_symbols;
    p->nb_fib_bits = p->nb_fic_bits / p->nb_fibs;
    p->nb_fib_cif_bits = p->nb_fib_bits * p->nb_fibs_per_cif;
    p->nb_cif_bits = p->nb_msc_bits / p->nb_cifs;
    return DAB_OK;
}

int dab_ensemble_subchannel_schedule(const dab_subchannel* sub, dab_vit_schedule* out, uint32_t* n_soft) {
    return subchannel_schedule(sub, out, n_soft);
}

dab_ensemble* dab_ensemble_create(const dab_parameters* params, const dab_ensemble_options* options, int* status) {
    auto fail = [&](int rc) -> dab_ensemble* { if (status) *status = rc; return nullptr; };
    if (!params || !options) return fail(set_error(DAB_ERR_INVALID, "null argument"));
    if (options->n_streams < 1) return fail(set_error(DAB_ERR_INVALID, "n_streams must be >= 1"));
    const dab_parameters& p = *params;
    if (p.nb_cifs < 1 || p.nb_fic_bits < 0 || p.nb_cif_bits < 0 || p.nb_cif_bits % 64 != 0 || p.nb_fibs_per_cif < 0 ||
        p.nb_fic_bits < p.nb_cifs * p.nb_fib_cif_bits || (p.nb_fib_cif_bits > 0 && p.nb_fibs_per_cif < 1))
        return fail(set_error(DAB_ERR_INVALID, "inconsistent DAB parameters"));
    int rc = select_device(options->device);
    if (rc != DAB_OK) return fail(rc);
    auto* e = new Ensemble();
    e->device = options->device;
    e->params = p;
    EnsGeom& g = e->g;
    g.n_streams = options->n_streams;
    g.nb_cifs = p.nb_cifs;
    g.max_subs = options->max_subchannels > 0 ? std::min(options->max_subchannels, DAB_ENSEMBLE_MAX_SUBCHANNELS) : DAB_ENSEMBLE_MAX_SUBCHANNELS;
    g.nb_fic_bits = p.nb_fic_bits;
    g.nb_fib_cif_bits = p.nb_fib_cif_bits;
    g.nb_cif_bits = p.nb_cif_bits;
    g.nb_fibs_per_cif = std::max(p.nb_fibs_per_cif, 1);
    // FIC_Decoder only knows the Mode I puncturing: 2304 soft bits -> 768 bits (fic_decoder.cpp:68-75); other sizes decode nothing
    g.fic_enabled = (p.nb_fib_cif_bits / 3 == (128 * 21 + 128 * 3 + 24) / 4 - 6 && p.nb_fib_cif_bits % 24 == 0) ? 1 : 0;
    g.plane_len = p.nb_cif_bits / 16;
    g.plane_pitch = (g.plane_len + 15) & ~15;
    g.ring_rows = ENS_DEINT_DEPTH + p.nb_cifs;
    g.cif_pitch = (p.nb_cif_bits + 15) & ~15;
    g.fib_group_bytes = p.nb_fib_cif_bits / 24;
    g.msc_cif_bytes = p.nb_cif_bits / 8;
    auto cuda_fail = [&](const char* what) -> dab_ensemble* {
        int r = set_error(DAB_ERR_CUDA, "ensemble create: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
        delete e;
        return fail(r);
    };
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) return cuda_fail("stream");
    e->stream = e->own_stream;
    if (cudaDeviceGetAttribute(&e->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device) != cudaSuccess) return cuda_fail("attribute");
    const size_t S = size_t(g.n_streams), C = size_t(g.nb_cifs), K = size_t(g.max_subs);
    bool ok = true;
    auto zero = [&](auto& buf, size_t n) {
        using T = std::remove_pointer_t<decltype(buf.ptr)>;
        if (!ok) return;
        ok = buf.reserve(std::max<size_t>(n, 1)) == cudaSuccess && cudaMemsetAsync(buf.ptr, 0, std::max<size_t>(n, 1) * sizeof(T), e->stream) == cudaSuccess;
    };
    zero(e->d_ring, S * g.ring_rows * 16 * size_t(g.plane_pitch));
    zero(e->d_deint, S * C * size_t(g.cif_pitch));
    zero(e->d_subs, S * K);
    zero(e->d_stored, S * K);
    zero(e->d_decoded, S);
    zero(e->d_present, S);
    zero(e->d_long_rank, K);
    zero(e->d_cif_count, S);
    zero(e->d_fib_bytes, S * C * size_t(g.fib_group_bytes));
    zero(e->d_fib_valid, S * C * size_t(g.nb_fibs_per_cif));
    zero(e->d_fic_error, S * C);
    zero(e->d_msc_bytes, S * C * size_t(g.msc_cif_bytes));
    zero(e->d_msc_nbytes, S * C * K);
    zero(e->d_msc_error, S * C * K);
    if (!ok) return cuda_fail("allocation");
    // AdditiveScrambler, syncword 0xFFFF (additive_scrambler.h:16-35; fic_decoder.cpp:47-48, msc_decoder.cpp:40-41)
    const size_t n_prbs = std::max<size_t>(std::max(g.fib_group_bytes, g.msc_cif_bytes), 16);
    std::vector<uint8_t> prbs(n_prbs);
    {
        uint16_t reg = 0xFFFF;
        for (size_t i = 0; i < n_prbs; i++) {
            uint8_t b = 0;
            for (int j = 0; j < 8; j++) {
                const uint8_t v = uint8_t(((reg >> 8) ^ (reg >> 4)) & 1u);
                b |= uint8_t(v << (7 - j));
                reg = uint16_t((reg << 1) | v);
            }
            prbs[i] = b;
        }
    }
    if (e->d_prbs.reserve(n_prbs) != cudaSuccess || cudaMemcpy(e->d_prbs.ptr, prbs.data(), n_prbs, cudaMemcpyHostToDevice) != cudaSuccess) return cuda_fail("prbs");
    // schedule 0: the FIC schedule PI_16 x 21 blocks, PI_15 x 3 blocks, PI_X (fic_decoder.cpp:77-88)
    dab_vit_schedule fic{};
    uint32_t left = 2304;
    add_segment(&fic, &left, 16, 128 * 21);
    add_segment(&fic, &left, 15, 128 * 3);
    add_segment(&fic, &left, 0, 24);
    fic.n_out_bytes = 96;
    DevSchedule d;
    rc = digest_schedule(&fic, &d);
    if (rc != DAB_OK) { delete e; return fail(rc); }
    e->schedule_keys.push_back(fic);
    e->schedules.push_back(d);
    e->subs.resize(S);
    SubDesc unused{0, 0, -1, 0};
    e->h_subs.assign(S * K, unused);
    e->tables_dirty = true;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) return cuda_fail("sync");
    if (status) *status = DAB_OK;
    return reinterpret_cast<dab_ensemble*>(e);
}

void dab_ensemble_destroy(dab_ensemble* h) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    if (e->decode_stream) cudaStreamSynchronize(e->decode_stream);
    if (e->ev_pushed) cudaEventDestroy(e->ev_pushed);
    if (e->ev_done) cudaEventDestroy(e->ev_done);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

int dab_ensemble_set_cuda_stream(dab_ensemble* h, void* cuda_stream) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
    return DAB_OK;
}

int dab_ensemble_set_decode_stream(dab_ensemble* h, void* cuda_stream) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    std::lock_guard<std::mutex> lock(e->mtx);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(e->device));
    { int rc_sync = sync_streams(e); if (rc_sync != DAB_OK) return rc_sync; }
    if (cuda_stream && !e->ev_pushed) {
        DAB_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_pushed, cudaEventDisableTiming));
        DAB_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming));
    }
    e->decode_stream = static_cast<cudaStream_t>(cuda_stream);
    e->done_pending = false;
    return DAB_OK;
}

int dab_ensemble_set_subchannels(dab_ensemble* h, int stream, const dab_subchannel* subs, int n_subs) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    const EnsGeom& g = e->g;
    if (stream < -1 || stream >= g.n_streams) return set_error(DAB_ERR_INVALID, "stream %d out of range", stream);
    if (n_subs < 0 || n_subs > g.max_subs || (n_subs > 0 && !subs)) return set_error(DAB_ERR_CAPACITY, "%d sub-channels, the handle holds %d per stream", n_subs, g.max_subs);
    std::lock_guard<std::mutex> lock(e->mtx);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(e->device));
    // validate + digest first so that a bad descriptor changes nothing
    std::vector<int> sched(static_cast<size_t>(n_subs));
    for (int k = 0; k < n_subs; k++) {
        dab_vit_schedule key;
        int rc = subchannel_schedule(&subs[k], &key, nullptr);
        if (rc != DAB_OK) return rc;
        rc = find_or_add_schedule(e, key);
        if (rc < 0) return rc;
        sched[size_t(k)] = rc;
    }
    // the de-interleaver counters live on the device: fetch, carry over for unchanged descriptors, write back
    const size_t K = size_t(g.max_subs);
    std::vector<int32_t> stored(size_t(g.n_streams) * K);
    { int rc_sync = sync_streams(e); if (rc_sync != DAB_OK) return rc_sync; }
    DAB_CUDA_CHECK(cudaMemcpy(stored.data(), e->d_stored.ptr, stored.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    auto same = [](const dab_subchannel& a, const dab_subchannel& b) {
        return a.start_address == b.start_address && a.length == b.length && (a.is_uep != 0) == (b.is_uep != 0) &&
               (a.is_uep ? a.uep_prot_index == b.uep_prot_index : (a.eep_prot_level == b.eep_prot_level && (a.eep_type_b != 0) == (b.eep_type_b != 0)));
    };
    const int s0 = stream < 0 ? 0 : stream, s1 = stream < 0 ? g.n_streams : stream + 1;
    for (int s = s0; s < s1; s++) {
        const std::vector<dab_subchannel>& old = e->subs[size_t(s)];
        std::vector<int32_t> carried(static_cast<size_t>(n_subs), 0);
        for (int k = 0; k < n_subs; k++)
            for (size_t j = 0; j < old.size(); j++)
                if (same(old[j], subs[k])) { carried[size_t(k)] = stored[size_t(s) * K + j]; break; }
        for (size_t k = 0; k < K; k++) {
            SubDesc& sd = e->h_subs[size_t(s) * K + k];
            if (int(k) < n_subs) {
                sd.start_cu = subs[k].start_address;
                sd.length_cu = subs[k].length;
                sd.schedule = sched[k];
                sd.overflow = (size_t(subs[k].start_address) + size_t(subs[k].length)) * 64u > size_t(g.nb_cif_bits) ? 1 : 0;
                stored[size_t(s) * K + k] = carried[k];
            } else {
                sd = SubDesc{0, 0, -1, 0};
                stored[size_t(s) * K + k] = 0;
            }
        }
        e->subs[size_t(s)].assign(subs, subs + n_subs);
    }
    // garbage-collect the puncturing schedules no sub-channel refers to any more (schedule 0 is the FIC's): a receiver that is
    // re-tuned for days must not grow the host and device tables without bound
    {
        std::vector<int> remap(e->schedules.size(), -1);
        remap[0] = 0;
        int next = 1;
        for (const SubDesc& sd : e->h_subs)
            if (sd.schedule > 0 && remap[size_t(sd.schedule)] < 0) remap[size_t(sd.schedule)] = next++;
        if (size_t(next) < e->schedules.size()) {
            std::vector<dab_vit_schedule> keys(static_cast<size_t>(next));
            std::vector<DevSchedule> scheds(static_cast<size_t>(next));
            for (size_t i = 0; i < remap.size(); i++)
                if (remap[i] >= 0) { keys[size_t(remap[i])] = e->schedule_keys[i]; scheds[size_t(remap[i])] = e->schedules[i]; }
            e->schedule_keys.swap(keys);
            e->schedules.swap(scheds);
            for (SubDesc& sd : e->h_subs)
                if (sd.schedule > 0) sd.schedule = remap[size_t(sd.schedule)];
        }
    }
    DAB_CUDA_CHECK(cudaMemcpy(e->d_stored.ptr, stored.data(), stored.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    DAB_CUDA_CHECK(cudaMemsetAsync(e->d_msc_nbytes.ptr + size_t(s0) * g.nb_cifs * K, 0, size_t(s1 - s0) * g.nb_cifs * K * sizeof(int32_t), e->stream));
    e->tables_dirty = true;
    return DAB_OK;
}

int dab_ensemble_decode_frames_device(dab_ensemble* h, const int8_t* d_bits, size_t stream_stride, const int32_t* d_frames_in_call, int slot) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    if (!d_bits || slot < 0) return set_error(DAB_ERR_INVALID, "null buffer");
    if (stream_stride < size_t(e->params.nb_fic_bits) + size_t(e->params.nb_cifs) * size_t(e->params.nb_cif_bits))
        return set_error(DAB_ERR_INVALID, "stream stride %zu is shorter than one frame", stream_stride);
    std::lock_guard<std::mutex> lock(e->mtx);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(e->device));
    return decode_device(e, d_bits, stream_stride, d_frames_in_call, slot);
}

int dab_ensemble_decode_frames(dab_ensemble* h, const int8_t* bits, const uint8_t* present) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    if (!bits) return set_error(DAB_ERR_INVALID, "null buffer");
    std::lock_guard<std::mutex> lock(e->mtx);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(e->device));
    const EnsGeom& g = e->g;
    const size_t frame = size_t(g.nb_fic_bits) + size_t(g.nb_cifs) * size_t(g.nb_cif_bits);
    const size_t pitch = (frame + 15) & ~size_t(15);
    DAB_CUDA_CHECK(e->d_in.reserve(pitch * size_t(g.n_streams)));
    DAB_CUDA_CHECK(cudaMemcpy2DAsync(e->d_in.ptr, pitch, bits, frame, frame, size_t(g.n_streams), cudaMemcpyHostToDevice, e->stream));
    const int32_t* mask = nullptr;
    std::vector<int32_t> present32;
    if (present) {
        present32.resize(size_t(g.n_streams));
        for (int s = 0; s < g.n_streams; s++) present32[size_t(s)] = present[s] ? 1 : 0;
        DAB_CUDA_CHECK(cudaMemcpyAsync(e->d_present.ptr, present32.data(), present32.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        mask = e->d_present.ptr;
    }
    int rc = decode_device(e, e->d_in.ptr, pitch, mask, 0);
    // `bits` / present32 are caller / stack memory: the staged copies must have left them before returning
    { int rc_sync = sync_streams(e); if (rc_sync != DAB_OK) return rc_sync; }
    return rc;
}

int dab_ensemble_device_results(dab_ensemble* h, dab_ensemble_results* out) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e || !out) return set_error(DAB_ERR_INVALID, "null argument");
    out->fib_bytes = e->d_fib_bytes.ptr;
    out->fib_valid = e->d_fib_valid.ptr;
    out->fic_error = reinterpret_cast<const uint64_t*>(e->d_fic_error.ptr);
    out->msc_bytes = e->d_msc_bytes.ptr;
    out->msc_nbytes = e->d_msc_nbytes.ptr;
    out->msc_error = reinterpret_cast<const uint64_t*>(e->d_msc_error.ptr);
    out->decoded = e->d_decoded.ptr;
    out->fib_group_bytes = size_t(e->g.fib_group_bytes);
    out->msc_cif_bytes = size_t(e->g.msc_cif_bytes);
    out->nb_cifs = e->g.nb_cifs;
    out->nb_fibs_per_cif = e->g.nb_fibs_per_cif;
    out->max_subchannels = e->g.max_subs;
    return DAB_OK;
}

int dab_ensemble_read_fic(dab_ensemble* h, int stream, uint8_t* fib_bytes, uint8_t* fib_valid, uint64_t* path_error) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    const EnsGeom& g = e->g;
    if (stream < 0 || stream >= g.n_streams) return set_error(DAB_ERR_INVALID, "stream %d out of range", stream);
    std::lock_guard<std::mutex> lock(e->mtx);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(e->device));
    { int rc_sync = sync_streams(e); if (rc_sync != DAB_OK) return rc_sync; }
    const size_t C = size_t(g.nb_cifs);
    if (fib_bytes && g.fib_group_bytes > 0)
        DAB_CUDA_CHECK(cudaMemcpy(fib_bytes, e->d_fib_bytes.ptr + size_t(stream) * C * g.fib_group_bytes, C * g.fib_group_bytes, cudaMemcpyDeviceToHost));
    if (fib_valid) DAB_CUDA_CHECK(cudaMemcpy(fib_valid, e->d_fib_valid.ptr + size_t(stream) * C * g.nb_fibs_per_cif, C * g.nb_fibs_per_cif, cudaMemcpyDeviceToHost));
    if (path_error) DAB_CUDA_CHECK(cudaMemcpy(path_error, e->d_fic_error.ptr + size_t(stream) * C, C * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return DAB_OK;
}

int dab_ensemble_read_msc(dab_ensemble* h, int stream, int cif, int sub_index, uint8_t* out, size_t capacity, int32_t* n_bytes, uint64_t* path_error) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    const EnsGeom& g = e->g;
    if (stream < 0 || stream >= g.n_streams || cif < 0 || cif >= g.nb_cifs) return set_error(DAB_ERR_INVALID, "stream %d / cif %d out of range", stream, cif);
    std::lock_guard<std::mutex> lock(e->mtx);
    if (sub_index < 0 || size_t(sub_index) >= e->subs[size_t(stream)].size()) return set_error(DAB_ERR_INVALID, "stream %d has no sub-channel slot %d", stream, sub_index);
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(e->device));
    { int rc_sync = sync_streams(e); if (rc_sync != DAB_OK) return rc_sync; }
    int32_t decoded = 0, n = 0;
    DAB_CUDA_CHECK(cudaMemcpy(&decoded, e->d_decoded.ptr + stream, sizeof(int32_t), cudaMemcpyDeviceToHost));
    const size_t idx = (size_t(stream) * g.nb_cifs + cif) * size_t(g.max_subs) + size_t(sub_index);
    DAB_CUDA_CHECK(cudaMemcpy(&n, e->d_msc_nbytes.ptr + idx, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (!decoded && n > 0) n = 0;   // the stream had no frame in the last call: nothing new
    if (n_bytes) *n_bytes = n;
    if (n > 0) {
        if (!out || capacity < size_t(n)) return set_error(DAB_ERR_CAPACITY, "sub-channel decoded %d bytes, buffer holds %zu", n, capacity);
        const size_t off = (size_t(stream) * g.nb_cifs + cif) * size_t(g.msc_cif_bytes) + size_t(e->subs[size_t(stream)][size_t(sub_index)].start_address) * 8u;
        DAB_CUDA_CHECK(cudaMemcpy(out, e->d_msc_bytes.ptr + off, size_t(n), cudaMemcpyDeviceToHost));
        if (path_error) DAB_CUDA_CHECK(cudaMemcpy(path_error, e->d_msc_error.ptr + idx, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    } else if (path_error) {
        *path_error = 0;
    }
    return DAB_OK;
}

int dab_ensemble_sync(dab_ensemble* h) {
    auto* e = reinterpret_cast<Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    DeviceGuard device_guard;
    DAB_CUDA_CHECK(cudaSetDevice(e->device));
    { int rc_sync = sync_streams(e); if (rc_sync != DAB_OK) return rc_sync; }
    return DAB_OK;
}

int dab_ensemble_schedule_count(const dab_ensemble* h) {
    auto* e = reinterpret_cast<const Ensemble*>(h);
    return e ? int(e->schedules.size()) : 0;
}

uint64_t dab_ensemble_kernel_launches(const dab_ensemble* h) {
    auto* e = reinterpret_cast<const Ensemble*>(h);
    return e ? e->launches : 0;
}

int dab_ensemble_last_work(const dab_ensemble* h, uint64_t* trellises, uint64_t* trellis_steps) {
    auto* e = reinterpret_cast<const Ensemble*>(h);
    if (!e) return set_error(DAB_ERR_INVALID, "null handle");
    if (trellises) *trellises = e->work_trellises;
    if (trellis_steps) *trellis_steps = e->work_steps;
    return DAB_OK;
}

}  // extern "C"
