/* dab_b200.h -- C ABI of the B200-native DAB receive hot path (libdab_b200.so).
 *
 * This is the drop-in boundary for the receive hot path and the callers either side of its Viterbi stage, nothing else:
 *   1. OFDM_Demod            /root/reference/src/ofdm/ofdm_demodulator.h:47-168   (IQ -> int8 soft bits)
 *   2. DAB_Viterbi_Decoder   /root/reference/src/dab/algorithms/dab_viterbi_decoder.h:12-45 (punctured soft bits -> bytes)
 *   3. FIC_Decoder / MSC_Decoder + CIF_Deinterleaver
 *                            /root/reference/src/dab/fic/fic_decoder.h:17-40, src/dab/msc/msc_decoder.h:18-43,
 *                            src/dab/msc/cif_deinterleaver.h:10-24   (frame soft bits -> FIBs + sub-channel bytes)
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.  The C++ mirror classes that keep the
 * reference's own signatures on top of these calls live in dab-radio_b200/cpp/ (see INTEGRATION.md).
 *
 * Every entry point needs a CUDA device (sm_100a).  There is no CPU fallback: without a usable device the create
 * calls fail with DAB_ERR_NO_DEVICE and every other call fails with DAB_ERR_INVALID on the NULL handle.
 *
 * All functions returning int return a dab_status (0 = ok, negative = error); dab_last_error() gives the text of the
 * most recent error on the calling thread.
 */
#ifndef DAB_B200_H
#define DAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DAB_API
#else
#define DAB_API __attribute__((visibility("default")))
#endif

typedef enum {
    DAB_OK = 0,
    DAB_ERR_INVALID = -1,    /* bad argument / NULL handle */
    DAB_ERR_NO_DEVICE = -2,  /* no CUDA device, or the device is not sm_100 */
    DAB_ERR_CUDA = -3,       /* a CUDA runtime call or kernel failed */
    DAB_ERR_UNDERRUN = -4,   /* Viterbi: fewer punctured symbols than the schedule consumes (dab_viterbi_decoder.cpp:158-162) */
    DAB_ERR_CAPACITY = -5,   /* block larger than the handle was created for, too many schedules, ... */
    DAB_ERR_TRACEBACK = -6   /* Viterbi: chainback longer than the decoded bits (viterbi_decoder_core.h:216-218) */
} dab_status;

DAB_API const char* dab_last_error(void);
DAB_API const char* dab_version(void);
/* number of usable sm_100 devices (0 => nothing in this library can run) */
DAB_API int dab_device_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * DAB tables (replaces get_DAB_OFDM_params / get_DAB_PRS_reference / get_DAB_mapper_ref,
 * src/ofdm/dab_ofdm_params_ref.cpp:10, dab_prs_ref.cpp:140, dab_mapper_ref.cpp:10; GetPunctureCode puncture_codes.h:69)
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { float re, im; } dab_c32; /* layout of std::complex<float> / float2 */

/* same fields and order as OFDM_Params (src/ofdm/ofdm_params.h:5-12) */
typedef struct {
    size_t nb_frame_symbols;
    size_t nb_symbol_period;
    size_t nb_null_period;
    size_t nb_cyclic_prefix;
    size_t nb_fft;
    size_t nb_data_carriers;
} dab_ofdm_params;

DAB_API int dab_get_ofdm_params(int transmission_mode, dab_ofdm_params* out);
DAB_API int dab_get_prs_reference(int transmission_mode, dab_c32* out, size_t nb_fft);
DAB_API int dab_get_mapper_reference(int* out, size_t nb_data_carriers, size_t nb_fft);
/* PI_1..PI_24 as 8 counts per 32 mother bits; pi = 0 returns the 6-entry tail code PI_X.  Returns the code length. */
DAB_API int dab_get_puncture_code(int pi, uint8_t out[8]);

/* ------------------------------------------------------------------------------------------------------------------
 * OFDM demodulator (replaces OFDM_Demod, src/ofdm/ofdm_demodulator.h:109-141)
 * One handle owns n_streams independent demodulators of one transmission mode on one GPU; every stream keeps its own
 * state machine (ofdm_demodulator.cpp:235-275) in device memory and streams are batched per kernel launch.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct dab_ofdm dab_ofdm;

/* OFDM_Demod_Config (ofdm_demodulator.h:24-45), flattened */
typedef struct {
    float signal_l1_update_beta;
    int signal_l1_nb_samples;
    int signal_l1_nb_decimate;
    float null_l1_thresh_null_start;
    float null_l1_thresh_null_end;
    float sync_fine_freq_update_beta;
    int sync_is_coarse_freq_correction;
    float sync_max_coarse_freq_correction_norm;
    float sync_coarse_freq_slow_beta;
    float sync_impulse_peak_threshold_db;
    float sync_impulse_peak_distance_probability;
} dab_ofdm_config;

/* OFDM_Demod::State (ofdm_demodulator.h:50-56) */
enum {
    DAB_OFDM_FINDING_NULL_POWER_DIP = 0,
    DAB_OFDM_READING_NULL_AND_PRS = 1,
    DAB_OFDM_RUNNING_COARSE_FREQ_SYNC = 2,
    DAB_OFDM_RUNNING_FINE_TIME_SYNC = 3,
    DAB_OFDM_READING_SYMBOLS = 4
};

/* the scalar getters of ofdm_demodulator.h:124-132 in one read */
typedef struct {
    int32_t state;
    int32_t fine_time_offset;
    int32_t total_frames_read;
    int32_t total_frames_desync;
    float signal_average;
    float fine_frequency_offset;
    float coarse_frequency_offset;
    int32_t reserved;
} dab_ofdm_state;

typedef struct {
    int64_t frame_start;      /* absolute sample index (per stream, since create/attach) of the PRS cyclic-prefix start */
    int32_t fine_time_offset;
    int32_t total_desync;
    float coarse_offset;      /* used by this frame's PLL */
    float fine_offset_used;   /* used by this frame's PLL */
    float fine_offset_after;  /* after this frame's cyclic-prefix update */
    float signal_average;     /* GetSignalAverage() after the Process() call that completed the frame */
} dab_ofdm_frame_info;

/* How a stream's samples are stored (examples/app_helpers/app_iq_readers.h:17-35,76-88,109-110,130-135: raw_u8, raw_s8,
 * raw_s16l/b, raw_u16l/b).  Raw integer pairs are uploaded as they come from the front end and dequantised on the device as
 * QuantisedIQ<T>::to_c32 does: (float(raw) - BIAS) * (1 / MAX_AMPLITUDE). */
typedef enum {
    DAB_IQ_F32 = 0,    /* std::complex<float> */
    DAB_IQ_U8 = 1,     /* (u8 - 127.5) / 127.5 */
    DAB_IQ_S8 = 2,     /* s8 / 127 */
    DAB_IQ_S16LE = 3,  /* s16 / 32767 */
    DAB_IQ_S16BE = 4,
    DAB_IQ_U16LE = 5,  /* (u16 - 32767.5) / 32767.5 */
    DAB_IQ_U16BE = 6
} dab_iq_format;
/* bytes per complex sample of a format (8, 2 or 4); 0 for an unknown format */
DAB_API size_t dab_iq_format_bytes(int format);

typedef struct {
    int n_streams;             /* >= 1 */
    int device;                /* CUDA ordinal */
    size_t max_block_samples;  /* largest n passed to a process call (0 => 262144) */
    int keep_debug_taps;       /* 1: also store frame FFT / DQPSK vectors for the GUI getters (doubles HBM writes) */
    int sample_format;         /* dab_iq_format of the stream buffers; DAB_IQ_U8 (= 1, the former raw_u8_ingest flag) and the other
                                  raw formats are fed through dab_ofdm_process_batch_raw and dequantised on the device */
} dab_ofdm_options;

/* Replaces On_OFDM_Frame().Attach(...) (ofdm_demodulator.h:140).  `bits` is owned by the library and valid only during the
 * call, like the reference's span over m_pipeline_out_bits (ofdm_demodulator.cpp:110,635).  Called on the thread that
 * invoked the process/sync entry point, in frame order per stream.
 * Re-entrancy: the callback runs inside the process call, which holds the handle's (recursive) lock.  It may call any
 * dab_ofdm_* function on the same handle from the same thread (GetState(), Reset() from an On_OFDM_Frame observer, as the
 * reference allows); other threads calling process / advance on the handle wait until the call returns.  The scalar getters
 * (dab_ofdm_get_state) and dab_ofdm_get_impulse_response / _coarse_frequency_response / _frame_data_bits are served from a host
 * snapshot taken at the end of the last synchronous process call and never wait for a call in flight on another thread. */
typedef void (*dab_ofdm_frame_cb)(void* user, int stream, const int8_t* bits, size_t n_bits, const dab_ofdm_frame_info* info);

/* A ready-made dab_ofdm_frame_cb for throughput measurements from languages whose own callbacks are slow (bench.py's e2e leg
 * delivers 1024 frames per step): `user` points at a dab_ofdm_frame_counter that accumulates the frames and soft bits
 * delivered and a checksum over the first and last 64 soft bits of every frame. */
typedef struct {
    uint64_t frames;
    uint64_t bits;
    uint64_t checksum;
} dab_ofdm_frame_counter;
DAB_API void dab_ofdm_count_frames_cb(void* user, int stream, const int8_t* bits, size_t n_bits, const dab_ofdm_frame_info* info);

/* OFDM_Demod::OFDM_Demod(params, prs_fft_ref, carrier_mapper, nb_desired_threads) -- the thread count has no meaning here */
DAB_API dab_ofdm* dab_ofdm_create(const dab_ofdm_params* params, const dab_c32* prs_fft_ref, const int* carrier_mapper,
                                  const dab_ofdm_options* options, int* status);
DAB_API void dab_ofdm_destroy(dab_ofdm* h);
/* run the kernels on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the handle's own stream */
DAB_API int dab_ofdm_set_cuda_stream(dab_ofdm* h, void* cuda_stream);
DAB_API int dab_ofdm_set_frame_callback(dab_ofdm* h, dab_ofdm_frame_cb cb, void* user);
/* GetConfig() is mutable between Process calls (examples/basic_radio_app.cpp:268-269); stream = -1 applies to all */
DAB_API int dab_ofdm_set_config(dab_ofdm* h, int stream, const dab_ofdm_config* cfg);
DAB_API int dab_ofdm_get_config(dab_ofdm* h, int stream, dab_ofdm_config* cfg);
DAB_API void dab_ofdm_default_config(dab_ofdm_config* cfg);

/* OFDM_Demod::Process(span<const complex<float>>) for one stream: host buffer, copied before return; frames completed by
 * this block are delivered through the callback before the call returns. */
DAB_API int dab_ofdm_process(dab_ofdm* h, int stream, const dab_c32* iq, size_t n);
/* The same for every stream at once: iq[s] / n[s] for s < n_streams (n[s] may be 0, iq[s] may then be NULL). */
DAB_API int dab_ofdm_process_batch(dab_ofdm* h, const dab_c32* const* iq, const size_t* n);
/* raw 8-bit IQ (examples/app_helpers/app_iq_readers.h:17-69): sample = (u8 - 127.5) / 127.5, dequantised on the device */
DAB_API int dab_ofdm_process_batch_u8(dab_ofdm* h, const uint8_t* const* iq_u8, const size_t* n);
/* any raw format: iq[s] points at n[s] samples in the handle's sample_format (options->sample_format) */
DAB_API int dab_ofdm_process_batch_raw(dab_ofdm* h, const void* const* iq, const size_t* n);

/* Device-resident streams (the batched B200 path): d_iq points at n_streams rows of `stride_samples` samples already in
 * HBM.  The demodulator then reads the rows in place -- no ring copy.  dab_ofdm_advance(h, n) is one Process() call of n[s]
 * further samples per stream.  Soft bits stay on the device (dab_ofdm_device_bits) unless a callback is attached. */
DAB_API int dab_ofdm_attach_device_streams(dab_ofdm* h, const dab_c32* d_iq, size_t stride_samples, size_t total_samples);
/* the same for rows of raw samples in the handle's sample_format (what a capture card DMAs into HBM): 2 or 4 bytes per sample
 * instead of 8 out of HBM. */
DAB_API int dab_ofdm_attach_device_streams_raw(dab_ofdm* h, const void* d_iq, size_t stride_samples, size_t total_samples);
/* dab_ofdm_advance* only queue work: internally the streams are split into pipeline ways on CUDA streams of their own, so that
 * consecutive calls overlap (no GPU-wide barrier per call).  dab_ofdm_join orders the handle's CUDA stream after everything
 * queued so far without blocking the host; dab_ofdm_device_bits, dab_ofdm_sync, the getters and the frame callback join
 * implicitly.  Work the caller queues on the handle's stream BEFORE a call is always ordered before that call. */
DAB_API int dab_ofdm_join(dab_ofdm* h);
/* The attached rows used as ring buffers: the caller has refilled (or laid out periodically) the region behind the demodulator
 * and moves every stream's origin forward by delta_samples -- sample index i of a row now means what index i + delta meant, so the
 * next dab_ofdm_advance continues at index (samples advanced so far - delta).  At least one frame + two NULL symbols + one symbol
 * must remain behind the cursor.  dab_ofdm_frame_info::frame_start counts from the new origin afterwards. */
DAB_API int dab_ofdm_rebase_device_streams(dab_ofdm* h, size_t delta_samples);
DAB_API int dab_ofdm_advance(dab_ofdm* h, const size_t* n);
DAB_API int dab_ofdm_advance_uniform(dab_ofdm* h, size_t n);
/* device pointer / geometry of the soft-bit output of the most recent process/advance call:
 * bits[stream][slot][n_bits], frames_in_call[stream] = how many slots are valid */
DAB_API int dab_ofdm_device_bits(dab_ofdm* h, const int8_t** d_bits, size_t* n_bits, int* slots_per_stream, const int32_t** d_frames_in_call);

DAB_API int dab_ofdm_reset(dab_ofdm* h, int stream);           /* OFDM_Demod::Reset() */
DAB_API int dab_ofdm_get_state(dab_ofdm* h, int stream, dab_ofdm_state* out);
DAB_API int dab_ofdm_sync(dab_ofdm* h);                        /* wait for queued GPU work, deliver pending callbacks */
DAB_API size_t dab_ofdm_frame_bits(const dab_ofdm* h);         /* (nb_frame_symbols-1)*nb_data_carriers*2 */
DAB_API int dab_ofdm_get_params(const dab_ofdm* h, dab_ofdm_params* out); /* GetOFDMParams() */
/* GUI getters (ofdm_demodulator.h:133-139), latest values for one stream copied to host memory */
DAB_API int dab_ofdm_get_impulse_response(dab_ofdm* h, int stream, float* out, size_t nb_fft);
DAB_API int dab_ofdm_get_coarse_frequency_response(dab_ofdm* h, int stream, float* out, size_t nb_fft);
DAB_API int dab_ofdm_get_frame_data_bits(dab_ofdm* h, int stream, int8_t* out, size_t n_bits);
/* GetCorrelationTimeBuffer(): the NULL + PRS window (nb_null_period + nb_symbol_period samples) the next / last sync works on */
DAB_API int dab_ofdm_get_correlation_time_buffer(dab_ofdm* h, int stream, dab_c32* out, size_t n);
DAB_API int dab_ofdm_get_frame_fft(dab_ofdm* h, int stream, dab_c32* out, size_t n);      /* needs keep_debug_taps */
DAB_API int dab_ofdm_get_frame_data_vec(dab_ofdm* h, int stream, dab_c32* out, size_t n); /* needs keep_debug_taps */
/* number of library kernels launched so far by this handle (bench.py's gpu_launches) */
DAB_API uint64_t dab_ofdm_kernel_launches(const dab_ofdm* h);
/* Per-kernel device timing for the roofline measurement: when enabled, every control / frame kernel launch is bracketed by
 * CUDA events on the launching stream.  Times are accumulated per pass index (pass 0 = first frame completion of a call). */
#define DAB_OFDM_TIMING_PASSES 8
typedef struct {
    double frame_ms[DAB_OFDM_TIMING_PASSES];
    uint64_t frame_launches[DAB_OFDM_TIMING_PASSES];
    double control_ms[DAB_OFDM_TIMING_PASSES];
    uint64_t control_launches[DAB_OFDM_TIMING_PASSES];
} dab_ofdm_kernel_times;
DAB_API int dab_ofdm_set_kernel_timing(dab_ofdm* h, int enable);                    /* also clears the accumulators */
DAB_API int dab_ofdm_get_kernel_times(dab_ofdm* h, dab_ofdm_kernel_times* out);     /* synchronises the stream */

/* Stage-level entry used by the parity tests and the roofline measurement: demodulate already aligned frames (complex float).
 * d_frames: n_frames rows of frame_stride samples, row = PRS + data symbols (nb_frame_symbols * nb_symbol_period samples);
 * freq_offset[n_frames] (host); d_bits: n_frames * frame_bits; d_phase_error: n_frames * nb_frame_symbols floats (per symbol). */
DAB_API int dab_ofdm_demod_frames_device(dab_ofdm* h, const dab_c32* d_frames, size_t frame_stride, int n_frames,
                                         const float* freq_offset, int8_t* d_bits, float* d_phase_error);

/* ------------------------------------------------------------------------------------------------------------------
 * Viterbi decoder (replaces DAB_Viterbi_Decoder, src/dab/algorithms/dab_viterbi_decoder.h:24-33, on top of
 * vendor/viterbi_decoder ViterbiDecoder_AVX_u16<7,4>): K = 7, rate 1/4 mother code G = {109, 79, 83, 109}, int8 soft input,
 * saturating u16 path metrics, tie -> predecessor 1, renormalisation at 60455.  Output is bit-exact with that decoder.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct dab_viterbi dab_viterbi;

#define DAB_VIT_MAX_SEGMENTS 8
#define DAB_VIT_MAX_SCHEDULES 1024

/* one update(punctured, puncture_code, requested_output_symbols) call (dab_viterbi_decoder.h:26-30) */
typedef struct {
    uint8_t counts[8];   /* puncture code as kept-symbol counts per group of 4 mother symbols (puncture_codes.h:42-67) */
    uint32_t code_len;   /* 1..8, the code is applied cyclically */
    uint32_t n_out;      /* requested_output_symbols, multiple of 4 */
} dab_vit_segment;

/* reset(start_state); update(...) x n_seg; chainback(n_out_bytes, end_state) */
typedef struct {
    dab_vit_segment seg[DAB_VIT_MAX_SEGMENTS];
    uint32_t n_seg;
    uint32_t n_out_bytes;
    uint32_t start_state;
    uint32_t end_state;
} dab_vit_schedule;

typedef struct {
    uint32_t schedule;     /* id from dab_viterbi_add_schedule */
    uint32_t n_soft;       /* punctured symbols available at soft_offset */
    uint64_t soft_offset;  /* into the soft buffer */
    uint64_t out_offset;   /* into the output byte buffer */
} dab_vit_job;

DAB_API dab_viterbi* dab_viterbi_create(int device, int* status);
DAB_API void dab_viterbi_destroy(dab_viterbi* h);
DAB_API int dab_viterbi_set_cuda_stream(dab_viterbi* h, void* cuda_stream);
DAB_API int dab_viterbi_add_schedule(dab_viterbi* h, const dab_vit_schedule* s); /* >= 0: id */
/* punctured symbols consumed by a schedule (what the update() calls would return in total) */
DAB_API int64_t dab_viterbi_schedule_soft_symbols(const dab_vit_schedule* s);
/* host buffers in, host buffers out; path_error[n_jobs] may be NULL; job_status[n_jobs] (dab_status per job) may be NULL */
DAB_API int dab_viterbi_decode_batch(dab_viterbi* h, const int8_t* soft, size_t soft_bytes, const dab_vit_job* jobs, int n_jobs,
                                     uint8_t* out, size_t out_bytes, uint64_t* path_error, int32_t* job_status);
/* device buffers in and out, jobs on the host; asynchronous on the handle's stream */
DAB_API int dab_viterbi_decode_batch_device(dab_viterbi* h, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* jobs,
                                            int n_jobs, uint8_t* d_out, size_t out_bytes, uint64_t* d_path_error, int32_t* d_job_status);
/* jobs already on the device too (nothing crosses PCIe; used by the roofline measurement) */
DAB_API int dab_viterbi_decode_jobs_device(dab_viterbi* h, const int8_t* d_soft, size_t soft_bytes, const dab_vit_job* d_jobs,
                                           int n_jobs, uint32_t max_steps, uint8_t* d_out, size_t out_bytes,
                                           uint64_t* d_path_error, int32_t* d_job_status);
/* A job list that is decoded again and again (the sub-channel layout of an ensemble changes rarely): uploads the jobs once,
 * orders them so that the 32 trellises a warp runs in lock step share a puncturing schedule, and launches the longest trellises
 * first so that the short ones (FIC groups next to 1.5-kbit sub-channel trellises) fill the gaps at the end of the launch instead
 * of delaying a long one.  Returns a plan id >= 1.  dab_viterbi_decode_prepared is asynchronous on the handle's stream; soft bits, output,
 * path errors and job status are device pointers laid out as the jobs' offsets say. */
DAB_API int dab_viterbi_prepare_jobs(dab_viterbi* h, const dab_vit_job* jobs, int n_jobs);
DAB_API int dab_viterbi_decode_prepared(dab_viterbi* h, int plan, const int8_t* d_soft, size_t soft_bytes, uint8_t* d_out, size_t out_bytes,
                                        uint64_t* d_path_error, int32_t* d_job_status);
DAB_API int dab_viterbi_release_jobs(dab_viterbi* h, int plan);
/* One trellis with an ad-hoc schedule (what the streaming reset/update/chainback mirror class issues on chainback) */
DAB_API int dab_viterbi_decode_one(dab_viterbi* h, const dab_vit_schedule* s, const int8_t* soft, size_t n_soft, uint8_t* out,
                                   uint64_t* path_error);
DAB_API int dab_viterbi_sync(dab_viterbi* h);
DAB_API uint64_t dab_viterbi_kernel_launches(const dab_viterbi* h);
/* ------------------------------------------------------------------------------------------------------------------
 * Ensemble decoder: everything between one OFDM frame of soft bits and the decoded bytes, batched over streams.
 * Replaces, per stream, BasicRadio::Process's split of the frame (src/basic_radio/basic_radio.cpp:42-63), one FIC_Decoder
 * (src/dab/fic/fic_decoder.cpp:53-116: Viterbi PI_16 x21 + PI_15 x3 + PI_X, energy dispersal, CRC16 per FIB) and one
 * MSC_Decoder per sub-channel (src/dab/msc/msc_decoder.cpp:46-170: CIF_Deinterleaver over 16 CIFs
 * (cif_deinterleaver.cpp:21-70), EEP / UEP puncturing schedule (subchannel_protection_tables.h), Viterbi, energy dispersal).
 * The soft bits never leave HBM between the demodulator and the decoded bytes.
 * ------------------------------------------------------------------------------------------------------------------ */
/* same fields and order as DAB_Parameters (src/dab/constants/dab_parameters.h:5-21) */
typedef struct {
    int nb_frame_bits, nb_symbols, nb_fic_symbols, nb_msc_symbols, nb_fibs, nb_cifs, nb_fibs_per_cif;
    int nb_sym_bits, nb_fic_bits, nb_msc_bits, nb_fib_bits, nb_fib_cif_bits, nb_cif_bits;
} dab_parameters;
DAB_API int dab_get_dab_parameters(int transmission_mode, dab_parameters* out); /* get_dab_parameters, dab_parameters.h:26-93 */

/* the fields of Subchannel (src/dab/database/dab_database_entities.h:179-190) the decoder reads */
typedef struct {
    int32_t id;
    int32_t start_address;   /* capacity units (64 soft bits) from the start of the CIF */
    int32_t length;          /* capacity units */
    int32_t is_uep;
    int32_t uep_prot_index;  /* row of UEP_PROTECTION_TABLE, 0..63 */
    int32_t eep_prot_level;  /* 0..3 = level 1..4 */
    int32_t eep_type_b;      /* 0 = EEP_Type::TYPE_A, 1 = TYPE_B */
    int32_t reserved;
} dab_subchannel;

#define DAB_ENSEMBLE_MAX_SUBCHANNELS 64
typedef struct {
    int n_streams;
    int device;
    int max_subchannels;  /* per stream, 0 => 64 */
} dab_ensemble_options;

typedef struct dab_ensemble dab_ensemble;
DAB_API dab_ensemble* dab_ensemble_create(const dab_parameters* params, const dab_ensemble_options* options, int* status);
DAB_API void dab_ensemble_destroy(dab_ensemble* h);
DAB_API int dab_ensemble_set_cuda_stream(dab_ensemble* h, void* cuda_stream);
/* Optional second stream (NULL: off).  With it dab_ensemble_decode_frames_device only INGESTS on the handle's stream (soft bits ->
 * de-interleaver ring, FIC soft bits and frames_in_call staged): once the work queued on that stream so far has run the caller may
 * overwrite d_bits / d_frames_in_call, e.g. by demodulating the next frame.  De-interleave, Viterbi, descramble, CRC and commit run on
 * the decode stream, concurrently with whatever the caller queues on the handle's stream next; the results are valid once the
 * decode stream has run (queue the read-back on it, or dab_ensemble_sync).  The next ingest waits for the previous decode. */
DAB_API int dab_ensemble_set_decode_stream(dab_ensemble* h, void* cuda_stream);
/* The sub-channel set of one stream (stream = -1: every stream), i.e. which MSC_Decoder objects exist
 * (basic_radio.cpp:100-153).  A sub-channel whose descriptor is unchanged keeps its de-interleaver history, a new one starts
 * empty and yields no bytes until 16 CIFs have been consumed (cif_deinterleaver.cpp:40-43). */
DAB_API int dab_ensemble_set_subchannels(dab_ensemble* h, int stream, const dab_subchannel* subs, int n_subs);
/* MSC_Decoder's update() schedule for a sub-channel as a dab_vit_schedule (msc_decoder.cpp:88-99,136-146); n_soft receives
 * length * 64.  A segment that would underrun is dropped exactly as the reference's release build does
 * (dab_viterbi_decoder.cpp:158-162). */
DAB_API int dab_ensemble_subchannel_schedule(const dab_subchannel* sub, dab_vit_schedule* out, uint32_t* n_soft);

/* One OFDM frame (nb_frame_bits soft bits: FIC then MSC) per stream, already on the device: stream s's frame starts at
 * d_bits + s * stream_stride.  d_frames_in_call (optional, dab_ofdm_device_bits) selects the streams that have a frame:
 * stream s is decoded iff d_frames_in_call == NULL or d_frames_in_call[s] > slot.  Asynchronous on the handle's stream. */
DAB_API int dab_ensemble_decode_frames_device(dab_ensemble* h, const int8_t* d_bits, size_t stream_stride,
                                              const int32_t* d_frames_in_call, int slot);
/* The same from host memory: bits[n_streams][nb_frame_bits]; present[n_streams] (optional) = 0 skips a stream. */
DAB_API int dab_ensemble_decode_frames(dab_ensemble* h, const int8_t* bits, const uint8_t* present);

/* Device-resident results of the most recent decode call (valid until the next one):
 *   fib_bytes [n_streams][nb_cifs][nb_fib_cif_bits/24]   descrambled FIB groups (FIC_Decoder::m_decoded_bytes)
 *   fib_valid [n_streams][nb_cifs][nb_fibs_per_cif]      1 = CRC16 matches (the FIB would be passed to OnFIB)
 *   fic_error [n_streams][nb_cifs]                       Viterbi path error
 *   msc_bytes [n_streams][nb_cifs][nb_cif_bits/8]        sub-channel k's bytes start at start_address * 8
 *   msc_nbytes[n_streams][nb_cifs][max_subchannels]      decoded bytes (0 while the de-interleaver fills, -1 = overflows the CIF)
 *   msc_error [n_streams][nb_cifs][max_subchannels]
 *   decoded   [n_streams]                                1 = the stream had a frame in this call */
typedef struct {
    const uint8_t* fib_bytes;
    const uint8_t* fib_valid;
    const uint64_t* fic_error;
    const uint8_t* msc_bytes;
    const int32_t* msc_nbytes;
    const uint64_t* msc_error;
    const int32_t* decoded;
    size_t fib_group_bytes, msc_cif_bytes;
    int nb_cifs, nb_fibs_per_cif, max_subchannels;
} dab_ensemble_results;
DAB_API int dab_ensemble_device_results(dab_ensemble* h, dab_ensemble_results* out);
/* host copies (synchronise, then copy): FIC of one stream / one sub-channel of one CIF */
DAB_API int dab_ensemble_read_fic(dab_ensemble* h, int stream, uint8_t* fib_bytes, uint8_t* fib_valid, uint64_t* path_error);
DAB_API int dab_ensemble_read_msc(dab_ensemble* h, int stream, int cif, int sub_index, uint8_t* out, size_t capacity,
                                  int32_t* n_bytes, uint64_t* path_error);
DAB_API int dab_ensemble_sync(dab_ensemble* h);
DAB_API uint64_t dab_ensemble_kernel_launches(const dab_ensemble* h);
/* puncturing schedules currently held (the FIC's + one per distinct sub-channel profile / length in use: unused ones are
 * garbage-collected by dab_ensemble_set_subchannels) */
DAB_API int dab_ensemble_schedule_count(const dab_ensemble* h);
/* trellises decoded / trellis steps run by the most recent decode call (host-side count from the sub-channel tables; upper
 * bound when d_frames_in_call masks streams) */
DAB_API int dab_ensemble_last_work(const dab_ensemble* h, uint64_t* trellises, uint64_t* trellis_steps);

#ifdef __cplusplus
}
#endif
#endif
