/* TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement, in plain C, of the reference's receive hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build, load or call this.
 * The product (dab-radio_b200/, include/dab_b200.h) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function below against the reference's own
 * sources compiled unmodified (oracle/_ref/libdabref.so) and against the committed fixtures in tests/golden/ that were
 * generated from that library (tests/golden/make_golden.py).  One boundary is unpinned by any reference test: FFT
 * rounding (FFTW3 is a third-party dependency absent from the tree, vcpkg.json:19-21) -- see DESIGN.md.
 */
#ifndef DAB_ORACLE_H
#define DAB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } orc_c32;

/* src/ofdm/ofdm_params.h:5-12 */
typedef struct {
    int nb_frame_symbols, nb_symbol_period, nb_null_period, nb_cyclic_prefix, nb_fft, nb_data_carriers;
} orc_params;

/* src/ofdm/ofdm_demodulator.h:24-45 (defaults in orc_default_config) */
typedef struct {
    float signal_l1_update_beta;
    int signal_l1_nb_samples;
    int signal_l1_nb_decimate;
    float thresh_null_start;
    float thresh_null_end;
    float fine_freq_update_beta;
    int is_coarse_freq_correction;
    float max_coarse_freq_correction_norm;
    float coarse_freq_slow_beta;
    float impulse_peak_threshold_db;
    float impulse_peak_distance_probability;
} orc_config;

typedef struct {
    int64_t frame_start;      /* absolute index of the PRS cyclic-prefix start picked by fine time sync */
    int32_t fine_time_offset;
    int32_t total_desync;
    float coarse_offset;      /* used by this frame's PLL */
    float fine_offset_used;   /* used by this frame's PLL */
    float fine_offset_after;  /* after this frame's cyclic-prefix update */
    float signal_average;
} orc_frame_info;

typedef struct {
    int32_t state;
    int32_t fine_time_offset;
    int32_t total_frames_read;
    int32_t total_frames_desync;
    float signal_average;
    float fine_offset;
    float coarse_offset;
    int32_t pad;
} orc_ofdm_state;

typedef void (*orc_frame_cb)(void* user, const int8_t* bits, size_t n_bits, const orc_frame_info* info);

/* tables: dab_ofdm_params_ref.cpp:10-57, dab_prs_ref.cpp:140-194, dab_mapper_ref.cpp:10-50 */
int orc_get_params(int mode, orc_params* out);
int orc_get_prs(int mode, orc_c32* out /* nb_fft */);
int orc_get_mapper(int mode, int* out /* nb_data_carriers */);
void orc_default_config(orc_config* cfg);

/* dsp: apply_pll.cpp:82-116 (AVX lane arithmetic), complex_conj_mul_sum.cpp:65-100, fftw3 semantics */
void orc_apply_pll(const orc_c32* x, orc_c32* y, size_t n, float freq_norm, float dt_norm);
orc_c32 orc_conj_mul_sum(const orc_c32* x0, const orc_c32* x1, size_t n);
void orc_fft(const orc_c32* in, orc_c32* out, int n, int sign /* -1 forward, +1 backward */);

/* test transmitter: ofdm_modulator.cpp:49-93 */
int orc_modulate(int mode, const uint8_t* bytes, size_t nbytes, orc_c32* frame_out, size_t nsamples);

/* OFDM_Demod in real-time order (single thread): ofdm_demodulator.cpp:235-950 */
typedef struct orc_ofdm orc_ofdm;
orc_ofdm* orc_ofdm_create(int mode);
orc_ofdm* orc_ofdm_create_custom(const orc_params* p, const orc_c32* prs_fft_ref, const int* mapper);
void orc_ofdm_destroy(orc_ofdm* d);
orc_config* orc_ofdm_config(orc_ofdm* d);
void orc_ofdm_set_callback(orc_ofdm* d, orc_frame_cb cb, void* user);
void orc_ofdm_process(orc_ofdm* d, const orc_c32* buf, size_t n);
void orc_ofdm_reset(orc_ofdm* d);
void orc_ofdm_get_state(const orc_ofdm* d, orc_ofdm_state* s);
/* collected frames (when no callback is set the oracle keeps every frame) */
size_t orc_ofdm_frames_done(const orc_ofdm* d);
int orc_ofdm_get_frame(const orc_ofdm* d, size_t index, orc_frame_info* info, int8_t* bits_out);
size_t orc_ofdm_frame_bits(const orc_ofdm* d);
/* stage taps of the latest frame / sync */
const orc_c32* orc_ofdm_frame_fft(const orc_ofdm* d);       /* (S+1)*nb_fft */
const orc_c32* orc_ofdm_frame_data_vec(const orc_ofdm* d);  /* (S-1)*nb_data_carriers */
const float* orc_ofdm_impulse_response(const orc_ofdm* d);  /* nb_fft */
const float* orc_ofdm_coarse_freq_response(const orc_ofdm* d);
/* GetCorrelationTimeBuffer() (ofdm_demodulator.h:139): nb_null_period + nb_symbol_period samples, the first *length filled */
const orc_c32* orc_ofdm_correlation_time_buffer(const orc_ofdm* d, size_t* length);
/* stage-level entry: demodulate one already-aligned frame (S symbols of nb_symbol_period) with a given net offset */
void orc_ofdm_demod_frame(const orc_params* p, const int* mapper, const orc_c32* frame, float freq_offset, int8_t* bits_out,
                          float* phase_error_sum);

/* DAB_Viterbi_Decoder with the SIMD (AVX2/SSE4.1 u16) semantics: dab_viterbi_decoder.cpp:27-181,
 * viterbi_decoder_avx_u16.h:47-170, viterbi_decoder_core.h:180-236 */
typedef struct orc_viterbi orc_viterbi;
orc_viterbi* orc_vit_create(void);
void orc_vit_destroy(orc_viterbi* v);
void orc_vit_set_traceback_length(orc_viterbi* v, size_t n);
size_t orc_vit_get_traceback_length(const orc_viterbi* v);
size_t orc_vit_get_current_decoded_bit(const orc_viterbi* v);
void orc_vit_reset(orc_viterbi* v, size_t start_state);
size_t orc_vit_update(orc_viterbi* v, const int8_t* soft, size_t n_soft, const uint8_t* code, size_t code_len, size_t n_out);
uint64_t orc_vit_chainback(orc_viterbi* v, uint8_t* out, size_t nbytes, size_t end_state);
/* full job = reset + segments + chainback; seg_codes is [n_seg][8] */
uint64_t orc_vit_decode_job(orc_viterbi* v, const int8_t* soft, size_t n_soft, const uint8_t* seg_codes, const uint32_t* seg_code_len,
                            const uint32_t* seg_n_out, uint32_t n_seg, uint8_t* out, size_t n_out_bytes, size_t* consumed);
/* puncture tables: puncture_codes.h:42-74 */
const uint8_t* orc_puncture_code(int pi /* 1..24 */);
const uint8_t* orc_puncture_code_tail(void); /* PI_X, 6 entries */
/* mother code encoder (for tests): convolutional_encoder_shift_register.h:44-62 semantics, one soft symbol per output bit */
size_t orc_conv_encode(const uint8_t* bytes, size_t nbytes, int8_t* soft_out /* (8*nbytes+6)*4 */);
size_t orc_puncture(const int8_t* mother, size_t n_mother, const uint8_t* seg_codes, const uint32_t* seg_code_len, const uint32_t* seg_n_out,
                    uint32_t n_seg, int8_t* out);
/* FIC / MSC decode around the Viterbi decoder:
 * additive_scrambler.h:10-35, crc.h:24-33 + fic_decoder.cpp:20-33, fic_decoder.cpp:53-116, cif_deinterleaver.cpp:9-70,
 * subchannel_protection_tables.h:21-154, msc_decoder.cpp:27-170 */
void orc_scrambler_bytes(uint16_t syncword, uint8_t* out, size_t n);
uint16_t orc_crc16_fib(const uint8_t* x, size_t n);
uint64_t orc_fic_decode_group(orc_viterbi* v, const int8_t* bits, size_t nb_encoded_bits, size_t nb_fibs, uint8_t* out, uint8_t* valid);
typedef struct orc_deint orc_deint;
orc_deint* orc_deint_create(size_t nb_bits);
void orc_deint_destroy(orc_deint* d);
int orc_deint_push(orc_deint* d, const int8_t* bits, int8_t* out); /* Consume + Deinterleave; 1 when out was written */
/* Subchannel (src/dab/database/dab_database_entities.h:179-190), the fields the decoder reads */
typedef struct {
    int32_t start_address, length, is_uep, uep_prot_index, eep_prot_level, eep_type_b;
} orc_subchannel;
int orc_uep_subchannel_size(int index);
int orc_msc_segments(const orc_subchannel* sc, int pi_out[5], uint32_t n_out[5]);
typedef struct orc_msc orc_msc;
orc_msc* orc_msc_create(const orc_subchannel* sc);
void orc_msc_destroy(orc_msc* m);
int64_t orc_msc_decode_cif(orc_msc* m, const int8_t* cif_bits, size_t n_bits, uint8_t* out, uint64_t* path_error);

/* CPU baselines for bench.py (port kind): wall seconds */
double orc_ofdm_bench(int mode, int n_threads, const orc_c32* iq, size_t n, size_t block, int repeats, uint64_t* frames_out);
double orc_vit_bench(int n_threads, const int8_t* soft, size_t soft_per_job, size_t n_jobs, const uint8_t* seg_codes,
                     const uint32_t* seg_code_len, const uint32_t* seg_n_out, uint32_t n_seg, size_t traceback_bits, uint8_t* out,
                     size_t out_bytes_per_job);

#ifdef __cplusplus
}
#endif
#endif
