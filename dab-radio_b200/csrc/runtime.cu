// Error plumbing, device selection and the DAB constant tables behind the C ABI (include/dab_b200.h).
#include <cmath>
#include <vector>

#include "common.cuh"

namespace dabb200 {

std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}

int set_error(int status, const char* fmt, ...) {
    char buf[512];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buf, sizeof(buf), fmt, args);
    va_end(args);
    last_error_ref() = buf;
    return status;
}

int select_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        return set_error(DAB_ERR_NO_DEVICE, "no CUDA device available (%s); libdab_b200 has no CPU fallback",
                         e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return set_error(DAB_ERR_INVALID, "device %d out of range [0, %d)", device, count);
    cudaDeviceProp prop;
    DAB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        return set_error(DAB_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", device, prop.name,
                         prop.major, prop.minor);
    }
    DAB_CUDA_CHECK(cudaSetDevice(device));
    return DAB_OK;
}

}  // namespace dabb200

using namespace dabb200;

extern "C" {

const char* dab_last_error(void) { return last_error_ref().c_str(); }
const char* dab_version(void) { return "dab_b200 0.1 (sm_100a)"; }

int dab_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    int usable = 0;
    for (int i = 0; i < count; i++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) usable++;
    }
    return usable;
}

// DAB transmission modes I-IV at 2.048 MS/s (ETSI EN 300 401 clause 14.2; reference dab_ofdm_params_ref.cpp:10-57)
int dab_get_ofdm_params(int mode, dab_ofdm_params* p) {
    if (!p) return set_error(DAB_ERR_INVALID, "null params");
    // {symbols per frame, symbol period, null period, fft size, data carriers}
    static const size_t T[4][5] = {{76, 2552, 2656, 2048, 1536}, {76, 638, 664, 512, 384}, {153, 319, 345, 256, 192}, {76, 1276, 1328, 1024, 768}};
    if (mode < 1 || mode > 4) return set_error(DAB_ERR_INVALID, "Invalid transmission mode %d", mode);
    const size_t* t = T[mode - 1];
    p->nb_frame_symbols = t[0];
    p->nb_symbol_period = t[1];
    p->nb_null_period = t[2];
    p->nb_fft = t[3];
    p->nb_cyclic_prefix = t[1] - t[3];
    p->nb_data_carriers = t[4];
    return DAB_OK;
}

// Phase reference symbol, ETSI EN 300 401 clause 14.3.2 tables 23 (i, n per block of 32 carriers) and 24 (h values);
// the mode II-IV block parameters are those of the reference's dab_prs_ref.cpp:77-123.  Each string holds one digit per
// 32-carrier block, carriers ordered -K/2..-1, +1..+K/2; value at carrier k is exp(j*pi/2*(h[i][(k-k_block) % 16] + n)).
int dab_get_prs_reference(int mode, dab_c32* out, size_t nb_fft) {
    static const char* const block_i[4] = {"012301230123012301230123032103210321032103210321", "012301210321", "012321", "012301230123032103210321"};
    static const char* const block_n[4] = {"120132232123123322211312311122102233021333303011", "232212022103", "230222", "011222033132010201222130"};
    static const char* const h_rows[4] = {"0200001120002211", "0323013021232330", "0002021322022013", "0121033223212132"};
    dab_ofdm_params p;
    int rc = dab_get_ofdm_params(mode, &p);
    if (rc != DAB_OK) return rc;
    if (!out || nb_fft < p.nb_data_carriers + 1) {
        return set_error(DAB_ERR_INVALID, "FFT buffer not large enough to fit phase reference symbol %zu<%zu", nb_fft, p.nb_data_carriers + 1);
    }
    for (size_t i = 0; i < nb_fft; i++) out[i] = dab_c32{0.0f, 0.0f};
    const int half = int(p.nb_data_carriers / 2);
    for (int c = 0; c < 2 * half; c++) {
        const int k = (c < half) ? (c - half) : (c - half + 1);
        const int blk = c / 32;
        const int h = h_rows[block_i[mode - 1][blk] - '0'][c % 16] - '0';
        const int n = block_n[mode - 1][blk] - '0';
        const float phi = 3.14159265358979323846f / 2.0f * float(h + n);
        out[(k < 0) ? (int(nb_fft) + k) : k] = dab_c32{std::cos(phi), std::sin(phi)};
    }
    return DAB_OK;
}

// Frequency interleaver, ETSI EN 300 401 clause 14.6.1: PI(i) = 13*PI(i-1) + N/4 - 1 mod N, kept when it lands on a data carrier.
int dab_get_mapper_reference(int* out, size_t nb_carriers, size_t nb_fft) {
    if (!out || nb_fft == 0 || nb_carriers >= nb_fft) return set_error(DAB_ERR_INVALID, "bad mapper geometry");
    const size_t dc = nb_fft / 2, lo = dc - nb_carriers / 2, hi = dc + nb_carriers / 2;
    size_t v = 0, n = 0;
    for (size_t i = 0; i < nb_fft && n < nb_carriers; i++) {
        if (i > 0) v = (13 * v + nb_fft / 4 - 1) % nb_fft;
        if (v < lo || v > hi || v == dc) continue;
        out[n++] = int(v - lo) - (v > dc ? 1 : 0);
    }
    return DAB_OK;
}

// ETSI EN 300 401 clause 11.1.2 table 13 expressed as kept-symbol counts per group of 4 mother bits:
// PI_p keeps (p-1)/8+1 bits per group, one more in ((p-1)%8)+1 groups chosen in bit-reversed order.
int dab_get_puncture_code(int pi, uint8_t out[8]) {
    if (!out || pi < 0 || pi > 24) return set_error(DAB_ERR_INVALID, "puncture index %d outside 0..24", pi);
    if (pi == 0) {
        for (int g = 0; g < 8; g++) out[g] = (g < 6) ? 2 : 0;
        return 6;
    }
    const int base = (pi - 1) / 8 + 1, extra = (pi - 1) % 8 + 1;
    for (int g = 0; g < 8; g++) {
        const int rev = ((g & 1) << 2) | (g & 2) | ((g >> 2) & 1);
        out[g] = uint8_t(base + (rev < extra ? 1 : 0));
    }
    return 8;
}

}  // extern "C"
