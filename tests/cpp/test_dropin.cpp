// Drop-in test: the caller code below is written against the REFERENCE interfaces only (OFDM_Demod as OFDM_Block uses it,
// examples/app_helpers/app_ofdm_blocks.h:25-58; DAB_Viterbi_Decoder as FIC_Decoder::DecodeFIBGroup uses it,
// src/dab/fic/fic_decoder.cpp:74-87) and links against the mirror classes in dab-radio_b200/cpp + libdab_b200.so.
//   test_dropin ofdm <mode> <block> <iq.c64> <out.bin>      -> per frame: int64 n_bits, int8 bits[n_bits]
//   test_dropin fic <soft.i8> <out.bin>                     -> per 2304-symbol group: 96 decoded bytes + u64 error
//   test_dropin ficdec <nb_bits> <nb_fibs> <soft.i8> <out.bin>   (FIC_Decoder as BasicRadio drives it, basic_radio.cpp:51-56)
//                                                           -> per CRC-valid FIB: int32 group, int32 n, n data bytes
//   test_dropin mscdec <start> <length> <is_uep> <uep_index> <eep_level> <eep_type_b> <cif_bits> <cifs.i8> <out.bin>
//                                                           -> per CIF: int32 n_bytes, bytes (MSC_Decoder::DecodeCIF)
#include <complex>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "dab/algorithms/dab_viterbi_decoder.h"
#include "dab/fic/fic_decoder.h"
#include "dab/msc/msc_decoder.h"
#include "ofdm/ofdm_helpers.h"
#include "dab_b200.h"

static std::vector<char> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static int run_ofdm(int mode, size_t block, const char* in_path, const char* out_path) {
    auto raw = slurp(in_path);
    const size_t n = raw.size() / sizeof(std::complex<float>);
    const auto* iq = reinterpret_cast<const std::complex<float>*>(raw.data());
    std::ofstream out(out_path, std::ios::binary);
    auto demod = Create_OFDM_Demodulator(mode, 1);
    demod->GetConfig().sync.impulse_peak_threshold_db = 20.0f;  // mutable config, as examples/basic_radio_app.cpp:268-269 does
    int frames = 0;
    demod->On_OFDM_Frame().Attach([&](tcb::span<const viterbi_bit_t> buf) {
        const int64_t nb = int64_t(buf.size());
        out.write(reinterpret_cast<const char*>(&nb), sizeof(nb));
        out.write(reinterpret_cast<const char*>(buf.data()), buf.size());
        frames++;
    });
    std::vector<std::complex<float>> buffer(block);
    for (size_t off = 0; off < n; off += block) {
        const size_t len = std::min(block, n - off);
        std::memcpy(buffer.data(), iq + off, len * sizeof(std::complex<float>));
        demod->Process(tcb::span<const std::complex<float>>(buffer.data(), len));
    }
    std::printf("frames=%d read=%d desync=%d state=%d fine=%.9g coarse=%.9g bits0=%d\n", frames, demod->GetTotalFramesRead(),
                demod->GetTotalFramesDesync(), int(demod->GetState()), demod->GetFineFrequencyOffset(), demod->GetCoarseFrequencyOffset(),
                int(demod->GetFrameDataBits()[0]));
    return 0;
}

static int run_fic(const char* in_path, const char* out_path) {
    auto raw = slurp(in_path);
    const auto* soft = reinterpret_cast<const viterbi_bit_t*>(raw.data());
    const size_t nb_encoded_bits = 2304, nb_decoded_bytes = 96;
    uint8_t PI_16[8], PI_15[8], PI_X[8];
    dab_get_puncture_code(16, PI_16);
    dab_get_puncture_code(15, PI_15);
    dab_get_puncture_code(0, PI_X);
    DAB_Viterbi_Decoder vitdec;
    vitdec.set_traceback_length(nb_decoded_bytes * 8);
    std::ofstream out(out_path, std::ios::binary);
    std::vector<uint8_t> decoded(nb_decoded_bytes);
    for (size_t g = 0; (g + 1) * nb_encoded_bits <= raw.size(); g++) {
        // FIC_Decoder::DecodeFIBGroup, fic_decoder.cpp:74-87
        vitdec.reset();
        tcb::span<const viterbi_bit_t> buf(soft + g * nb_encoded_bits, nb_encoded_bits);
        size_t N;
        N = vitdec.update(buf, tcb::span<const uint8_t>(PI_16, 8), 128 * 21);
        buf = buf.subspan(N);
        N = vitdec.update(buf, tcb::span<const uint8_t>(PI_15, 8), 128 * 3);
        buf = buf.subspan(N);
        N = vitdec.update(buf, tcb::span<const uint8_t>(PI_X, 6), 24);
        buf = buf.subspan(N);
        if (!buf.empty() || vitdec.get_current_decoded_bit() != 774) return 3;
        const uint64_t error = vitdec.chainback(decoded);
        out.write(reinterpret_cast<const char*>(decoded.data()), decoded.size());
        out.write(reinterpret_cast<const char*>(&error), sizeof(error));
    }
    // underrun behaves like the reference: update() consumes nothing and returns 0
    vitdec.reset();
    if (vitdec.update(tcb::span<const viterbi_bit_t>(soft, 10), tcb::span<const uint8_t>(PI_16, 8), 128) != 0) return 4;
    return 0;
}

static int run_ficdec(size_t nb_bits, size_t nb_fibs, const char* in_path, const char* out_path) {
    auto raw = slurp(in_path);
    const auto* soft = reinterpret_cast<const viterbi_bit_t*>(raw.data());
    std::ofstream out(out_path, std::ios::binary);
    FIC_Decoder fic_decoder(nb_bits, nb_fibs);
    int32_t group = 0;
    fic_decoder.OnFIB().Attach([&](tcb::span<const uint8_t> buf) {
        const int32_t n = int32_t(buf.size());
        out.write(reinterpret_cast<const char*>(&group), sizeof(group));
        out.write(reinterpret_cast<const char*>(&n), sizeof(n));
        out.write(reinterpret_cast<const char*>(buf.data()), buf.size());
    });
    for (size_t g = 0; (g + 1) * nb_bits <= raw.size(); g++, group++)
        fic_decoder.DecodeFIBGroup(tcb::span<const viterbi_bit_t>(soft + g * nb_bits, nb_bits), g % 4);
    return 0;
}

static int run_mscdec(char** a) {
    Subchannel sub(7);
    sub.start_address = subchannel_addr_t(std::atoi(a[0]));
    sub.length = subchannel_size_t(std::atoi(a[1]));
    sub.is_uep = std::atoi(a[2]) != 0;
    sub.uep_prot_index = uep_protection_index_t(std::atoi(a[3]));
    sub.eep_prot_level = eep_protection_level_t(std::atoi(a[4]));
    sub.eep_type = std::atoi(a[5]) ? EEP_Type::TYPE_B : EEP_Type::TYPE_A;
    const size_t cif_bits = size_t(std::atol(a[6]));
    auto raw = slurp(a[7]);
    const auto* soft = reinterpret_cast<const viterbi_bit_t*>(raw.data());
    std::ofstream out(a[8], std::ios::binary);
    MSC_Decoder msc_decoder(sub);
    for (size_t c = 0; (c + 1) * cif_bits <= raw.size(); c++) {
        auto bytes = msc_decoder.DecodeCIF(tcb::span<const viterbi_bit_t>(soft + c * cif_bits, cif_bits));
        const int32_t n = int32_t(bytes.size());
        out.write(reinterpret_cast<const char*>(&n), sizeof(n));
        out.write(reinterpret_cast<const char*>(bytes.data()), bytes.size());
    }
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc == 6 && std::string(argv[1]) == "ofdm") return run_ofdm(std::atoi(argv[2]), size_t(std::atol(argv[3])), argv[4], argv[5]);
        if (argc == 4 && std::string(argv[1]) == "fic") return run_fic(argv[2], argv[3]);
        if (argc == 6 && std::string(argv[1]) == "ficdec") return run_ficdec(size_t(std::atol(argv[2])), size_t(std::atol(argv[3])), argv[4], argv[5]);
        if (argc == 11 && std::string(argv[1]) == "mscdec") return run_mscdec(argv + 2);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    std::fprintf(stderr, "usage: test_dropin ofdm <mode> <block> <iq.c64> <out.bin> | fic <soft.i8> <out.bin>\n");
    return 1;
}
