// Stand-in for the reference's src/viterbi_config.h: the soft-bit contract between the demodulator and the decoder.
#pragma once
#include <stdint.h>
typedef int8_t viterbi_bit_t;
static constexpr viterbi_bit_t SOFT_DECISION_VITERBI_HIGH = +127;
static constexpr viterbi_bit_t SOFT_DECISION_VITERBI_LOW = -127;
static constexpr viterbi_bit_t SOFT_DECISION_VITERBI_PUNCTURED = 0;
