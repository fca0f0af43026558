// Stand-in for the reference's src/ofdm/ofdm_params.h for builds OUTSIDE its tree: OFDM_Params is the C ABI's dab_ofdm_params
// (include/dab_b200.h, same fields in the same order), so the two never drift apart.  Inside the reference tree its own header
// is found first and ofdm_demodulator.cpp checks that the layouts agree.
#pragma once
#include "dab_b200.h"
using OFDM_Params = dab_ofdm_params;
