#!/bin/bash
# Viterbi / ensemble timing line of bench.py (not part of the product)
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sustained 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().split(chr(10))[-1])
v = d['viterbi']
print('viterbi ms', v['ms_per_launch'], 'ensemble ms', v['ensemble']['ms_per_call'], v['bit_exact_vs_oracle'], '|', v['ensemble']['bit_exact_vs_oracle'])
"
