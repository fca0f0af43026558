#!/bin/bash
# ncu captures of the kernel families (run on the GPU box through gpurun); reports land in gpurun_out/
tag=${1:-r01g}
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:ofdm_frame_v3 -s 8 -c 1 -f -o $out/${tag}_frame python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-viterbi > $out/${tag}_frame.log 2>&1
DAB_B200_PIPELINE_WAYS=1 $NCU -k regex:ofdm_control -s 9 -c 3 -f -o $out/${tag}_control python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-viterbi > $out/${tag}_control.log 2>&1
# the two forms of the Viterbi kernel on the same 77 824-trellis batch: bulk (one trellis per thread) and warp-cooperative
$NCU -k regex:viterbi_lanes_kernel -s 2 -c 1 -f -o $out/${tag}_viterbi_lanes python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_viterbi_lanes.log 2>&1
DAB_B200_VITERBI_LANES=0 $NCU -k regex:viterbi_kernel -s 2 -c 1 -f -o $out/${tag}_viterbi_warp python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_viterbi_warp.log 2>&1
# ensemble decoder: push, de-interleave, Viterbi, commit of one timed call
$NCU -k regex:ens_ -s 28 -c 4 -f -o $out/${tag}_ensemble python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_ensemble.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ofdm_ -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-viterbi > $out/${tag}_launches.log 2>&1
ls -la $out
