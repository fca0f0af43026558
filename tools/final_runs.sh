#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): tests, both bench arms, launch list, frame-kernel ncu capture.
# usage: tools/final_runs.sh <tag> [ncu]     -> gpurun_out/<tag>_*   (the three ncu --set full reports, ~20 MB each, only with `ncu`:
# gpurun copies back at most 64 MiB)
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; tail -c 600 $out/${tag}_bench_reference.json
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; tail -c 300 $out/${tag}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ofdm_ -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-viterbi --no-sustained > $out/${tag}_launches.log 2>&1
[ "$2" = ncu ] || { ls -la $out | tail -12; exit 0; }
ncu --set full --clock-control none --import-source on -k regex:ofdm_frame_v3 -s 40 -c 1 -f -o $out/${tag}_frame_way python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-viterbi --no-sustained > $out/${tag}_ncu_frame_way.log 2>&1
DAB_B200_PIPELINE_WAYS=1 ncu --set full --clock-control none --import-source on -k regex:ofdm_frame_v3 -s 14 -c 1 -f -o $out/${tag}_frame_step python tools/ncu_target.py 9 > $out/${tag}_ncu_frame_step.log 2>&1
ls -la $out | tail -12
DAB_B200_PIPELINE_WAYS=1 ncu --set full --clock-control none --import-source on -k regex:ofdm_control -s 21 -c 3 -f -o $out/${tag}_control python tools/ncu_target.py 9 > $out/${tag}_ncu_control.log 2>&1
