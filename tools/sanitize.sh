#!/bin/bash
# compute-sanitizer passes over the library's kernels (run on the GPU box through gpurun); logs land in gpurun_out/<tag>_sanitizer_*.log
# memcheck / initcheck / synccheck: smoke() + the back-to-back probe (Modes I and II: every OFDM kernel, pipeline ways)
# + the Viterbi / ensemble smoke paths.  racecheck (shared-memory hazards, ~100x slower): 64 streams.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
export PROBE_STEP=16 PROBE_REPS=1
for tool in memcheck initcheck synccheck; do
  log=$out/${tag}_sanitizer_${tool}.log
  : > $log
  for cmd in "python __graft_entry__.py smoke" "python tools/async_probe.py 2 4096" "python tools/async_probe.py 1 196608" \
             "python -m pytest tests/test_ensemble_gpu.py -x -q -m gpu -k demod_to_bytes"; do
    echo "##### compute-sanitizer --tool $tool $cmd" >> $log
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 $cmd 2>&1 | grep -v "^$" | tail -12 >> $log
  done
done
log=$out/${tag}_sanitizer_racecheck.log
: > $log
export PROBE_STREAMS=64 PROBE_STEP=8
for cmd in "python tools/async_probe.py 2 4096" "python tools/async_probe.py 1 196608" "python __graft_entry__.py smoke"; do
  echo "##### compute-sanitizer --tool racecheck $cmd" >> $log
  timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 $cmd 2>&1 | grep -v "^$" | tail -12 >> $log
done
for f in $out/${tag}_sanitizer_*.log; do echo "== $f"; tail -n 5 $f; done
