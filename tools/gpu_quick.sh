#!/bin/bash
# Quick GPU iteration: parity tests, bench (no CPU leg), optional ncu capture of some kernels.
# Usage: bash tools/gpu_quick.sh <tag> ["<profile_ops args>" "<kernel regex>" [<skip> <count>]]
tag=${1:-q}; ops=${2:-}; rx=${3:-}; skip=${4:-0}; cnt=${5:-6}
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_$tag.log
python bench.py --no-cpu --steps 10 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
if [ -n "$rx" ]; then
  python tools/profile_ops.py $ops > $out/plain_profile_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f \
      -o $out/prof_$tag python tools/profile_ops.py $ops > $out/ncu_full_$tag.log 2>&1
  echo "ncu exit $?"
fi
