#!/bin/bash
# torchrun --no-python bash benchmarks/rank0_under_ncu.sh <bench args>: rank 0 runs under ncu (duration-only pass), the others plain
if [ "$RANK" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${NCU_SKIP:-900} -c ${NCU_COUNT:-520} --csv --log-file gpurun_out/r2_launches_n2.csv python bench.py "$@"
else
  exec python bench.py "$@"
fi
