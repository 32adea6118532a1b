#!/bin/bash
# one `ncu --set full` capture per named kernel (first 2 launches after the warm-up step), bench workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > gpurun_out/plain_ncu.log 2>&1 || { echo "plain run failed"; exit 1; }
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 2 -f -o "gpurun_out/prof_${k}" \
      python bench.py --steps 2 --warmup 1 --quick > "gpurun_out/ncu_${k}.log" 2>&1
  echo "$k rc $?"
done
