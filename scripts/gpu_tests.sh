#!/bin/bash
# Run every GPU test file in its own process (a CUDA fault in one file must not poison the others),
# each under its own timeout; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
status=0
for f in tests/test_gpu_probe.py tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_extras.py tests/test_gpu_parallel.py "$@"; do
  name=$(basename "$f" .py)
  echo "=== $f"
  timeout 900 python -m pytest "$f" -q -rA -m gpu --maxfail=25 --no-header -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  rc=$?
  tail -n 25 "gpurun_out/${name}.log"
  echo "=== $f exit $rc"
  [ $rc -ne 0 ] && status=1
done
exit $status
