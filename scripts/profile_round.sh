#!/bin/bash
# One round of profiling evidence on the GPU box (run through gpurun; results land in gpurun_out/, small files only):
#   1. the plain run (must exit 0 without ncu), 2. the ncu launch list of the same command,
#   3. `ncu --set full` of the conv kernels -> summary + conv_traffic.json, 4. the same for the bandwidth kernels changed this round.
# The .ncu-rep files are summarised on the box and deleted (gpurun copies back at most 64 MiB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --quick --no-graph --no-stock --no-cpu"
$CMD > gpurun_out/plain_r2f.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_r2f.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2_final.csv $CMD > gpurun_out/ncu_launches_r2f.log 2>&1
echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel" -s 30 -c 26 -f -o /tmp/prof_conv_halo_r2f $CMD > gpurun_out/ncu_full_r2f.log 2>&1
echo "conv full rc $?"
python scripts/ncu_summary.py /tmp/prof_conv_halo_r2f.ncu-rep > gpurun_out/conv_halo_r2f_ncu_summary.txt 2>&1
python scripts/ncu_traffic.py /tmp/prof_conv_halo_r2f.ncu-rep "bench.py --steps 2 --warmup 1 --quick --no-graph, fp16 mode, round 2 final" > gpurun_out/ncu_traffic_r2f.log 2>&1
cp profiles/conv_traffic.json gpurun_out/conv_traffic.json
ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_reduce_pool|bn_bwd_apply_pool|tail_bwd_fused|upsample2" -s 10 -c 12 -f -o /tmp/prof_misc_r2f $CMD > gpurun_out/ncu_misc_r2f.log 2>&1
echo "misc full rc $?"
python scripts/ncu_summary.py /tmp/prof_misc_r2f.ncu-rep > gpurun_out/misc_kernels_r2f_ncu_summary.txt 2>&1
gzip -f gpurun_out/launches_r2_final.csv
ls -la gpurun_out/ | tail -12
