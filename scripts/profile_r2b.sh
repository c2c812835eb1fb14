#!/bin/bash
# Round-2 final profile pass (run under gpurun; one call): launch list of the default bench command + ncu --set full of
# the kernels DESIGN.md quotes.  Each ncu command runs only after the same command line exited 0 without ncu.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu"
$B > gpurun_out/r2b_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/r2b_ncu_launches.log 2>&1
K='window_score|topk_cluster|gather_rows|h2o_kernel|ada_phase'
for job in "one_layer:6:6:python scripts/one_layer.py auto 1 16" "cfg5_layer:6:6:python scripts/cfg5_layer.py" "h2o_layer:1:1:python scripts/h2o_layer.py 8192" \
           "one_dyn:3:1:python scripts/one_dyn.py 32768 32 8 dyn" "one_half:3:1:python scripts/one_dyn.py 32768 32 8 half"; do
  name=${job%%:*}; rest=${job#*:}; skip=${rest%%:*}; rest=${rest#*:}; cnt=${rest%%:*}; cmd=${rest#*:}
  $cmd > gpurun_out/r2b_${name}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s $skip -c $cnt -f -o gpurun_out/r2b_${name} $cmd > gpurun_out/r2b_${name}_ncu.log 2>&1
done
ls -la gpurun_out | grep r2b
