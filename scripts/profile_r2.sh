#!/bin/bash
# Round-2 profile pass (run under gpurun; one call): launch list of the default bench command + ncu --set full of the
# kernels DESIGN.md quotes.  Each ncu command runs only after the same command line exited 0 without ncu.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu"
$B > gpurun_out/r2_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
for job in "one_layer:python scripts/one_layer.py auto 1 16" "cfg5_layer:python scripts/cfg5_layer.py" "h2o_layer:python scripts/h2o_layer.py 8192" "one_dyn:python scripts/one_dyn.py"; do
  name=${job%%:*}; cmd=${job#*:}
  $cmd > gpurun_out/r2_${name}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'window_score|topk_cluster|gather_rows|h2o_kernel' -s 6 -c 6 -o gpurun_out/r2_${name} $cmd > gpurun_out/r2_${name}_ncu.log 2>&1
done
ls -la gpurun_out | tail -20
