#!/bin/bash
# Round-end evidence, one gpurun call: bench lines, the ncu launch list of the bench command and
# `--set full` captures of the dominant kernels.  Everything lands in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || exit 1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 120 python tools/con_tune.py c3 bf16 > gpurun_out/con_tune_c3.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:EpiStats" -s 30 -c 2 \
    -o gpurun_out/prof_r01_epistats python tools/con_tune.py c3 bf16 > gpurun_out/ncu_epistats.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:EpiGrad|EpiStore|normalize_bwd|prep_all" -s 24 -c 8 \
    -o gpurun_out/prof_r01_bwd python tools/con_tune.py c3 bf16 > gpurun_out/ncu_bwd.log 2>&1
timeout 120 python tools/ot_tune.py c3 bf16 > gpurun_out/ot_tune_c3.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ot_" -s 30 -c 4 \
    -o gpurun_out/prof_r01_ot python tools/ot_tune.py c3 bf16 > gpurun_out/ncu_ot.log 2>&1
tail -n 2 gpurun_out/con_tune_c3.log; tail -n 2 gpurun_out/ot_tune_c3.log
