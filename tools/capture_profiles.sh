#!/bin/bash
# Round-end evidence, one gpurun call: bench lines, the ncu launch list of the bench command and
# `--set full` captures of the dominant kernels.  Everything lands in gpurun_out/ (copy what is to be
# judged into profiles/, named per round).  A command runs under ncu only after the same command line
# has exited 0 without it (`&&` directly before).
set -x
cd "$(dirname "$0")/.."
R=${ROUND:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err || exit 1
timeout 900 python bench.py --impl reference --steps 7 --warmup 2 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_short.json 2> gpurun_out/${R}_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${R}_ncu_launches.log 2>&1
timeout 120 python tools/ot_tune.py c3 bf16 > gpurun_out/${R}_ot_tune_c3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ot_stream" -s 32 -c 1 \
    -o gpurun_out/prof_${R}_ot_stream python tools/ot_tune.py c3 bf16 > gpurun_out/${R}_ncu_ot.log 2>&1
timeout 120 python tools/con_tune.py c3 bf16 > gpurun_out/${R}_con_tune_c3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:umma_gemm|bwd_|normalize_bwd|prep_all|fwd_items" -s 100 -c 20 \
    -o gpurun_out/prof_${R}_gemm_chain python tools/con_tune.py c3 bf16 > gpurun_out/${R}_ncu_gemm.log 2>&1
timeout 120 python tools/ot_tune.py c4 bf16 > gpurun_out/${R}_ot_tune_c4.log 2>&1
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_c2.json 2> gpurun_out/${R}_bench_c2.err
tail -n 2 gpurun_out/${R}_con_tune_c3.log; tail -n 3 gpurun_out/${R}_ot_tune_c3.log; tail -n 2 gpurun_out/${R}_ot_tune_c4.log
