#!/bin/bash
# One gpurun call: the GPU suite, smoke(), the default and c4 bench lines, the ncu launch list of the c4 bench
# command and `--set full` captures of the wide-plan OT kernels (csrc/ot_wide.cu + the ragged IPOT solver).
# A command runs under ncu only after the same command line has exited 0 without it.
set -x
cd "$(dirname "$0")/.."
R=${ROUND:-r02b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -n 5 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -n 2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err
timeout 200 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_c4_short.json 2> gpurun_out/${R}_bench_c4_short.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${R}_launches_c4.csv \
    python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${R}_ncu_launches_c4.log 2>&1
timeout 120 python tools/ot_tune.py c4 bf16 > gpurun_out/${R}_ot_tune_c4.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ot_wide|ot_ipot" -s 48 -c 3 \
    -o gpurun_out/prof_${R}_ot_c4 python tools/ot_tune.py c4 bf16 > gpurun_out/${R}_ncu_ot_c4.log 2>&1
timeout 120 python tools/ot_tune.py c3 bf16 > gpurun_out/${R}_ot_tune_c3.log 2>&1
tail -n 1 gpurun_out/${R}_ot_tune_c4.log; tail -n 2 gpurun_out/${R}_ot_tune_c3.log
head -c 1500 gpurun_out/${R}_bench_c4.json
