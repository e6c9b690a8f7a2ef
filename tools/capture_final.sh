#!/bin/bash
# One gpurun call with the round's final evidence: GPU suite, smoke(), bench lines (default c3, c2, c4, reference arm),
# the ncu launch list of the default bench command and a `--set full` capture of one pass of the contrastive kernel chain.
# A command runs under ncu only after the same command line has exited 0 without it.
set -x
cd "$(dirname "$0")/.."
R=${ROUND:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -n 15 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -n 2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_c2.json 2> gpurun_out/${R}_bench_c2.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err
if [ -z "$SKIP_REF" ]; then
timeout 600 python bench.py --impl reference --steps 7 --warmup 2 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
fi
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_short.json 2> gpurun_out/${R}_bench_short.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches_c3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${R}_ncu_launches_c3.log 2>&1
timeout 120 python tools/con_tune.py c3 bf16 > gpurun_out/${R}_con_tune_c3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:umma_gemm|bwd_|normalize_bwd|prep_all|fwd_items" -s 100 -c 18 \
    -o gpurun_out/prof_${R}_gemm_chain python tools/con_tune.py c3 bf16 > gpurun_out/${R}_ncu_gemm.log 2>&1
for f in default c2 c4; do python - <<EOF
import json
d = json.loads(open("gpurun_out/${R}_bench_$f.json").read().strip().splitlines()[-1])
print("$f", "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["ms"], 4),
      "secondary", round(d["roofline_secondary"]["frac"], 3), round(d["roofline_secondary"]["ms"], 4), "launches/step", d.get("gpu_launches_per_step"), "e2e", round(d["e2e"]["value"]))
EOF
done
