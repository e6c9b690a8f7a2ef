#!/bin/bash
# One gpurun call: the GPU suite with programmatic dependent launch on (CE_PDL=1), the in-process A/B of the bench step
# (tools/pdl_check.py: parity of every replay + CUDA-event timing of both graphs) and the bench lines with it on.
set -x
cd "$(dirname "$0")/.."
R=${ROUND:-r02d}
mkdir -p gpurun_out
CE_PDL=1 timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu_pdl.log 2>&1; tail -n 6 gpurun_out/${R}_pytest_gpu_pdl.log
timeout 200 python tools/pdl_check.py c3 bf16 > gpurun_out/${R}_pdl_check_c3_bf16.log 2>&1; tail -n 5 gpurun_out/${R}_pdl_check_c3_bf16.log
timeout 200 python tools/pdl_check.py c2 bf16 > gpurun_out/${R}_pdl_check_c2_bf16.log 2>&1; tail -n 5 gpurun_out/${R}_pdl_check_c2_bf16.log
timeout 200 python tools/pdl_check.py c4 bf16 > gpurun_out/${R}_pdl_check_c4_bf16.log 2>&1; tail -n 5 gpurun_out/${R}_pdl_check_c4_bf16.log
timeout 200 python tools/pdl_check.py c3 fp32 20 > gpurun_out/${R}_pdl_check_c3_fp32.log 2>&1; tail -n 5 gpurun_out/${R}_pdl_check_c3_fp32.log
CE_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_default_pdl1.json 2> gpurun_out/${R}_bench_default_pdl1.err
CE_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_default_pdl0.json 2> gpurun_out/${R}_bench_default_pdl0.err
for f in pdl1 pdl0; do python - <<PY
import json
d = json.loads(open("gpurun_out/${R}_bench_default_$f.json").read().strip().splitlines()[-1])
print("$f", "ms/step", round(d["ms_per_step"], 4), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["ms"], 4),
      "secondary", round(d["roofline_secondary"]["frac"], 3), round(d["roofline_secondary"]["ms"], 4), "losses", d["losses"])
PY
done
