#!/bin/bash
# One gpurun call that produces everything a round needs judged (about 4 GPU-minutes on 1xB200):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round_check.sh r2'
# outputs under gpurun_out/<tag>_*: copy what you keep into profiles/.
tag=${1:-rN}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 300 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_n1.json"))
print("bench:", d["value"], "ms/step  e2e", d["e2e"]["value"], d["breakdown_ms_per_step"], "int_pipe", d["roofline"]["int_pipe"]["frac"])
PY
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference.json 2>&1
timeout 100 python tools/check_large_msm.py 23 2>&1 | tail -4 | tee gpurun_out/${tag}_large_msm.txt
# the dominant kernel, one full capture (largest commit of the step)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -c 1 -o gpurun_out/${tag}_accumulate \
    python tools/affine_one.py 0 16 1 > gpurun_out/${tag}_ncu.log 2>&1
# launch list of two bench steps (ncu serialises every launch: ~2 minutes)
if [ "$2" = "launches" ]; then
    timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
fi
