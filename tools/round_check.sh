#!/bin/bash
# Everything round 2 wants judged, in ONE gpurun call on 1xB200 (about 10 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round_check.sh r2f'
# outputs under gpurun_out/<tag>_*: the ones kept are copied into profiles/.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 300 gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
print("bench:", d["value"], "ms/step  e2e", d["e2e"]["value"], "seq", d["sequential_phases_ms"], d["verified"], d["breakdown_ms_per_step"], "launches", d["gpu_launches"])
print("roofline:", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["int_pipe"])
k = d["extra"]["k20"]; print("k20:", k["value"], k["e2e"]["value"], k["verified"], k["sequential_phases_ms"], k["breakdown_ms_per_step"])
print({x: d["extra"][x] for x in d["extra"] if x != "k20"}); print(d.get("e2e_host_abi", {}).get("value"), d.get("stages"), d.get("cpu_baseline", {}).get("value"))
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference.json 2> /dev/null
# launch list of two bench steps (ncu serialises every launch and every graph node; eager launches so that each kernel is its own row)
SB_MSM_GRAPH=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-k20 --no-host-abi --no-extra --no-verify > gpurun_out/${tag}_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_bench_launches.csv | tee gpurun_out/${tag}_bench_launches_summary.txt | tail -32
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  SB_MSM_GRAPH=0 timeout 300 ncu --set full --clock-control none -k regex:"$rx" --launch-skip $skip -c $cnt -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap ${tag}_step_acc_ct "k_accumulate|sb_ct" 6 6 python tools/shard_profile.py --world 1 --steps 1 --warmup 1
cap ${tag}_tail_w8 "k_fixup|k_rowcol_coop|k_weighted_coop|k_exchange_combine" 12 12 python tools/shard_profile.py --world 8 --steps 1 --warmup 1
cap ${tag}_ntt "k_ntt_pass" 6 2 python tools/quick_ntt_timing.py 20
python tools/ncu_summary.py gpurun_out/${tag}_step_acc_ct_raw.csv gpurun_out/${tag}_tail_w8_raw.csv gpurun_out/${tag}_ntt_raw.csv | tee gpurun_out/${tag}_ncu_full_summary.txt | cut -c1-220
du -sh gpurun_out
