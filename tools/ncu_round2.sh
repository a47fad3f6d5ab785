# ncu --set full captures of round 2's kernels; the reports are reduced to text on the box (raw CSV + details page) because
# gpurun returns at most 64 MiB.  Run under gpurun on ONE GPU:  bash tools/ncu_round2.sh
set -x
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --clock-control none -k regex:"$rx" --launch-skip $skip -c $cnt -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
# 1. dominant kernel + compiled cross terms inside the k = 17 step
cap r2_step_acc_ct "k_accumulate|sb_ct" 6 6 python tools/shard_profile.py --world 1 --steps 1 --warmup 1
# 2. NTT passes at k = 20 and k = 17
cap r2_ntt "k_ntt_pass" 6 3 python tools/quick_ntt_timing.py 20,17
# 3. Protogalaxy: compiled leaf kernel (blend of 2 traces), beta tree, witness fold at k = 17
cap r2_pg "sb_ct|k_beta_tree|k_lincomb" 40 4 python bench.py --workload cyclefold_poseidon --k 17 --steps 1 --warmup 1 --no-verify
# 4. the commitment tail on the cooperative group law, one rank's share of the 8-GPU step
cap r2_tail_w8 "k_fixup_coop|k_rowcol_coop|k_weighted_coop|k_exchange_combine" 8 8 python tools/shard_profile.py --world 8 --steps 1 --warmup 1
du -sh gpurun_out
