#!/bin/bash
# Round-2 evidence on one B200 (run through gpurun from the repo root): bench lines, the ncu launch list of one C3b step, per-launch
# ncu metrics of that step, and `--set full` captures of the kernels DESIGN.md discusses.  Everything lands in gpurun_out/r2/;
# scripts/summarize_r2.py turns it into the tracked summaries under profiles/.  A number printed under ncu is never a bench value.
set -u
O=gpurun_out/r2
mkdir -p $O/prof
M=$(cat scripts/ncu_metrics.txt)
BENCH1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --profile-steps 0 --legs staged"
# 1. bench lines (no profiler)
DCGANSR_PROFILE_DIR=$O/prof timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
DCGANSR_PROFILE_DIR=$O/prof timeout 600 python bench.py --workload C5 --steps 2 --warmup 3 --legs staged,profile --no-extra --no-cpu-baseline --profile-steps 1 > $O/bench_c5.json 2> $O/bench_c5.err
# 2. launch list of the whole run (time only); the summariser cuts out the last step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c3b.csv $BENCH1 > $O/ncu_launch.log 2>&1
# 3. per-launch metrics of the LAST step (the step count before it is read from the launch list)
read SKIP COUNT < <(python scripts/summarize_r2.py --last-step $O/launches_c3b.csv)
echo "last step: skip $SKIP count $COUNT" > $O/last_step.txt
timeout 900 ncu --metrics $M --clock-control none --launch-skip $SKIP --launch-count $COUNT --csv --log-file $O/step_metrics_c3b.csv $BENCH1 > $O/ncu_step.log 2>&1
# 4. --set full single-kernel captures (3rd launch of the op: after the two warm-ups)
cap() {   # name kernel-regex one_layer-args...
  local name=$1 rx=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip 2 --launch-count 1 -f -o $O/full_$name python scripts/one_layer.py "$@" > $O/full_$name.log 2>&1
  ncu -i $O/full_$name.ncu-rep --page raw --csv > $O/full_$name.csv 2>/dev/null
  ls -la $O/full_$name.ncu-rep | awk '{print $5}' >> $O/rep_sizes.txt
}
cap halo_fc48_24_fwd      '^tapconv_halo'      1 48 256 24 4 2 1 128 0
cap halo_fc96_48_fwd_pair '^tapconv_halo'      1 96 128 48 4 2 1 128 0
cap halo_fc64_32_fwd_c2   '^tapconv_halo'      1 64 128 32 4 2 1 64 0
cap wgrad_halo_fc48_24    '^wgrad_halo'        1 48 256 24 4 2 1 128 2
cap tc3_fc256_128_fwd     '^tapconv_tc3'       1 256 64 128 4 2 1 64 0
cap tc3_c128_256_fwd      '^tapconv_tc3'       0 128 128 256 4 2 1 64 0
cap tc2pair_c256_512_fwd  '^tapconv_tc2_pair'  0 256 16 512 4 2 1 256 0
cap wgradpair_fc256_128   '^wgrad_tc_pair'     1 256 64 128 4 2 1 64 2
# the reports stay only while they are small (gpurun_out/ merges back at most 64 MiB)
for f in $O/full_*.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 7000000 ]; then rm -f $f; fi; done
du -sh $O
