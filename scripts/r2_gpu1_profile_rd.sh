# round 2, GPU run 1: full-occupancy ncu captures of the range-dependent kernels and EqGlobal (before the rework),
# and the per-ray step counts of config 5 in full (for the long-ray scheduling design)
set -x
P=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > ${P}_smi.csv
for W in "config4 20000 0" "config5 20000 300" "config3 50000 0"; do
  set -- $W
  B="python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu-baseline --rays-cap $2"
  if [ "$3" != "0" ]; then B="$B --ray-limit $3"; fi
  timeout 600 $B > ${P}_plain_$1.json 2> ${P}_plain_$1.err && \
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -f -o ${P}_prof_$1 $B > ${P}_ncu_$1.log 2>&1
  tail -c 400 ${P}_plain_$1.json
  ncu -i ${P}_prof_$1.ncu-rep --page details > ${P}_ncu_details_$1.txt 2>/dev/null
  ncu -i ${P}_prof_$1.ncu-rep --page raw --csv > ${P}_ncu_raw_$1.csv 2>/dev/null
done
timeout 900 python scripts/r2_config5_steps.py > ${P}_config5_steps.log 2>&1; tail -20 ${P}_config5_steps.log
ls -la gpurun_out | tail -20
