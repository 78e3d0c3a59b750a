# round 2, GPU run 8: full-occupancy section captures of the range-dependent kernels (single launch: under ncu kernels run one
# after another, so the exclusive long-region launch would drain both regions and leave the main launch empty), listing tool
set -x
P=gpurun_out/r2z
export GEOAC_B200_EXCLUSIVE=0
SEC="--section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section InstructionStats --section LaunchStats --section Occupancy"
cap() { name=$1; shift; timeout 600 ncu $SEC --clock-control none -k regex:trace_kernel -s 1 -c 1 "$@" > ${P}_ncu_$name.txt 2>&1; grep -E "Duration|L1/TEX Hit|Issue Slots" ${P}_ncu_$name.txt; }
cap Eq3DRD_config4_full_occupancy python bench.py --workload config4 --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-e2e --rays-cap 40000
cap EqGlobalRD_config5_full_occupancy python bench.py --workload config5 --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-e2e --rays-cap 40000 --ray-limit 200
unset GEOAC_B200_EXCLUSIVE
timeout 300 python tools/list_near_threshold.py --workload config2 --every 100 --limit 60 > ${P}_near_threshold_config2.txt 2>&1; head -4 ${P}_near_threshold_config2.txt
timeout 300 python tools/list_near_threshold.py --workload config1 --every 1 --limit 60 > ${P}_near_threshold_config1.txt 2>&1; head -4 ${P}_near_threshold_config1.txt
