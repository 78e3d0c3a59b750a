# round 2, GPU run 5: stratified kernels after the common-factor / nu^2-polynomial / step-size changes (parity + bench), ncu of the
# cooperative kernel, per-launch timings of the config-5 shard
set -x
P=gpurun_out/r2e
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or seeded or absorption or config2_every or config3_every or config2_full or edge or determinism" > ${P}_pytest.log 2>&1; tail -4 ${P}_pytest.log
for W in config2 config3 config1; do
  timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline --no-strong > ${P}_$W.json 2> ${P}_$W.err
  python - <<PY
import json
d=json.loads(open("${P}_$W.json").read().strip().splitlines()[-1]); print("RESULT $W", round(d["ms_per_step"],1), "ms", round(d["rk4_steps_per_sec"]/1e9,3), "Gsteps/s frac", round(d["roofline"]["frac"],4), "occ", d["config"]["lane_occupancy"], "e2e", round(d["e2e"]["value"]))
PY
done
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
run() { name=$1; wl=$2; shard=$3; shift 3; env "$@" timeout 400 $B --workload $wl $shard > ${P}_$name.json 2> ${P}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("${P}_$name.json").read().strip().splitlines()[-1]); print("RESULT $name", round(d["ms_per_step"]), "ms", round(d["rk4_steps_per_sec"]/1e6), "Msteps/s frac", round(d["roofline"]["frac"],3), "occ", d["config"]["lane_occupancy"], d["config"].get("schedule"))
except Exception as e: print("RESULT $name failed", e)
PY
}
run c5s8_coop config5 "--shard-of 0/8" A=1
run c5s8_serial config5 "--shard-of 0/8" GEOAC_B200_LONG_WIDTH=32
run c5s8_serial_idx config5 "--shard-of 0/8" GEOAC_B200_LONG_WIDTH=32 GEOAC_B200_RD_GROUP=0
E="GEOAC_B200_LONG_ALPHA=1 GEOAC_B200_LONG_SM_PCT=90 GEOAC_B200_LPT=2"
env $E timeout 200 $B --workload config4s > ${P}_c4s_plain.json 2> ${P}_c4s_plain.err && \
env $E timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_coop_kernel -c 1 -f -o ${P}_prof_coop $B --workload config4s > ${P}_ncu_coop.log 2>&1
ncu -i ${P}_prof_coop.ncu-rep --page details > ${P}_ncu_coop_details.txt 2>/dev/null
ncu -i ${P}_prof_coop.ncu-rep --page source --csv > ${P}_ncu_coop_source.csv 2>/dev/null
rm -f ${P}_prof_coop.ncu-rep
ls -la gpurun_out | tail -12
