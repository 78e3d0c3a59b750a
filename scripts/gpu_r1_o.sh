for K in 8 16 32 64; do
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
GEOAC_B200_SCOUT_COARSE=$K timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/r1o_launches_$K.csv $B > gpurun_out/r1o_$K.log 2>&1
echo K=$K; grep -E "scout|trace_kernel" gpurun_out/r1o_launches_$K.csv | head -4 | awk -F'","' '{print substr($5,1,40), $NF}'
python -c "
import json
d=json.loads(open('gpurun_out/r1o_$K.log').read().strip().splitlines()[-1]); print('occ', d['config']['lane_occupancy'])"
done
