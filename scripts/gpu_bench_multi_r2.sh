# bench.py on N GPUs exactly as the driver launches it (torchrun, one rank per GPU): config 2 weak scaling, the single-process
# multi-device sub-record (geoac_trace_multi from rank 0) and the config-5 strong-scaling pass.   bash scripts/gpu_bench_multi_r2.sh N
N=${1:-2}
set -x
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv | head -10
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 \
  > gpurun_out/r2z_bench_gpus$N.json 2> gpurun_out/r2z_bench_gpus$N.err
tail -c 1500 gpurun_out/r2z_bench_gpus$N.json
tail -5 gpurun_out/r2z_bench_gpus$N.err
