set -x
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 12 --warmup 3 --no-e2e > gpurun_out/r02_bench_${N}gpu_dr.json 2> gpurun_out/r02_bench_${N}gpu_dr.err
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"kernel_ms_steps_rank0": \[[^]]*\]\|"clocks": {[^}]*}' gpurun_out/r02_bench_${N}gpu_dr.json
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit,temperature.gpu --format=csv
