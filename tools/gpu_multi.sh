# usage: bash tools/gpu_multi.sh N   — the round-end multi-GPU measurements on one box with N GPUs
N=${1:-2}
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "ms_per_step")}, l["e2e"])
PY
DEB_DEBUG_TIMING=1 python bench.py --gpus $N --single-process --steps 5 --warmup 3 --no-extra \
    > gpurun_out/r02_bench_${N}gpu_single_process.json 2> gpurun_out/r02_bench_${N}gpu_single_process.err
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_${N}gpu_single_process.json").read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "ms_per_step")}, l["e2e"])
PY
grep -E "all devices done|shard finished" gpurun_out/r02_bench_${N}gpu_single_process.err | tail -$((N+1))
python -m pytest tests -m gpu -q --timeout 600 -k "device or Device or shard" > gpurun_out/r02_t_${N}gpu.log 2>&1; tail -3 gpurun_out/r02_t_${N}gpu.log
