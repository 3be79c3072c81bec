set -x
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/host_bw_probe.py > gpurun_out/r02_hostbw_${N}gpu.json 2>&1
DEB_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
DEB_DEBUG_TIMING=1 python bench.py --gpus $N --single-process > gpurun_out/r02_bench_${N}gpu_single.json 2> gpurun_out/r02_bench_${N}gpu_single.err
python -m pytest tests/test_abi9_gpu.py -q -x -k "device_list" > gpurun_out/r02_t_${N}gpu.log 2>&1; tail -3 gpurun_out/r02_t_${N}gpu.log
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"library_call_ms_rank0": [0-9.]*' gpurun_out/r02_bench_${N}gpu.json gpurun_out/r02_bench_${N}gpu_single.json
grep "deb timing" gpurun_out/r02_bench_${N}gpu.err | tail -16
grep "deb timing" gpurun_out/r02_bench_${N}gpu_single.err | tail -8
grep -o '"results": {.*}, "topology"' gpurun_out/r02_hostbw_${N}gpu.json | cut -c1-900
