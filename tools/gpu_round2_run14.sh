set -x
N=${1:-4}
DEB_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"library_call_ms_rank0": [0-9.]*' gpurun_out/r02_bench_${N}gpu.json
grep "deb timing" gpurun_out/r02_bench_${N}gpu.err | tail -8 | sed 's/.*first block done/first block done/; s/; kernel.*//'
python -m pytest tests/test_abi9_gpu.py tests/test_full_size_gpu.py -q -x -k "device_list or host_path_2m or sde" > gpurun_out/r02_t_${N}gpu.log 2>&1; tail -3 gpurun_out/r02_t_${N}gpu.log
