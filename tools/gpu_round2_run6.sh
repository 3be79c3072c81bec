set -x
python -m pytest tests/test_abi9_gpu.py -q -x -k "device_list" > gpurun_out/r02_t_2gpu_b.log 2>&1; tail -3 gpurun_out/r02_t_2gpu_b.log
DEB_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_bench_2gpu_dbg.json 2> gpurun_out/r02_bench_2gpu_dbg.err
grep "deb timing" gpurun_out/r02_bench_2gpu_dbg.err | tail -8; grep -o '"e2e": {[^}]*}' gpurun_out/r02_bench_2gpu_dbg.json | cut -c1-400
DEB_DEBUG_TIMING=1 python bench.py --gpus 2 --single-process > gpurun_out/r02_bench_2gpu_single_dbg.json 2> gpurun_out/r02_bench_2gpu_single_dbg.err
grep "deb timing" gpurun_out/r02_bench_2gpu_single_dbg.err | tail -4
