set -x
tools/microbench/fp64_operands > gpurun_out/r02_fp64_operands.jsonl 2>&1
python -m pytest tests/test_abi9_gpu.py -q -x -k "device_list or statistics or row_major" > gpurun_out/r02_t_2gpu.log 2>&1; tail -5 gpurun_out/r02_t_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/host_bw_probe.py > gpurun_out/r02_hostbw_2gpu.json 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
python bench.py --gpus 2 --single-process > gpurun_out/r02_bench_2gpu_single.json 2> gpurun_out/r02_bench_2gpu_single.err
tail -c 1200 gpurun_out/r02_bench_2gpu.json; tail -c 800 gpurun_out/r02_bench_2gpu_single.json; tail -3 gpurun_out/r02_bench_2gpu_single.err; cat gpurun_out/r02_fp64_operands.jsonl
