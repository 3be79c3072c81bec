# the round-end single-GPU measurements: the driver's default bench line, the reference arm, the ncu launch list and DRAM traffic
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_1gpu_10M_final.json 2> gpurun_out/r02_bench_1gpu_final.err; cut -c1-600 gpurun_out/r02_bench_1gpu_10M_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>&1; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_1M.csv \
    python bench.py --n-traj 1000000 --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/ncu_launch_bench.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dp_ensemble_kernel -c 1 --csv --log-file gpurun_out/r02_dram_traffic_10M.csv \
    python bench.py --steps 1 --warmup 0 --no-extra --no-cpu --no-e2e > gpurun_out/ncu_dram_bench.log 2>&1
tail -3 gpurun_out/r02_dram_traffic_10M.csv
