set -x
python bench.py > gpurun_out/r02_bench_10M_a.json 2> gpurun_out/r02_bench_10M_a.err; tail -c 600 gpurun_out/r02_bench_10M_a.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dp_ensemble -c 1 --csv --log-file gpurun_out/r02_dram_traffic_10M.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-extra > gpurun_out/r02_dram_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_1M.csv python bench.py --n-traj 1000000 --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dp_ensemble -c 1 -o gpurun_out/r02_dp_a python bench.py --n-traj 1000000 --steps 1 --warmup 0 --no-e2e --no-cpu --no-extra > gpurun_out/r02_ncu_full.log 2>&1
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:sde_ensemble -c 4 --csv --log-file gpurun_out/r02_ncu_sde_inst.csv python tools/bench_configs.py --c3 0 --c5 0 --c4 10000000 --reps 1 > gpurun_out/r02_sde_ncu.log 2>&1
ls -la gpurun_out | tail -12
