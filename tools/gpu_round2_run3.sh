set -x
tools/microbench/fp64_issue_mix > gpurun_out/r02_fp64_issue_mix.jsonl 2>&1
bash tools/ab_libs.sh 2000000 > gpurun_out/r02_ab_stcs.log 2>&1
bash tools/ab_libs.sh 2000000 >> gpurun_out/r02_ab_stcs.log 2>&1
DEB200_LIB=$PWD/build/alt/stcs.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dp_ensemble -c 1 --csv --log-file gpurun_out/r02_dram_traffic_10M_stcs.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-extra > gpurun_out/r02_dram_bench_stcs.log 2>&1
cat gpurun_out/r02_ab_stcs.log; tail -3 gpurun_out/r02_dram_traffic_10M_stcs.csv | cut -c 180-; cat gpurun_out/r02_fp64_issue_mix.jsonl
