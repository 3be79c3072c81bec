# the round-end single-GPU measurements: GPU tests, recorder timings, the driver's default bench line, the reference arm
set -x
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_t_final.log 2>&1; tail -3 gpurun_out/r2_t_final.log
python tools/bench_recorders.py > gpurun_out/r02_recorders.jsonl 2>&1; cut -c1-160 gpurun_out/r02_recorders.jsonl | sed "s/\"accepted\".*//"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_1gpu_10M_final.json 2> gpurun_out/r02_bench_1gpu_final.err; cut -c1-300 gpurun_out/r02_bench_1gpu_10M_final.json
