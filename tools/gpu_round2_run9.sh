set -x
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/host_bw_probe.py > gpurun_out/r02_hostbw_${N}gpu_b.json 2>&1
for K in 1 2 4; do
DEB_COPY_STREAMS=$K DEB_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$K bench.py --gpus $N --e2e-only > gpurun_out/r02_bench_${N}gpu_e2e_cs$K.json 2> gpurun_out/r02_bench_${N}gpu_e2e_cs$K.err
echo "== copy streams $K"; grep -o '"e2e": {"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/r02_bench_${N}gpu_e2e_cs$K.json | tail -3
grep "deb timing" gpurun_out/r02_bench_${N}gpu_e2e_cs$K.err | tail -8 | sed 's/.*first block done/first block done/; s/; kernel.*//'
done
python - <<PY
import json
for l in open('gpurun_out/r02_hostbw_${N}gpu_b.json'):
    if l.startswith('{'):
        d=json.loads(l)
        for k,v in d['results'].items(): print(k, [round(x,1) for x in v['per_rank_gbs']], round(v['aggregate_gbs_by_wall'],1))
PY
