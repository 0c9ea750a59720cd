set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2p_bench_2gpu.json 2> gpurun_out/r2p_bench_2gpu.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_2gpu.json').read().strip().splitlines()[-1])
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'])
print(d.get('full_step'))
PY
tail -3 gpurun_out/r2p_bench_2gpu.err
