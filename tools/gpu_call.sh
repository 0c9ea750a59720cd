cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for flag in "" "--graph-allreduce"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 300 --warmup 10 --no-extra $flag 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$flag', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['allreduce'])"
done
