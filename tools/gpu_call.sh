cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 28 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_step_bench.py 256 10 > gpurun_out/r2z8_full_2gpu.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2z8_full_2gpu.log
