set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/dev_single.py > gpurun_out/dev_single.log 2>&1; echo "rc=$?" >> gpurun_out/dev_single.log
tail -40 gpurun_out/dev_single.log
