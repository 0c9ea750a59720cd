set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_gpu.log
tail -8 gpurun_out/r2m_pytest_gpu.log
