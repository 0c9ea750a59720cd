cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r2z7_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z7_pytest_gpu.log
tail -3 gpurun_out/r2z7_pytest_gpu.log
timeout 60 python tools/time_gru.py 256 2>&1 | grep -i "pack" > gpurun_out/r2z7_time_gru.log; cat gpurun_out/r2z7_time_gru.log
timeout 100 python tools/train_step_bench.py 256 30 > gpurun_out/r2z7_full_native.log 2>&1; tail -1 gpurun_out/r2z7_full_native.log
