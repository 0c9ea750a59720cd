cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gru.py -x -q > gpurun_out/r2z_gru_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_gru_pytest.log
tail -25 gpurun_out/r2z_gru_pytest.log
timeout 90 python tools/time_gru.py 256 > gpurun_out/r2z_time_gru.log 2>&1; cat gpurun_out/r2z_time_gru.log | tail -12
timeout 120 python tools/train_step_bench.py 256 30 > gpurun_out/r2z_full_native.log 2>&1; tail -2 gpurun_out/r2z_full_native.log
timeout 120 python tools/train_step_bench.py 256 30 libgru > gpurun_out/r2z_full_libgru.log 2>&1; tail -2 gpurun_out/r2z_full_libgru.log
