cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gru.py tests/test_gpu_heads.py tests/test_gpu_model.py -x -q > gpurun_out/r2z4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z4_pytest.log
grep -v "^$" gpurun_out/r2z4_pytest.log | tail -8
timeout 120 python tools/train_step_bench.py 256 30 > gpurun_out/r2z4_full_native.log 2>&1; tail -1 gpurun_out/r2z4_full_native.log
timeout 60 python tools/time_gru.py 256 2>&1 | grep -i "native" > gpurun_out/r2z4_time_gru.log; cat gpurun_out/r2z4_time_gru.log
