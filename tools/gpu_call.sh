cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_model.py tests/test_gpu_gru.py -q -k "auralnet or graph or aural" > gpurun_out/r2z5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z5_pytest.log
grep -v "^$" gpurun_out/r2z5_pytest.log | tail -30
