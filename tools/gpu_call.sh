cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gru.py -x -q -s -k "element_wise or encoders or graph" > gpurun_out/r2z3_gru_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z3_gru_pytest.log
grep -v "^$" gpurun_out/r2z3_gru_pytest.log | tail -12
timeout 90 python tools/time_gru.py 256 > gpurun_out/r2z3_time_gru.log 2>&1; tail -3 gpurun_out/r2z3_time_gru.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gru_(fwd|bwd)" -s 4 -c 2 -o gpurun_out/r2z3_gru -f python tools/time_gru.py 256 profile > gpurun_out/r2z3_ncu.log 2>&1; tail -3 gpurun_out/r2z3_ncu.log
