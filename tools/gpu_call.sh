set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_heads.py -x -q > gpurun_out/r2n_pytest_heads.log 2>&1; tail -3 gpurun_out/r2n_pytest_heads.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:heads --csv --log-file gpurun_out/r2n_heads_launches.csv python -m pytest tests/test_gpu_heads.py -x -q -k "256" > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2n_heads_launches.csv 10
timeout 600 python tools/train_step_bench.py 256 30 2>&1 | tail -1
