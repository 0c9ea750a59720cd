cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_single.py tests/test_gpu_parity.py -k single -x -q 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"seq1|seq_bwd" --csv --log-file gpurun_out/r2y_single_launches.csv python tools/time_single.py 256 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2y_single_launches.csv 6
