cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2zz_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2zz_pytest_gpu.log
tail -4 gpurun_out/r2zz_pytest_gpu.log; echo "t=$(( $(date +%s) - S ))"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2zz_smoke.log; tail -2 gpurun_out/r2zz_smoke.log; echo "t=$(( $(date +%s) - S ))"
timeout 240 python bench.py > gpurun_out/r2zz_bench_final.json 2> gpurun_out/r2zz_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2zz_bench_final.json; echo; echo "t=$(( $(date +%s) - S ))"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gru_|heads_|seq_|wgrad|stft|cc_fwd|prepare|q_reg" --csv --log-file gpurun_out/r2zz_full_step_ours.csv python tools/train_step_bench.py 256 2 > gpurun_out/r2zz_ncu_full.log 2>&1
python tools/launch_summary.py gpurun_out/r2zz_full_step_ours.csv 20 > gpurun_out/r2zz_full_step_ours.summary.txt 2>&1; head -16 gpurun_out/r2zz_full_step_ours.summary.txt; echo "t=$(( $(date +%s) - S ))"
