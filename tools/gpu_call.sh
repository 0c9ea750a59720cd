# Final-validation call of a round (run as: gpurun --timeout 300 -- 'bash tools/gpu_call.sh'): the whole GPU test suite, the
# smoke check, the default bench line and the launch list of this package's kernels inside the full training step.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log
tail -4 gpurun_out/final_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log; tail -2 gpurun_out/final_smoke.log
timeout 240 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gru_|heads_|seq_|wgrad|stft|cc_fwd|prepare|q_reg" --csv --log-file gpurun_out/final_full_step_ours.csv python tools/train_step_bench.py 256 2 > gpurun_out/final_ncu_full.log 2>&1
python tools/launch_summary.py gpurun_out/final_full_step_ours.csv 20 > gpurun_out/final_full_step_ours.summary.txt 2>&1; head -16 gpurun_out/final_full_step_ours.summary.txt
