set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest_gpu.log
tail -3 gpurun_out/r2x_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seq_fwd2|seq_bwd|wgrad_tc|stft_fwd|cc_fwd" -c 10 -o gpurun_out/r2x_full python tools/run_once.py 256 2 > gpurun_out/r2x_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r2x_full.ncu-rep profiles/r2_ncu_summary > gpurun_out/r2x_ncu_summary.log 2>&1; cp profiles/r2_ncu_summary.json gpurun_out/r2x_ncu_summary.json; cp profiles/r2_ncu_summary.txt gpurun_out/r2x_ncu_summary.txt; cat profiles/r2_ncu_summary.txt
python tools/ncu_lines.py gpurun_out/r2x_full.ncu-rep seq_fwd2 > gpurun_out/r2x_seq_fwd2_hot_lines.txt 2>&1; head -5 gpurun_out/r2x_seq_fwd2_hot_lines.txt
rm -f gpurun_out/r2x_full.ncu-rep
timeout 600 python bench.py > gpurun_out/r2x_bench_final.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
head -c 260 gpurun_out/r2x_bench_final.json; echo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; tail -1 gpurun_out/r2x_smoke.log
