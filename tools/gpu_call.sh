set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest_gpu.log
tail -4 gpurun_out/r2v_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2v_bench_final.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
head -c 300 gpurun_out/r2v_bench_final.json; echo; tail -2 gpurun_out/r2v_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; tail -1 gpurun_out/r2v_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2v_launches_final.csv python bench.py --eager --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2v_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r2v_launches_final.csv 20 > gpurun_out/r2v_launches_final.summary.txt; head -14 gpurun_out/r2v_launches_final.summary.txt
