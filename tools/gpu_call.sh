set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transparent or graphed or single" > gpurun_out/r2e_pytest_graph.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest_graph.log
tail -4 gpurun_out/r2e_pytest_graph.log
timeout 900 python tools/run_reference_pipeline.py --clips 6144 --batch 256 --log-dir gpurun_out > gpurun_out/r2e_pipeline.log 2>&1
tail -3 gpurun_out/r2e_pipeline.log
grep "biear_b200 timing" gpurun_out/r2_train_biear_unchanged.log gpurun_out/r2_evaluate_biear_unchanged.log
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 100 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
head -c 300 gpurun_out/r2e_bench.json; echo; tail -3 gpurun_out/r2e_bench.err
