set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_gpu.log
tail -5 gpurun_out/r2i_pytest_gpu.log
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 200 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
head -c 300 gpurun_out/r2i_bench.json; echo; tail -3 gpurun_out/r2i_bench.err
timeout 300 python tools/phase_prof.py 256 > gpurun_out/r2i_phase_cycles.txt 2>&1
cat gpurun_out/r2i_phase_cycles.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; tail -2 gpurun_out/r2i_smoke.log
