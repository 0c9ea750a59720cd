set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# 1. correctness of the st.async hand-over kernels first (bounded: a protocol bug traps)
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dual_adaptive or multi_tile or nonfinite or properties_full or streamed or graphed" > gpurun_out/r2b_pytest_seq.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest_seq.log
tail -4 gpurun_out/r2b_pytest_seq.log
# 2. timing
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 100 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
head -c 300 gpurun_out/r2b_bench.json; echo
# 3. phase profile with the band-stage / last-phase split
timeout 300 python tools/phase_prof.py 256 > gpurun_out/r2b_phase_cycles.txt 2>&1
cat gpurun_out/r2b_phase_cycles.txt
# 4. full parity
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
tail -5 gpurun_out/r2b_pytest_gpu.log
# 5. the element-wise table with the round-1 library (numerics of the old band stage) for comparison
BIEAR_B200_LIB=$PWD/biear_b200/lib/libbiear_b200_r1.so timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -s > gpurun_out/r2b_pytest_full_r1lib.log 2>&1
tail -3 gpurun_out/r2b_pytest_full_r1lib.log
