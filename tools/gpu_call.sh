cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dual or fused or graph or nonfinite or multi_tile" 2>&1 | tail -2
for i in 1 2; do
BIEAR_B200_LIB=$PWD/biear_b200/lib/libbiear_b200_prev.so timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 300 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prev:', d['ms_per_step'], d['roofline']['us_per_launch'])"
timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 300 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new (snake + philox hoist): ', d['ms_per_step'], d['roofline']['us_per_launch'])"
done
