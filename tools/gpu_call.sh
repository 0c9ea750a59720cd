set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_single.py tests/test_gpu_parity.py -k "single" -x -q -s > gpurun_out/r2k_pytest_single.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_single.log
tail -25 gpurun_out/r2k_pytest_single.log
timeout 600 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
print(json.dumps(d.get('variants'), indent=1))
"
tail -3 gpurun_out/r2k_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2k_fullstep_launches.csv python tools/train_step_bench.py 256 3 > gpurun_out/r2k_fullstep_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r2k_fullstep_launches.csv 60 > gpurun_out/r2k_fullstep_launches.summary.txt 2>&1
head -70 gpurun_out/r2k_fullstep_launches.summary.txt
