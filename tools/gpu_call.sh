cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BIEAR_B200_LIB=$PWD/biear_b200/lib/libbiear_b200_prev.so timeout 600 ncu --set full --clock-control none -k regex:seq_fwd2 -c 2 -o gpurun_out/r2t_prev python tools/run_once.py 256 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:seq_fwd2 -c 2 -o gpurun_out/r2t_new python tools/run_once.py 256 2 > /dev/null 2>&1
for v in prev new; do
python tools/ncu_summary.py gpurun_out/r2t_$v.ncu-rep gpurun_out/r2t_ncu_$v > /dev/null 2>&1; cat gpurun_out/r2t_ncu_$v.txt | head -12
done
rm -f gpurun_out/r2t_prev.ncu-rep
