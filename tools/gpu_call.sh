cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/ubench/mma_rate > gpurun_out/r2q_mma_rate.txt 2>&1; cat gpurun_out/r2q_mma_rate.txt
