# round 2, call o: L2 random-gather microbenchmark (load flavours for the global-table walk)
mkdir -p gpurun_out
timeout 120 profiles/microbench/_build/gather_l2 > gpurun_out/r02_o_gather_l2.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_o_gather_l2.txt
cat gpurun_out/r02_o_gather_l2.txt
