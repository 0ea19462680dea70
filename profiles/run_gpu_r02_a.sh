# round 2, call a: shared-memory ranking microbenchmark + host facts of the GPU box
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_a_host.txt; nproc >> gpurun_out/r02_a_host.txt; free -g >> gpurun_out/r02_a_host.txt
timeout 120 profiles/microbench/_build/smem_rank > gpurun_out/r02_a_smem_rank.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_a_smem_rank.txt
cat gpurun_out/r02_a_smem_rank.txt gpurun_out/r02_a_host.txt
