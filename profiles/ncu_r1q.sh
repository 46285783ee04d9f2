# round 1, capture q: --set full of the shared-memory chain kernel (C2 shape) and of the general path's pass kernel.
set -x
python benchmarks/configs.py c2 > gpurun_out/c2_plain_r1q.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:smem_chain_kernel -s 1 -c 1 -o gpurun_out/small_r1q python benchmarks/configs.py c2 > /dev/null 2>&1
python profiles/summarize_ncu.py gpurun_out/small_r1q.ncu-rep gpurun_out/r1q_small_ncu_full.json
python benchmarks/configs.py general > gpurun_out/general_plain_r1q.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gen_pass_kernel -s 20 -c 2 -o gpurun_out/general_r1q python benchmarks/configs.py general > /dev/null 2>&1
python profiles/summarize_ncu.py gpurun_out/general_r1q.ncu-rep gpurun_out/r1q_general_ncu_full.json
rm -f gpurun_out/small_r1q.ncu-rep gpurun_out/general_r1q.ncu-rep
