# round 1, capture l: launch lists (time + DRAM bytes) of the next-row kernels and the general path.
set -x
G="python benchmarks/configs.py general"
$G > gpurun_out/general_plain_r1l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gen_ -c 600 --csv --log-file gpurun_out/launches_general_r1l.csv $G > /dev/null 2>&1
python benchmarks/next_rows.py > gpurun_out/next_rows_plain_r1l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"classical_chain_kernel|trajectory_kernel|state_moments_kernel|kl_constants" --csv --log-file gpurun_out/launches_next_rows_r1l.csv python benchmarks/next_rows.py > /dev/null 2>&1
tail -2 gpurun_out/next_rows_plain_r1l.log
