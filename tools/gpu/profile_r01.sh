set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_launches_v4.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 26 -c 4 -o gpurun_out/r01_gemm_f16_v4 $B > gpurun_out/ncu_g1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_resid -s 24 -c 2 -o gpurun_out/r01_gemm_resid_v4 $B > gpurun_out/ncu_g2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 13 -c 2 -o gpurun_out/r01_att_v6 $B > gpurun_out/ncu_a.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_launches_v4.csv
