set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"toponet_tc_kernel|layernorm_f16_kernel|EpiDecFinal|EpiLN" -s 30 -c 8 -o gpurun_out/r01_tail_v4 $B > gpurun_out/ncu_t.log 2>&1
tail -2 gpurun_out/ncu_t.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc80 -s 2 -c 2 -o gpurun_out/r01_att80_v1 python tools/bench_configs.py --steps 1 > gpurun_out/ncu_t2.log 2>&1
tail -2 gpurun_out/ncu_t2.log
ls -la gpurun_out/r01_tail_v4.ncu-rep gpurun_out/r01_att80_v1.ncu-rep
