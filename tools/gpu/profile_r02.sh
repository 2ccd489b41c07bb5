# ncu captures behind profiles/r02_* and profiles/ncu_metrics.json (run on the B200 box: gpurun -- 'bash tools/gpu/profile_r02.sh')
set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scene"
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $B > $O/ncu_l.log 2>&1
# per step: gemm_tc2_kernel = [qkv, lin1] x 12 + decoder stage 2; gemm_tc2_resid = patch-embed + [proj, lin2] x 12; attention = w w g w w g ...
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 25 -c 4 -o $O/r02_gemm_f16 $B > $O/ncu_g1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_resid -s 26 -c 4 -o $O/r02_gemm_resid $B > $O/ncu_g2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 13 -c 2 -o $O/r02_att $B > $O/ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"layernorm_f16_kernel|toponet_tc_kernel" -s 30 -c 3 -o $O/r02_tail $B > $O/ncu_t.log 2>&1
python tools/ncu_extract.py $O/ncu_metrics.json c2 \
  $O/r02_gemm_f16.ncu-rep:gemm_qkv,gemm_mlp_lin1,gemm_qkv,gemm_mlp_lin1 \
  $O/r02_gemm_resid.ncu-rep:gemm_proj,gemm_mlp_lin2,gemm_proj,gemm_mlp_lin2 \
  $O/r02_att.ncu-rep:attention_window,attention_global > $O/ncu_extract.log 2>&1
for r in r02_gemm_f16 r02_gemm_resid r02_att r02_tail; do python tools/ncu_top.py $O/$r.ncu-rep 14 > $O/${r}_top.txt 2>&1; done
ls -la $O/*.ncu-rep
