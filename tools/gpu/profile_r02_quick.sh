# sanity subset + the ncu captures that feed ncu_metrics.json (no LayerNorm / TopoNet capture): for a late kernel change
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu -k "attention or vitb_512-512 or determinism" 2>&1 | tail -3
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scene"
O=gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $B > $O/ncu_l.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 13 -c 2 -o $O/r02_att $B > $O/ncu_a.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 25 -c 4 -o $O/r02_gemm_f16 $B > $O/ncu_g1.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_resid -s 26 -c 4 -o $O/r02_gemm_resid $B > $O/ncu_g2.log 2>&1
python tools/ncu_extract.py $O/ncu_metrics.json c2 \
  $O/r02_gemm_f16.ncu-rep:gemm_qkv,gemm_mlp_lin1,gemm_qkv,gemm_mlp_lin1 \
  $O/r02_gemm_resid.ncu-rep:gemm_proj,gemm_mlp_lin2,gemm_proj,gemm_mlp_lin2 \
  $O/r02_att.ncu-rep:attention_window,attention_global > $O/ncu_extract.log 2>&1
tail -6 $O/ncu_extract.log
