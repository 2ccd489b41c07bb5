# attention kernel A/B: op-level + model parity tests, then c2 / c1 bench lines (gpurun -- 'bash tools/gpu/att_ab.sh')
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/att_tests.log; tail -8 gpurun_out/att_tests.log
for w in c2 c1; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-scene > gpurun_out/att_ab_$w.json 2> gpurun_out/att_ab_$w.err || tail -3 gpurun_out/att_ab_$w.err
  python - <<PY
import json
d = json.load(open('gpurun_out/att_ab_$w.json'))
print('$w', 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'clk', d['clocks']['sm_mhz'])
for k, v in d['kernels'].items():
    if 'att' in k or 'topo_gemm' in k: print(f"    {k:20s} {v['ms_per_step']:.3f} ms  {v['tflops']:.0f} TF/s")
PY
done
# (round 2: compact Q tiles for the last window column -- a narrower TMA box for windows with only edge_w real
#  token columns -- measured with this script: window attention 2.22 -> 2.30 ms at c2, 0.869 -> 0.926 ms at c1,
#  i.e. slower; window units are bound by their latency chain, not by the number of active softmax warps. Reverted.)
