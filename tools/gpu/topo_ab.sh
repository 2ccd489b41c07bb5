# TopoNet kernel A/B: op-level + model parity tests, then the c2 / c4 bench lines (gpurun -- 'bash tools/gpu/topo_ab.sh')
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu 2>&1 | tail -15 > gpurun_out/topo_tests.log; tail -8 gpurun_out/topo_tests.log
for w in c2 c4; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-scene > gpurun_out/topo_ab_$w.json 2> gpurun_out/topo_ab_$w.err || tail -3 gpurun_out/topo_ab_$w.err
  python - <<PY
import json
d = json.load(open('gpurun_out/topo_ab_$w.json'))
print('$w', 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'clk', d['clocks']['sm_mhz'])
for k, v in d['kernels'].items():
    if 'topo' in k: print(f"    {k:20s} {v['ms_per_step']:.3f} ms  {v['tflops']:.0f} TF/s")
PY
done
