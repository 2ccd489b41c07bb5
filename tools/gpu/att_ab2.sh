# attention epilogue A/B (transposed stores): the whole GPU suite, then short c2 / c1 bench lines with per-kernel times
timeout 1700 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/tests.log; tail -4 gpurun_out/tests.log
for w in c2 c1; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-scene > gpurun_out/att_ab_$w.json 2> gpurun_out/att_ab_$w.err || tail -3 gpurun_out/att_ab_$w.err
  python - <<PY
import json
d = json.load(open('gpurun_out/att_ab_$w.json'))
print('$w', 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'clk', d['clocks']['sm_mhz'])
for k, v in d['kernels'].items():
    if 'att' in k or 'lin1' in k: print(f"    {k:20s} {v['ms_per_step']:.3f} ms  {v['tflops']:.0f} TF/s")
PY
done
