# whole GPU suite + the default bench line (run on the B200 box: gpurun -- 'bash tools/gpu/full_check.sh')
timeout 1700 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/tests.log; tail -15 gpurun_out/tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -5 gpurun_out/bench_c2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2.json'))
print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'])
print('clocks', d['clocks'])
for k, v in list(d['kernels'].items())[:14]:
    print(f"  {k:20s} {v['ms_per_step']:.3f} ms  {v['tflops']:.0f} TF/s  {v['gbs']:.0f} GB/s")
for tag, s in (d.get('e2e_scene') or {}).items():
    for tie in ('numpy', 'stable'):
        print(tag, tie, round(s[tie]['value'], 1), 'tiles/s', s[tie]['ms_per_scene'], 'ms', s[tie]['stages_ms'], s[tie]['graph_stats'], 'pts', s[tie]['n_points'], 'edges', s[tie]['n_edges'], 'samples', s[tie]['topo_samples'])
print('cpu', d['cpu_baseline'])
PY
