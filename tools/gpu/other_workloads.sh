# bench lines of the other BASELINE configurations + the error budget (gpurun -- 'bash tools/gpu/other_workloads.sh')
for w in c1 c3 c4 c5 c2_samdec; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --scene-runs 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || tail -3 gpurun_out/bench_$w.err
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_$w.json'))
print('$w', d['config']['workload'], 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['path_tensor_frac'], 3), 'clk', d['clocks']['sm_mhz'])
for k, v in list(d['kernels'].items())[:6]:
    print(f"    {k:20s} {v['ms_per_step']:.3f} ms  {v['tflops']:.0f} TF/s  {v['gbs']:.0f} GB/s")
for tag, s in (d.get('e2e_scene') or {}).items():
    for tie in ('numpy', 'stable'):
        print('   ', tag, tie, round(s[tie]['value'], 1), 'tiles/s', round(s[tie]['ms_per_scene'], 2), 'ms', s[tie]['stages_ms'], s[tie]['graph_stats'].get('us'), 'pts', s[tie]['n_points'])
PY
done
timeout 600 python tools/error_budget.py --patch 512 --gain 12 > gpurun_out/error_budget_512.json 2> gpurun_out/error_budget.err; tail -2 gpurun_out/error_budget.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/error_budget_512.json'))
for g, v in d.items():
    print(g, 'feat absmax', round(v['feat_absmax'], 2), 'mask range', [round(x, 2) for x in v['mask_logit_range']], 'topo range', [round(x, 2) for x in v['topo_logit_range']])
    for k, r in v['maxabs_vs_fp32_oracle'].items():
        print(f"   {k:60s} feat {r['feat']:.2e}  mask {r['mask_logit']:.2e}  topo {r['topo_logit_valid']:.2e}")
PY
