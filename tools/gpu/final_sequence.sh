bash tools/gpu/profile_r02_extra.sh 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err; tail -2 gpurun_out/bench_c2_final.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2_final.json'))
print('c2 value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'clk', d['clocks']['sm_mhz'], 'tensor%', d.get('tensor_pipe_pct_ncu'), 'roofline', d['roofline'])
PY
bash tools/gpu/other_workloads.sh
