timeout 600 python -m pytest tests -q -m gpu --timeout 200 2>&1 | tail -3
timeout 200 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'])
for k,v in list(d['kernels'].items())[:10]: print(k, round(v['ms_per_step'],3), round(v['tflops'],1))
"
