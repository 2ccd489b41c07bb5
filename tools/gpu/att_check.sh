timeout 500 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 150 2>&1 | tail -2
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v25.json 2> gpurun_out/bench_v25.err; tail -2 gpurun_out/bench_v25.err
python -c "
import json; d=json.load(open('gpurun_out/bench_v25.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items())[:14]: print(k, round(v['ms_per_step'],3), round(v['tflops'],1))
"
