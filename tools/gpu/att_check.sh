timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 100 -k "gemm" 2>&1 | tail -3
timeout 100 python tools/gemm_probe.py lin1 qkv k3072 lin2
timeout 150 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 120 -k "vitb_512-512 or vitb_256-256" 2>&1 | tail -1
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v15.json 2> gpurun_out/bench_v15.err; tail -2 gpurun_out/bench_v15.err
python -c "
import json; d=json.load(open('gpurun_out/bench_v15.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for k,v in list(d['kernels'].items())[:8]: print(k, round(v['ms_per_step'],3), round(v['tflops'],1))
"
