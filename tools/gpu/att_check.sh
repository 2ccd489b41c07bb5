timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 100 -k "gemm_f32" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 150 -k "vitb_512-512 or vitb_256-256" 2>&1 | tail -1
for mode in 0 32 0 32; do
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --debug-gemm-mode $mode > gpurun_out/bench_ab_$mode.json 2> gpurun_out/bench_ab.err; tail -1 gpurun_out/bench_ab.err
python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$mode.json'))
k=d['kernels']
print('mode $mode', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), ' '.join(f\"{n}={k[n]['ms_per_step']:.3f}\" for n in ('layernorm','gemm_proj','gemm_mlp_lin2','gemm_qkv','gemm_mlp_lin1')))
"
done
