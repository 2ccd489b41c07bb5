# N-GPU checks (gpurun --gpus N -- 'bash tools/gpu/multi_gpu_check.sh N'): NCCL scene test + the weak-scaling bench line
N=${1:-2}
timeout 420 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -5
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -E "bench mem|NCCL INFO (Connected|Channel 00/|comm .* rank)|NVLS" gpurun_out/bench_n$N.err | head -24
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").readline())
print('N', d['n_gpus'], 'value', round(d['value'], 1), 'per gpu', round(d['value'] / d['n_gpus'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
print('clocks', d['clocks'])
for tag, s in (d.get('e2e_scene') or {}).items():
    for tie in ('numpy', 'stable'):
        print(tag, tie, round(s[tie]['value'], 1), 'tiles/s', round(s[tie]['ms_per_scene'], 2), 'ms', s[tie]['stages_ms'])
PY
# A/B: the same step without the exchange (what the slowest GPU alone costs)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-exchange --no-scene --no-cpu-baseline > gpurun_out/bench_n${N}_noex.json 2> gpurun_out/bench_n${N}_noex.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n${N}_noex.json').readline())
print('NO-EXCHANGE N', d['n_gpus'], 'value', round(d['value'], 1), 'per gpu', round(d['value'] / d['n_gpus'], 1), 'ms/step', round(d['ms_per_step'], 3), 'by rank', d['ms_per_step_by_rank'])
d = json.loads(open('gpurun_out/bench_n$N.json').readline())
print('WITH EXCHANGE by rank', d['ms_per_step_by_rank'], d['config']['exchange'])
PY
