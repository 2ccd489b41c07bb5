N=${1:-8}
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -E "NCCL INFO (Connected all|comm .* rank .* nRanks)" gpurun_out/bench_n$N.err | head -10
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$N.json').readline())
print('N', d['n_gpus'], 'value', round(d['value'], 1), 'per gpu', round(d['value'] / d['n_gpus'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), d['config']['exchange'])
print('by rank', d['ms_per_step_by_rank'], 'clocks', d['clocks']['sm_mhz'])
for tag, s in (d.get('e2e_scene') or {}).items():
    for tie in ('numpy', 'stable'):
        print(tag, tie, round(s[tie]['value'], 1), 'tiles/s', round(s[tie]['ms_per_scene'], 2), 'ms', s[tie]['stages_ms'])
PY
