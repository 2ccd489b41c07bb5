for mode in "" "--no-exchange" ""; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-scene --no-cpu-baseline $mode > gpurun_out/n2ab.json 2> gpurun_out/n2ab.err
python - <<PY
import json
d = json.loads(open('gpurun_out/n2ab.json').readline())
print('mode [$mode]', 'ms/step', round(d['ms_per_step'], 3), 'by rank', d['ms_per_step_by_rank'], d['config']['exchange'], 'clk', d['clocks']['sm_mhz'])
PY
done
