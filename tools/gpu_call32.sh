# the driver's multi-GPU command at N = 2 on the final state (data parallel + strong in one line)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/c32_bench2.json 2> gpurun_out/c32_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c32_bench2.json') if l.startswith('{')][-1])
print('N=2 dp', round(d['ms_per_step'],2), round(d['value']), 'strong', round(d['strong']['ms_per_step'],2), round(d['strong']['value']))
PY
