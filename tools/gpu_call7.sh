python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/c7_bench8.json 2> gpurun_out/c7_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
txt=open('gpurun_out/c7_bench8.json').read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('N=8 dp', round(d['ms_per_step'],2), d['value'], 'strong', d['strong']['ms_per_step'], d['strong']['value'])
print(d['strong']['per_op_ms_per_step_rank0'])
print(d['strong']['collectives'])
PY
