# 8 GPUs: NCCL parity of the row-sharded step, the default bench line (data parallel + strong), one chunk-count variant
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29520 tools/check_rowshard_nccl.py collab 64 > gpurun_out/c25_check8.log 2>&1; echo "check8 rc=$?"; grep -v "^\*\|OMP\|Warn" gpurun_out/c25_check8.log | tail -12
timeout 400 $TR --master-port 29521 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/c25_bench8.json 2> gpurun_out/c25_bench8.err; echo "bench8 rc=$?"
TWOWL_ROWSHARD_CHUNKS=2 timeout 300 $TR --master-port 29522 bench.py --gpus 8 --shard rows --steps 8 --warmup 3 > gpurun_out/c25_rows8_c2.json 2> gpurun_out/c25_rows8_c2.err; echo "rows c2 rc=$?"
python - <<'PY'
import json
def load(f):
    txt=open(f).read(); return json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
try:
    d=load('gpurun_out/c25_bench8.json'); print('N=8 dp', round(d['ms_per_step'],2), round(d['value']), 'strong', d['strong']['ms_per_step'], round(d['strong']['value']), d['strong']['per_op_ms_per_step_rank0'], d['strong']['collectives'])
except Exception as e: print('ERR', e)
for f in ('c25_rows8_c2',):
    try:
        d=load('gpurun_out/%s.json'%f); po=d['roofline']['per_op']
        print(f, round(d['ms_per_step'],2), round(d['value']), {k:round(v['ms']/d['steps'],2) for k,v in po.items()})
    except Exception as e: print(f,'ERR',e)
PY
