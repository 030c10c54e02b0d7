TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29520 tools/check_rowshard_nccl.py collab 64 > gpurun_out/c16_check8.log 2>&1; echo "check8 rc=$?"; grep -v "^\*\|OMP" gpurun_out/c16_check8.log | head -30
$TR --master-port 29521 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/c16_bench8.json 2> gpurun_out/c16_bench8.err; echo "bench8 rc=$?"
TWOWL_ROWSHARD_CHUNKS=1 $TR --master-port 29522 bench.py --gpus 8 --shard rows --steps 8 --warmup 3 > gpurun_out/c16_rows8_c1.json 2> gpurun_out/c16_rows8_c1.err; echo "rows c1 rc=$?"
$TR --master-port 29523 bench.py --gpus 8 --shard rows --hidden 256 --steps 4 --warmup 3 > gpurun_out/c16_rows8_h256.json 2> gpurun_out/c16_rows8_h256.err; echo "h256 rc=$?"; tail -3 gpurun_out/c16_rows8_h256.err
python - <<'PY'
import json
def load(f):
    txt=open(f).read(); return json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
try:
    d=load('gpurun_out/c16_bench8.json'); print('N=8 dp', round(d['ms_per_step'],2), round(d['value']), 'strong', d['strong']['ms_per_step'], round(d['strong']['value']), d['strong']['per_op_ms_per_step_rank0'])
except Exception as e: print('ERR', e)
for f in ('c16_rows8_c1','c16_rows8_h256'):
    try:
        d=load('gpurun_out/%s.json'%f); po=d['roofline']['per_op']
        print(f, round(d['ms_per_step'],2), round(d['value']), {k:round(v['ms']/d['steps'],2) for k,v in po.items()})
    except Exception as e: print(f,'ERR',e)
PY
