# 2 GPUs: A/B of the NCCL stream priority and the number of node ranges on the row-sharded step
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" timeout 300 $TR --master-port $PORT bench.py --gpus 2 --shard rows --steps 8 --warmup 3 > gpurun_out/c26_$name.json 2> gpurun_out/c26_$name.err; echo "$name rc=$?"; PORT=$((PORT+1)); }
PORT=29530
run c2_prio0 TWOWL_ROWSHARD_CHUNKS=2 TWOWL_NCCL_HIGH_PRIO=0
run c2_prio1 TWOWL_ROWSHARD_CHUNKS=2 TWOWL_NCCL_HIGH_PRIO=1
run c4_prio1 TWOWL_ROWSHARD_CHUNKS=4 TWOWL_NCCL_HIGH_PRIO=1
python - <<'PY'
import json
def load(f):
    txt=open(f).read(); return json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
for f in ('c2_prio0','c2_prio1','c4_prio1'):
    try:
        d=load('gpurun_out/c26_%s.json'%f); po=d['roofline']['per_op']
        print(f, round(d['ms_per_step'],2), round(d['value']), {k:round(v['ms']/d['steps'],2) for k,v in po.items() if v['ms']/d['steps']>=0.2})
    except Exception as e: print(f,'ERR',e)
PY
