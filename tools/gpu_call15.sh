python -m pytest tests -m gpu -q > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c15_pytest.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/c15_pytest.log | tail -6
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c15_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c15_smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/c15_bench2.json 2> gpurun_out/c15_bench2.err; echo "bench2 rc=$?"
python bench.py --workload collab --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c15_collab1.json 2> gpurun_out/c15_collab1.err; echo "collab1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload collab --steps 10 --warmup 3 > gpurun_out/c15_collab2.json 2> gpurun_out/c15_collab2.err; echo "collab2 rc=$?"
python bench.py --workload cora --graph --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c15_cora_graph.json 2> gpurun_out/c15_cora.err; echo "cora rc=$?"
python - <<'PY'
import json
def load(f):
    txt=open(f).read(); return json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
for f in ('c15_bench2','c15_collab1','c15_collab2','c15_cora_graph'):
    try:
        d=load('gpurun_out/%s.json'%f)
        print(f, 'N', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],3), 'strong', (d.get('strong') or {}).get('ms_per_step'), (d.get('strong') or {}).get('per_op_ms_per_step_rank0'))
    except Exception as e: print(f,'ERR',e)
PY
