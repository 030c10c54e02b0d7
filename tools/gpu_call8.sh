export TWOWL_PARITY_REPORT_ONLY=1
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/c8_pytest.log | tail -12
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/c8_bench1.json 2>gpurun_out/c8_b1.err; echo "bench1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/c8_bench2.json 2> gpurun_out/c8_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
def load(f):
    txt=open(f).read(); return json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
d=load('gpurun_out/c8_bench1.json'); po=d['roofline']['per_op']
print('N=1', round(d['ms_per_step'],2), {k:round(v['ms']/8,2) for k,v in po.items()})
d=load('gpurun_out/c8_bench2.json')
print('N=2 dp', round(d['ms_per_step'],2), 'strong', d['strong']['ms_per_step'], d['strong']['per_op_ms_per_step_rank0'])
PY
