python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/c9_pytest.log | tail -12
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c9_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c9_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/c9_bench.json 2>gpurun_out/c9_b.err; echo "bench rc=$?"
TWOWL_PAIR_CONV_DUAL=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_dual.json 2>gpurun_out/c9_bd.err; echo "dual rc=$?"
python bench.py --hidden 128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_h128.json 2>gpurun_out/c9_b128.err; echo "h128 rc=$?"; tail -3 gpurun_out/c9_b128.err
python - <<'PY'
import json
def load(f):
    txt=open(f).read(); return json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
for f in ('c9_bench','c9_bench_dual','c9_bench_h128'):
    try:
        d=load('gpurun_out/%s.json'%f); po=d['roofline']['per_op']
        print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), {k:(round(v['ms']/d['steps'],2), v['GBps']) for k,v in list(po.items())[:6]})
    except Exception as e: print(f,'ERR',e)
PY
