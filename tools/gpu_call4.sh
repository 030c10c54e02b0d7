export TWOWL_PARITY_REPORT_ONLY=1
python -m pytest tests -m gpu -q --durations=8 -x > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c4_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"
grep -E "passed|failed|rc=|Error" gpurun_out/c4_pytest.log | tail -8; tail -3 gpurun_out/c4_smoke.log; tail -5 gpurun_out/c4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c4_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'])
print({k:(v['ms'],v['GBps']) for k,v in d['roofline']['per_op'].items()})
print(json.dumps(d['same_config'])[:1500])
PY
