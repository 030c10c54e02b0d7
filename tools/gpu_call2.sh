export TWOWL_PARITY_REPORT_ONLY=1
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c2_smoke.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
grep -E "passed|failed|rc=" gpurun_out/c2_pytest.log | tail -5; tail -3 gpurun_out/c2_smoke.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/c2_bench.json'))
print(d['ms_per_step'], d['value'], {k:(v['ms'],v['GBps']) for k,v in d['roofline']['per_op'].items()})
PY
