# re-entry check of HEAD: GPU tests, smoke, the default bench line, the per-op list of one step
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/c17_pytest.log | tail -6
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c17_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c17_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/c17_bench1.json 2> gpurun_out/c17_bench1.err; echo "bench1 rc=$?"; tail -c 600 gpurun_out/c17_bench1.json
