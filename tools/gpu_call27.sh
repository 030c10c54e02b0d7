# final state: GPU tests, smoke, the default bench line (with the CPU legs), the reference arm
python -m pytest tests -m gpu -q > gpurun_out/c27_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c27_pytest.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c27_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c27_smoke.log
python bench.py > gpurun_out/c27_bench1.json 2> gpurun_out/c27_bench1.err; echo "bench1 rc=$?"; tail -c 1500 gpurun_out/c27_bench1.json
cp gpurun_out/parity_report.json gpurun_out/c27_parity_report.json
