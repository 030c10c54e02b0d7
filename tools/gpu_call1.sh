export TWOWL_PARITY_REPORT_ONLY=1
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
TWOWL_LINEAR_IMPL=0 TWOWL_PARITY_REPORT=gpurun_out/parity_report_simt.json python -m pytest tests/test_gpu_model.py -m gpu -q > gpurun_out/c1_pytest_simt.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c1_smoke.log
tail -5 gpurun_out/c1_pytest.log; tail -3 gpurun_out/c1_smoke.log
