python -m pytest tests -m gpu -q -x > gpurun_out/c34_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c34_pytest.log | tail -3
