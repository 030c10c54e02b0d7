python -m pytest tests/test_gpu_rowshard.py -m gpu -q -x > gpurun_out/c33_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/c33_pytest.log | tail -4
timeout 200 python tools/rowshard_dry.py 8 64 0 > gpurun_out/c33_dry.log 2>&1; echo "dry rc=$?"; grep "^rank\|^   [a-z]" gpurun_out/c33_dry.log
