python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/c23_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c23_pytest.log | tail -5
timeout 600 python tools/diag_node.py rmat 64 > gpurun_out/c23_node.log 2>&1; echo "node rc=$?"; grep -v Warn gpurun_out/c23_node.log | tail -9
DIAG_PATHS=1 timeout 600 python tools/diag_fullsize.py rmat 64 > gpurun_out/c23_diag.log 2>&1; echo "diag rc=$?"; grep -v Warning gpurun_out/c23_diag.log | tail -7
