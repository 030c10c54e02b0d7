python tools/diag_parity.py > gpurun_out/c3_diag.log 2>&1; cat gpurun_out/c3_diag.log
