out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -m gpu -q > $out/pytest_r2f.log 2>&1; tail -3 $out/pytest_r2f.log
timeout 300 python tools/diag_batch.py 20 > $out/diag_batch_r2f.log 2>&1; cat $out/diag_batch_r2f.log
timeout 300 python tools/diag_c4.py 8 > $out/diag_c4_r2f.log 2>&1; tail -6 $out/diag_c4_r2f.log
B200REG_GICP_TRACE=1 timeout 300 python tools/profile_gicp.py 2 > $out/gicp_trace_r2f.log 2>&1; tail -8 $out/gicp_trace_r2f.log
