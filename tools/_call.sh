SPN_TRACE_BLOCK=1 timeout 120 python tools/trace_fwd_ts.py 1048576 train > gpurun_out/r02d_trace_ts_train_peer.txt 2>&1; tail -1 gpurun_out/r02d_trace_ts_train_peer.txt
