mkdir -p gpurun_out
NPSL_TRACE_KEEP=gpurun_out/s14_trace8.npy python tools/trace_timeline.py S64_disc 8 0 > gpurun_out/s14_t8.txt 2>&1
NPSL_TRACE_KEEP=gpurun_out/s14_trace1.npy python tools/trace_timeline.py S64_disc 1 0 > gpurun_out/s14_t1.txt 2>&1
