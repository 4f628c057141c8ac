python meshgraphnets.jl_b200/build.py --trace -f > /dev/null 2>&1
python tools/trace_kernels.py > gpurun_out/r2_trace.txt 2> gpurun_out/r2_trace.err
tail -3 gpurun_out/r2_trace.err; wc -l gpurun_out/r2_trace.txt
