# Profiles of the final kernels (one GPU): launch list, DRAM bytes per launch, --set full of the three tensor-core kernels.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 1 --no-graph --no-shooting-leg --cpu-seconds 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2f_launches.csv $B > /dev/null 2> gpurun_out/r2f_launches.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"mlp_fwd_kernel|mlp_bwd_chain_kernel|mlp_bwd_input_kernel" -s 200 -c 97 --csv --log-file gpurun_out/r2f_dram_bytes.csv $B > /dev/null 2> gpurun_out/r2f_dram.err
ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_kernel -s 104 -c 2 -f -o gpurun_out/r2f_fwd $B > /dev/null 2> gpurun_out/r2f_ncu_fwd.err
ncu --set full --clock-control none --import-source on -k regex:"mlp_bwd_chain_kernel|mlp_bwd_input_kernel" -s 70 -c 4 -f -o gpurun_out/r2f_bwd $B > /dev/null 2> gpurun_out/r2f_ncu_bwd.err
ls -la gpurun_out/r2f_*
