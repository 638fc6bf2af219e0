ncu --set full --clock-control none --import-source on -k regex:gate_kernel -c 12 -o gpurun_out/prof_gate4 -f python tools/kernel_bench.py --iters 1 --only gate > gpurun_out/ncu_gate4.log 2>&1
ls -la gpurun_out/
