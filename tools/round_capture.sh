#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): GPU tests, smoke, the bench line, the ncu launch list of the timed
# regions, per-launch DRAM / tensor-pipe metrics of a step's conv launches, one --set full capture of the dominant launch.
# Every ncu pass runs only after the same command has exited 0 without ncu.
TAG=${1:-v6}
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err && python tools/print_bench.py gpurun_out/bench_$TAG.json
python bench.py --steps 2 --warmup 3 --legs none --no-cpu-baseline > /dev/null 2>&1 && \
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/one_step_eager.py 4 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:conv_igemm -s 195 -c 65 --csv --log-file gpurun_out/conv_step_$TAG.csv python tools/one_step_eager.py 4 > gpurun_out/ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 249 -c 1 -f -o /tmp/aspp_full python tools/one_step_eager.py 4 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/aspp_full.ncu-rep --page raw --csv > gpurun_out/aspp_full_raw_$TAG.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
