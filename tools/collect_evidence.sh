#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): GPU tests, the bench line, the reference arm, the ncu launch list of the
# bench command and `--set full` captures of the dominant GEMM and the attention kernel.  Everything lands in gpurun_out/.
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/${R}_pytest_gpu.log
tail -3 $O/${R}_pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/${R}_clocks.csv &
SMI=$!
timeout 900 python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
kill $SMI
tail -c 600 $O/${R}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2> $O/${R}_bench_reference.err
tail -c 400 $O/${R}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${R}_launch_list_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/${R}_launch_list_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 1 -c 2 -o $O/${R}_prof_conv \
    python tools/prof_kernels.py conv > $O/${R}_prof_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_fwd -s 1 -c 2 -o $O/${R}_prof_attn \
    python tools/prof_kernels.py attn > $O/${R}_prof_attn.log 2>&1
ls -la $O
