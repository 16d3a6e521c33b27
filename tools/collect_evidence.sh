#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): GPU tests, the bench line, the reference arm, the ncu launch list of the
# bench command and `--set full` captures of the dominant GEMM, the Wavenet two-pass GEMM, the streaming RMSNorm and the
# alignment kernel.  Everything lands in gpurun_out/.
set -u
R=${1:-r02k}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/${R}_pytest_gpu.log
tail -3 $O/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
timeout 900 python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
tail -c 300 $O/${R}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2> $O/${R}_bench_reference.err
tail -c 300 $O/${R}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${R}_launch_list_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/${R}_launch_list_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 1 -c 1 -o $O/${R}_prof_conv \
    python tools/prof_kernels.py conv > $O/${R}_prof_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2w_kernel -s 1 -c 1 -o $O/${R}_prof_wavenet \
    python tools/prof_kernels.py wavenet > $O/${R}_prof_wavenet.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rmsnorm_stream -s 1 -c 1 -o $O/${R}_prof_norm \
    python tools/prof_kernels.py norm > $O/${R}_prof_norm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mas_ -s 2 -c 2 -o $O/${R}_prof_mas \
    python tools/aligner_bench.py > $O/${R}_prof_mas.log 2>&1
ls -la $O | tail -20
