#!/bin/bash
# One GPU-box pass at the end of a round: parity tests, smoke, the bench (all workloads + reference GPU baseline), the CPU arm,
# then the ncu evidence for profiles/ (launch lists of the same bench command + one --set full capture per dominant kernel).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_end.sh r02'
# Everything lands in gpurun_out/ (summaries are made afterwards with tools/ncu_summary.py).
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -W ignore > $O/${R}_pytest.log 2>&1; tail -2 $O/${R}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py 2> $O/${R}_bench.err > $O/${R}_bench.json; cut -c1-170 $O/${R}_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/${R}_bench_ref.json 2>/dev/null; cut -c1-170 $O/${R}_bench_ref.json
for w in rgb sam mask; do
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_$w.csv \
        python bench.py --workload $w --only --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:render_kernel --launch-skip 2 -c 1 -f \
    -o $O/${R}_rgb python bench.py --workload rgb --only --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name 'regex:render_kernel|samvit_mlp_kernel' --launch-skip 4 -c 2 -f \
    -o $O/${R}_sam python bench.py --workload sam --only --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name 'regex:render_kernel|mask_head_kernel' --launch-skip 4 -c 2 -f \
    -o $O/${R}_mask python bench.py --workload mask --only --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
ls -la $O/${R}_* | awk '{print $5, $9}'
