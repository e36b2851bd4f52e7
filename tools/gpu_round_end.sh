#!/bin/bash
# One GPU-box pass: parity tests, smoke, the three bench workloads, then the ncu evidence for profiles/.
#   gpurun --timeout 900 -- 'bash tools/gpu_round_end.sh'
# Everything lands in gpurun_out/ (summaries are made afterwards with tools/ncu_summary.py).
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; tail -2 $O/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for w in rgb sam mask; do
    timeout 200 python bench.py --workload $w 2> $O/final_bench_$w.err > $O/final_bench_$w.json; cut -c1-170 $O/final_bench_$w.json
done
timeout 200 python bench.py --impl reference > $O/final_bench_ref.json 2>/dev/null; cut -c1-170 $O/final_bench_ref.json
for w in rgb sam mask; do
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches_$w.csv \
        python bench.py --workload $w --steps 2 --warmup 1 > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:render_kernel --launch-skip 2 -c 1 -f \
    -o $O/final_rgb python bench.py --workload rgb --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name 'regex:render_kernel|mask_head_kernel' --launch-skip 4 -c 2 -f \
    -o $O/final_mask python bench.py --workload mask --steps 2 --warmup 3 > /dev/null 2>&1
ls -la $O/final_* | awk '{print $5, $9}'
