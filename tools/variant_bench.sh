#!/bin/bash
# Kernel experiments on the GPU box: for every library variant built under sanerf_hq_b200/lib_<name>/ (SANERF_LIB_VARIANT, see
# sanerf_hq_b200/build.py) run the parity tests that exercise the fused kernels, then a short bench of the given workloads.
#   gpurun -- 'bash tools/variant_bench.sh "base xpair" "rgb"'
set -u
O=gpurun_out
mkdir -p $O
for v in $1; do
    export SANERF_LIB_VARIANT=$v
    [ "$v" = base ] && unset SANERF_LIB_VARIANT
    if [ "${3:-test}" = test ]; then
        timeout 600 python -m pytest tests/test_render_gpu.py tests/test_render_options_gpu.py tests/test_heads_gpu.py -m gpu -x -q -p no:cacheprovider -W ignore 2>&1 | tail -2 | sed "s/^/[$v] /"
    fi
    for w in $2; do
        timeout 300 python bench.py --workload $w --only --no-cpu-baseline --steps 15 --warmup 3 2>/dev/null > $O/var_${v}_$w.json
        python - <<PY
import json
for ln in open("$O/var_${v}_$w.json"):
    if ln.startswith("{"):
        b = json.loads(ln)
        print("[$v] $w", round(b["value"], 2), "Mrays/s", round(b["ms_per_step"], 3), "ms", [(k["kernel"].split("::")[1], round(k["ms"], 3)) for k in b["kernels"]], b.get("clocks", {}).get("sm_mhz"))
PY
    done
done
