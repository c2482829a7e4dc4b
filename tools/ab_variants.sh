#!/bin/bash
# Dev tool: rebuild libisb.so on the GPU box with different -D tuning flags and print the stage split of each.
#   tools/ab_variants.sh "name1:-DISB_WARP_BLOCK_H=64" "name2:-DISB_WARP_MIN_CTAS=3 -DISB_WARP_BLOCK_H=128" ...
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ISB_NVCC_FLAGS="$flags" python -m image_stitching_b200.build --force > gpurun_out/ab_$name.build 2>&1
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e-pipeline 2> gpurun_out/ab_$name.err | tail -1 > gpurun_out/ab_$name.json
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{name}.json"))
    print(name, "ms/step", d["ms_per_step"], "stages", d["roofline"].get("stages_ms"), "parity", d.get("parity"))
except Exception as e:
    print(name, "FAILED", e)
PY
done
# leave the default build behind
python -m image_stitching_b200.build --force > /dev/null 2>&1
