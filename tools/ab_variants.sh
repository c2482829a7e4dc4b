#!/bin/bash
# Dev tool: rebuild libisb.so on the GPU box with different -D tuning flags and run tools/ab_env.py on each build.
#   tools/ab_variants.sh "name1:-DISB_WARP_BLOCK_H=64" "name2:-DISB_WARP_MIN_CTAS=3 -DISB_WARP_BLOCK_H=128" ...
# (the flags are part of the library's source hash, so they must stay exported while the build is in use)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  export ISB_NVCC_FLAGS="$flags"
  python -m image_stitching_b200.build --force > gpurun_out/ab_$name.build 2>&1
  echo "== $name ($flags)"
  python tools/ab_env.py --variants "${AB_ENV:-ISB_PDL=1}" --steps 30 --workload "${AB_WORKLOAD:-cfg2}" --out gpurun_out/ab_$name.json 2>&1 | tail -2
done
unset ISB_NVCC_FLAGS
python -m image_stitching_b200.build --force > /dev/null 2>&1   # leave the default build behind
