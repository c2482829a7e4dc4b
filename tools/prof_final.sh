set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
K='regex:blend|pyrdown|warp_tiles|seam|mirror'
for w in cfg2 cfg4 cfg3; do
  timeout 400 ncu --metrics $M --clock-control none -k "$K" -c 90 --csv --log-file gpurun_out/r02f_${w}_launches.csv python tools/ab_env.py --workload $w --steps 1 --variants ISB_PDL=1 > gpurun_out/r02f_${w}_ncu.log 2>&1
  echo "rc $w $?"
  python tools/traffic_from_launches.py gpurun_out/r02f_${w}_launches.csv | tail -25
done
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:warp_tiles|pyrdown_tma|blend_pipe' --launch-skip 15 --launch-count 5 -o gpurun_out/prof_r02f python tools/ab_env.py --workload cfg2 --steps 1 --variants ISB_PDL=1 > gpurun_out/r02f_full.log 2>&1
echo "rc full $?"
ls -la gpurun_out/prof_r02f*
