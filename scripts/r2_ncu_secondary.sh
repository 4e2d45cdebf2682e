#!/bin/bash
# ncu --set full of the secondary kernels at reduced sizes (evidence + tuning): MC rollouts, car kernels, tc sweep, knn
OUT=gpurun_out/ncu2; mkdir -p $OUT
run() {  # name, kernel regex, count, config, scale
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -c $3 -o $OUT/$1 -f \
      python bench_configs.py --configs $4 --scale $5 > $OUT/$1.log 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.csv 2>/dev/null
  python scripts/ncu_summary.py $OUT/$1.csv > $OUT/$1_summary.txt 2>&1
  tail -40 $OUT/$1_summary.txt
}
run mc 'mc_rollout_kernel' 1 C5 0.02
run cars 'car_cost_kernel|car_edges_free_kernel' 4 F4 0.25
run tc 'tc_rball_kernel' 1 C3 0.1
