#!/bin/bash
# For each library variant and small frame-group size: kernel durations + DRAM bytes with natural L2 state.
for so in build_variants/liblars_*.so; do
  name=$(basename $so .so)
  for f in 1 2; do
    B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --frames $f"
    LARS_B200_LIB=$PWD/$so $B > /dev/null 2>&1 && LARS_B200_LIB=$PWD/$so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --cache-control none -k regex:"fused_index|wb_hist" -c 10 --csv --log-file gpurun_out/l2_${name}_f$f.csv $B > /dev/null 2>&1
  done
done
