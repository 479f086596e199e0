#!/bin/bash
# Run the device-resident bench once per tuning variant of the library (build_variants/liblars_*.so).
for so in build_variants/liblars_*.so; do
  name=$(basename $so .so)
  LARS_B200_LIB=$PWD/$so timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', 'value %.0f' % d['value'], 'step_ms %.4f' % d['ms_per_step'], 'k2_ms %.4f' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'])"
done
