#!/bin/bash
# Pass-1 (K1) time per tuning variant: whole step minus the fused pass, from bench.py's own numbers.
for so in build_variants/liblars_*.so; do
  name=$(basename $so .so)
  LARS_B200_LIB=$PWD/$so timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', 'value %.0f' % d['value'], 'step_ms %.4f' % d['ms_per_step'], 'k2_ms %.4f' % d['roofline']['ms_per_launch'], 'rest_ms %.4f' % (d['ms_per_step'] - d['roofline']['ms_per_launch']))"
done
