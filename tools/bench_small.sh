#!/bin/bash
for f in 1 2 16; do
for so in build_variants/liblars_*.so; do
  name=$(basename $so .so)
  LARS_B200_LIB=$PWD/$so timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --frames $f 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frames $f', '$name', 'value %.0f' % d['value'], 'step_ms %.4f' % d['ms_per_step'], 'k2_ms %.4f' % d['roofline']['ms_per_launch'])"
done; done
