#!/bin/bash
# Multi-GPU session of round 2 (run on the GPU box through gpurun --gpus N): the NCCL test at every visible GPU,
# the host<->device copy ceiling at 1..N concurrent ranks, and bench.py for the four workloads at the given GPU counts.
#   tools/run_scale.sh "<list of N>" "<workloads>" [extra bench flags]
# Writes gpurun_out/r02_bench_<workload>_n<N>.json (+ .err), r02_pytest_multigpu_n<G>.log, r02_pcie_probe_n<N>.log.
set -u
NS=${1:-"2"}
WLS=${2:-"c2 c4 c5 c3"}
shift 2 || true
EXTRA="$*"
mkdir -p gpurun_out
G=$(python -c "import torch; print(torch.cuda.device_count())")
echo "visible GPUs: $G"
timeout 900 python -m pytest tests/test_multigpu_nccl.py -m gpu -q -x --timeout 800 -rs > gpurun_out/r02_pytest_multigpu_n${G}.log 2>&1
echo "nccl pytest exit $? ($(tail -n 1 gpurun_out/r02_pytest_multigpu_n${G}.log))"
port=29500
for N in $NS; do
  port=$((port + 1))
  if [ "$N" = 1 ]; then
    timeout 300 python tools/pcie_probe.py > gpurun_out/r02_pcie_probe_n1.log 2>&1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      tools/pcie_probe.py > gpurun_out/r02_pcie_probe_n${N}.log 2>&1
  fi
  echo "pcie probe N=$N exit $?: $(grep -c aggregate gpurun_out/r02_pcie_probe_n${N}.log) lines"
  for wl in $WLS; do
    port=$((port + 1))
    steps=20; [ $wl = c3 ] && steps=5; [ $wl = c5 ] && steps=5; [ $wl = c4 ] && steps=20
    out=gpurun_out/r02_bench_${wl}_n${N}
    t0=$(date +%s)
    if [ "$N" = 1 ]; then
      timeout 900 python bench.py --workload $wl --gpus 1 --steps $steps --warmup 3 $EXTRA > $out.json 2> $out.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
        bench.py --workload $wl --gpus $N --steps $steps --warmup 3 $EXTRA > $out.json 2> $out.err
    fi
    echo "bench $wl N=$N exit $? in $(( $(date +%s) - t0 )) s: $(python -c "
import json,sys
try:
    d=json.load(open('$out.json')); print(round(d['value']), d['unit'], 'K2 frac', round(d['roofline']['frac'],3), 'e2e', round(d.get('e2e',{}).get('value',0)))
except Exception as e: print('no line', e)")"
  done
done
nvidia-smi topo -m > gpurun_out/r02_box_topology_n${G}.txt 2>&1; nproc >> gpurun_out/r02_box_topology_n${G}.txt; free -g >> gpurun_out/r02_box_topology_n${G}.txt
