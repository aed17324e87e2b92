#!/bin/bash
# Multi-GPU evidence in one gpurun call (row g of the coverage table: BASELINE.json configs[2..4] at
# their GPU counts, the distributed parity tests, per-rank phases).  Usage on the GPU box:
#   bash tools/run_multigpu_suite.sh <ngpus> <tag>        -> gpurun_out/<tag>_*.json / .log
N=${1:-8}; TAG=${2:-r02_multi}
OUT=gpurun_out; mkdir -p $OUT
run() {  # run <gpus> <name> <bench args...>
  local g=$1 name=$2; shift 2
  local port=$((29700 + RANDOM % 200))
  if [ "$g" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > $OUT/${TAG}_${name}_n1.json 2> $OUT/${TAG}_${name}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $g "$@" > $OUT/${TAG}_${name}_n${g}.json 2> $OUT/${TAG}_${name}_n${g}.err
  fi
  echo "== $name gpus=$g rc=$? $(head -c 300 $OUT/${TAG}_${name}_n${g}.json)"
}
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader | head -8
# distributed parity tests (torchrun worlds 2..4, FXB200_NO_PEER, single-process multi-GPU context)
timeout 1500 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x -s 2>&1 | tail -30 > $OUT/${TAG}_dist_tests.log
tail -5 $OUT/${TAG}_dist_tests.log
run 1 headline --steps 8 --warmup 3 --no-cpu-baseline --transform-m 0
for g in 2 4 8; do
  [ $g -le $N ] || continue
  run $g headline --steps 8 --warmup 3 --no-cpu-baseline
  run $g C3 --config C3 --steps 5 --warmup 3 --no-cpu-baseline --transform-m 0
done
run $N C4 --config C4 --steps 2 --warmup 1
run $N C5 --config C5 --steps 3 --warmup 2 --no-cpu-baseline
if [ "$N" = "8" ]; then   # A/B of the solver's latency hiding at 8 GPUs
  FXB200_NO_LAG=1 run 8 headline_nolag --steps 8 --warmup 3 --no-cpu-baseline --transform-m 0
  FXB200_MAX_BLOCKS=5 run 8 headline_mb5 --steps 8 --warmup 3 --no-cpu-baseline --transform-m 0
  FXB200_MAX_BLOCKS=3 run 8 headline_mb3 --steps 8 --warmup 3 --no-cpu-baseline --transform-m 0
fi
