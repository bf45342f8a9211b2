#!/bin/bash
# Benchmark lines of all five named shapes at N GPUs (one node): usage  tools/run_all_shapes.sh N [steps]
# c3 (default line; for N >= 2 it also carries the node-partitioned config-4 block), c1, c2, c5 data-parallel, c4 node-partitioned
# (fp32 and bf16-storage).  Writes gpurun_out/shapes_n${N}_<workload>.json (the JSON line only).
N=${1:-1}; STEPS=${2:-5}; PORT=29600
mkdir -p gpurun_out
run() {  # name, extra args...
  local name=$1; shift
  PORT=$((PORT+1))
  if [ "$N" -gt 1 ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps $STEPS --warmup 3 "$@" 2> gpurun_out/shapes_n${N}_${name}.err | grep "^{" > gpurun_out/shapes_n${N}_${name}.json
  else
    timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup 3 "$@" 2> gpurun_out/shapes_n${N}_${name}.err | grep "^{" > gpurun_out/shapes_n${N}_${name}.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/shapes_n${N}_${name}.json"))
    np_ = d.get("node_partitioned") or {}
    print("${name} N=${N}: value %.4g  ms/step %.2f  e2e %s  node_partitioned %s" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), (np_.get("value"), np_.get("ms_per_step"), np_.get("snapshots"))))
except Exception as exc:
    print("${name} N=${N}: FAILED", exc)
PY
}
run c3
run c1 --workload c1
run c2 --workload c2
run c5 --workload c5
if [ "$N" -gt 1 ]; then
  run c4 --workload c4 --steps 3
  run c4_bf16 --workload c4 --steps 3 --qkv-storage bf16 --no-e2e
fi
