#!/bin/bash
# one multi-GPU box: bench lines for the configs given as arguments, e.g.  scale_trip.sh 8 7b-ssr 13b-actorder
N=$1; shift
for CFG in "$@"; do
  out=gpurun_out/r02_bench_${CFG}_n${N}.json
  STEPS=2; WARM=3; [ "$CFG" = "13b-actorder" ] && { STEPS=1; WARM=2; }
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
      bench.py --gpus $N --config $CFG --steps $STEPS --warmup $WARM > $out 2> gpurun_out/scale_trip.err
  echo "== $CFG N=$N rc=$?"; tail -2 gpurun_out/scale_trip.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("$out").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", (d.get("e2e") or {}).get("value"), "frac", d["roofline"]["frac"])
    print("parity", json.dumps(d.get("parity"))[:700])
    print("crit", json.dumps((d.get("critical_path") or {}).get("ranks")))
except Exception as e:
    print("no line:", e)
PY
done
