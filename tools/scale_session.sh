#!/bin/bash
# Multi-GPU measurement session (round 2): the BASELINE configs nobody had run at every rank count.
# usage: tools/scale_session.sh G      (G = 2, 4 or 8; run under `gpurun --gpus G`)
G=$1; OUT=gpurun_out; P=29600
run() {  # name, extra args...
  name=$1; shift
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P bench.py --gpus $G "$@" \
      > $OUT/r02_${name}_n${G}.json 2> $OUT/r02_${name}_n${G}.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/r02_${name}_n${G}.json"))
    print("$name G=$G: %.1f G inter/s, %.3f ms/step, b2b %.1f, e2e %.1f, launches/step %s, parity %s" % (d["value"], d["ms_per_step"], d["value_back_to_back"], d["e2e"]["value"], d["config"].get("launches_per_step"), d.get("parity")))
except Exception as e:
    print("$name G=$G FAILED:", e); print(open("$OUT/r02_${name}_n${G}.err").read()[-1500:])
PY
}
WEAK=$(python -c "import math; print(int(round(1048576*math.sqrt($G)/1024))*1024)")
run c4_push --steps 10 --warmup 3
run c4_nccl --steps 10 --warmup 3 --exchange nccl
run c3_f64 --steps 20 --warmup 3 --precision f64 --bodies 65536 --no-cpu-baseline
run c5 --steps 2 --warmup 1 --bodies 4194304 --no-energy
run weak --steps 3 --warmup 1 --bodies $WEAK --no-energy
run c2 --steps 20 --warmup 3 --bodies 131072
