#!/bin/bash
# Follow-up multi-GPU session (after the "peers seen" flag): headline config, C3 and C2.  usage: tools/scale_session2.sh G
G=$1; OUT=gpurun_out; P=29700
run() {
  name=$1; shift
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P bench.py --gpus $G "$@" \
      > $OUT/r02e_${name}_n${G}.json 2> $OUT/r02e_${name}_n${G}.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/r02e_${name}_n${G}.json"))
    print("$name G=$G: %.1f G inter/s, %.4f ms/step, b2b %.1f, e2e %.1f, launches/step %s" % (d["value"], d["ms_per_step"], d["value_back_to_back"], d["e2e"]["value"], d["config"].get("launches_per_step")))
except Exception as e:
    print("$name G=$G FAILED:", e); print(open("$OUT/r02e_${name}_n${G}.err").read()[-1500:])
PY
}
run c4_push --steps 10 --warmup 3
run c3_f64 --steps 20 --warmup 3 --precision f64 --bodies 65536 --no-cpu-baseline
run c2 --steps 20 --warmup 3 --bodies 131072
