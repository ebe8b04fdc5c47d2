#!/bin/bash
# C3 (N = 65536 FP64) on G GPUs: stream-K shapes and one phase vs two.  usage: tools/scale_session3.sh G
G=$1; OUT=gpurun_out; P=29800
run() {
  name=$1; shift
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P bench.py --gpus $G --steps 20 --warmup 3 --precision f64 --bodies 65536 --no-cpu-baseline --no-energy "$@" \
      > $OUT/r02c_${name}_n${G}.json 2> $OUT/r02c_${name}_n${G}.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/r02c_${name}_n${G}.json"))
    print("$name G=$G: %.1f G inter/s, %.4f ms/step, b2b %.1f (%.4f ms), variant %s grid %s phases %s" % (d["value"], d["ms_per_step"], d["value_back_to_back"], d["ms_per_step_back_to_back"], d["config"]["force_variant"], d["config"].get("stream_grid"), d["config"].get("stream_phases")))
except Exception as e:
    print("$name G=$G FAILED:", e); print(open("$OUT/r02c_${name}_n${G}.err").read()[-800:])
PY
}
run c3_default
run c3_onephase --overlap 0
run c3_v6 --variant 6
run c3_v6_onephase --variant 6 --overlap 0
run c3_v7 --variant 7
run c3_splitgrid --stream 0
