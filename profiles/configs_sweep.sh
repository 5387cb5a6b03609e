#!/bin/bash
# configs_sweep.sh tag : the other BASELINE.json sizes through bench.py (device-resident / e2e / periodic), one line each
tag=$1
run() {  # waters frames_per_step steps
  python bench.py --waters $1 --frames-per-step $2 --steps $3 --warmup 3 --no-cpu --no-find --no-cpp --no-sustained > gpurun_out/${tag}_w$1.json 2> gpurun_out/${tag}_w$1.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_w$1.json"))
print("waters=$1 frames/step=$2 value=%.0f e2e=%.0f periodic=%.0f pair_evals/s=%.3g ms/step=%.3f frac=%.3f" % (d["value"], d["e2e"]["value"], d["periodic"]["value"] if d["periodic"] else 0, d["pair_evals_per_s"], d["ms_per_step"], d["roofline"]["frac"]))
PY
}
run 216 100 20
run 4000 500 20
run 333333 24 10
