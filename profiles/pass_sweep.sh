#!/bin/bash
# pass_sweep.sh tag "pf1 pf2 ..." : device-resident leg of bench.py at several internal pass sizes (env passes through)
tag=$1; shift
for pf in $1; do
  python bench.py --no-cpu --no-find --no-pbc --steps 12 --warmup 3 --pass-frames $pf > gpurun_out/${tag}_pf$pf.json 2> gpurun_out/${tag}_pf$pf.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_pf$pf.json"))
print("$tag pf=$pf value=%.0f ms/step=%.3f stages=%s" % (d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}))
PY
done
