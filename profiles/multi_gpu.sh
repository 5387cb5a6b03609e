#!/bin/bash
# multi_gpu.sh N tag : bench.py under torchrun on N GPUs, the C++ drop-in on N GPUs, the sharded-driver test
N=$1; tag=$2; lite=$3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-find \
  > gpurun_out/${tag}_n$N.json 2> gpurun_out/${tag}_n$N.err
tail -c 300 gpurun_out/${tag}_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_n$N.json").read().strip().splitlines()[-1])
print("N=$N value=%.0f e2e=%.0f sustained=%s pcie=%s stats_reduce=%s" % (d["value"], d["e2e"]["value"], d["sustained"] and round(d["sustained"]["value"]), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["pcie"].items() if k!="note"}, d["stats_reduce"]))
PY
if [ -z "$lite" ]; then gmx-acqueduct_b200/wn_bench -gpus $N -batches 12 -warmup 2 -find 0 > gpurun_out/${tag}_cpp_n$N.jsonl 2>> gpurun_out/${tag}_n$N.err; cut -c1-400 gpurun_out/${tag}_cpp_n$N.jsonl; fi
gmx-acqueduct_b200/wn_bench -gpus $N -batches 12 -warmup 2 -find 0 -mode stage >> gpurun_out/${tag}_cpp_n$N.jsonl 2>> gpurun_out/${tag}_n$N.err; tail -1 gpurun_out/${tag}_cpp_n$N.jsonl | cut -c1-400
if [ -z "$lite" ]; then python -m pytest tests/test_gpu_host_cpp.py -q -m gpu -k "sharding" 2>&1 | tail -2; fi
