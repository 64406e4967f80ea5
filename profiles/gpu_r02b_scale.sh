#!/bin/bash
# multi-GPU bench line (weak scaling value + strong_scaling block), launched the way the driver does
N=${N:-8}
OUT=gpurun_out/r2b
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
tail -c 600 $OUT/bench_n$N.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "strong", json.dumps(d.get("strong_scaling"))[:600])
PY
