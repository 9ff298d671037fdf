#!/bin/bash
# Round-2 scaling session on N GPUs of one box: weak (B = 256 per GPU) and strong (global B = 2048, shards balanced by sum n^2)
# guided sampler step.  usage: gpurun --gpus N -- tools/gpu_r2_scale.sh N
set -u
N=${1:-2}; O=gpurun_out/r2scale; mkdir -p $O
run() { # name, extra args
  if [ "$N" = "1" ]; then timeout 900 python bench.py --gpus 1 --no-cpu-baseline $2 > $O/$1_n$N.json 2> $O/$1_n$N.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
       bench.py --gpus $N --no-cpu-baseline $2 > $O/$1_n$N.json 2> $O/$1_n$N.err; fi
  tail -c 900 $O/$1_n$N.json | head -c 900; echo; tail -2 $O/$1_n$N.err
}
run weak_guided ""
run strong_guided_b2048 "--strong --batch 2048"
python - <<P
import json
for n in ("weak_guided","strong_guided_b2048"):
    try:
        d=json.loads([l for l in open("$O/%s_n$N.json"%n) if l.startswith("{")][-1])
        print(n, "N=$N", round(d["ms_per_step"],3), "ms/step", round(d["value"],2), "mol/s e2e", round(d["e2e"]["value"],2), d.get("per_rank_ms_per_step"))
    except Exception as e: print(n, "ERR", e)
P
