#!/bin/bash
# 8-GPU box: which NCCL algorithm / protocol serves the 480-KB per-iteration all-reduce best (bench N = 8, resident leg)
export PYTHONPATH=$PWD
O=gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29720 + RANDOM % 200)) \
    bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 1 > $O/r2_nccl_$name.json 2> $O/r2_nccl_$name.err
  python - "$name" "$O/r2_nccl_$name.json" <<'PY'
import json, sys
try:
    j = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
    print(sys.argv[1], "ms/step", round(j["ms_per_step"], 3), "mv", round(j["roofline"]["avg_ms"], 4))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV
grep -h -i "nvls\|algo" $O/r2_nccl_default.err | head -8 | cut -c1-200
run tree NCCL_ALGO=Tree
run nvls NCCL_ALGO=NVLS
run ll128 NCCL_PROTO=LL128
run ll NCCL_PROTO=LL
run tree_ll NCCL_ALGO=Tree NCCL_PROTO=LL
run depth3 B200SFM_PCG_DEPTH=3
