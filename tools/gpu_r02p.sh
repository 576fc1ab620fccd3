#!/bin/bash
mkdir -p gpurun_out
HLALA_TYPING_PROFILE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 0 --strong-pairs 1000000 > gpurun_out/r02p_bench2.json 2> gpurun_out/r02p_bench2.err; echo "rc $?"
grep -E "typing-profile|typing-device" gpurun_out/r02p_bench2.err | tail -150 | awk '{k=$1" "$4" "$5; if ($1=="[typing-profile]") k=$1" "$2; n[k]++; s[k]+=$(NF-1)} END{for (k in s) printf "%s n=%d sum=%.1f ms\n", k, n[k], s[k]}' | sort
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench2.json').read().strip().splitlines()[-1])
print(json.dumps(d['strong_scaling_config3'], indent=1))
PY
