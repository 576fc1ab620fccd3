#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_typing.py -x -q 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --strong-pairs 1000000 > gpurun_out/r02k_bench2.json 2> gpurun_out/r02k_bench2.err; echo "rc $?"
grep "\[bench\]" gpurun_out/r02k_bench2.err; tail -3 gpurun_out/r02k_bench2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['n_gpus']); print(json.dumps(d['strong_scaling_config3'], indent=1))
PY
