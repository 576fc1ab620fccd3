#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/cli_e2e.py --pairs 200000 --levels 600000 --alleles 200 --gpus 2 > gpurun_out/r02m_cli.json 2> gpurun_out/r02m_cli.err; echo "cli rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r02m_cli.json')); print(json.dumps(d, indent=1)[:3500])"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --strong-pairs 1000000 > gpurun_out/r02m_bench2.json 2> gpurun_out/r02m_bench2.err; echo "rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['n_gpus']); print(json.dumps(d['strong_scaling_config3'], indent=1))
PY
timeout 900 python -m pytest tests/test_gpu_typing.py -x -q 2>&1 | tail -3
