#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_2gpu_bench.json 2> gpurun_out/r02_2gpu_bench.err; echo "bench rc $?"
grep "\[bench\]" gpurun_out/r02_2gpu_bench.err | cut -c1-250
python -c "
import json
d=json.loads(open('gpurun_out/r02_2gpu_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['n_gpus']); print(json.dumps(d['strong_scaling_config3'])[:900])"
timeout 900 python tools/cli_e2e.py --pairs 200000 --levels 1000000 --alleles 200 --gpus 2 > gpurun_out/r02_2gpu_cli.json 2> gpurun_out/r02_2gpu_cli.err; echo "cli rc $?"; tail -c 700 gpurun_out/r02_2gpu_cli.json
