#!/bin/bash
mkdir -p gpurun_out
HLALA_TYPING_PROFILE=1 timeout 900 python tools/cli_e2e.py --pairs 200000 --levels 600000 --alleles 1000 --gpus 2 > gpurun_out/r02n_cli.json 2> gpurun_out/r02n_cli.err; echo "cli rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r02n_cli.json')); print(json.dumps(d, indent=1)[:3500])"
