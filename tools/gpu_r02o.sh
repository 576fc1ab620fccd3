#!/bin/bash
mkdir -p gpurun_out
HLALA_TYPING_PROFILE=1 timeout 900 python tools/cli_e2e.py --pairs 200000 --levels 600000 --alleles 1000 > gpurun_out/r02o_cli.json 2> gpurun_out/r02o_cli.err; echo "cli rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r02o_cli.json')); print(d['phases_s']); print(d['stderr'][-3000:])"
