#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02q_pytest.log; tail -6 gpurun_out/r02q_pytest.log | cut -c1-300
timeout 900 python tools/cli_e2e.py --pairs 400000 --levels 1000000 --alleles 200 > gpurun_out/r02q_cli.json 2> gpurun_out/r02q_cli.err; echo "cli rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r02q_cli.json')); print(d['wall_s'], d['phases_s'], d['info'])"
