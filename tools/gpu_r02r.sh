#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_long_reads.py -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02r_pytest.log; tail -30 gpurun_out/r02r_pytest.log | cut -c1-400
timeout 1200 python tools/long_reads_probe.py 2000 8000 100 > gpurun_out/r02r_lr.json 2> gpurun_out/r02r_lr.err; echo "probe rc $?"; cat gpurun_out/r02r_lr.json; tail -5 gpurun_out/r02r_lr.err
