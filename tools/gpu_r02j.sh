#!/bin/bash
mkdir -p gpurun_out
python -X faulthandler tools/scale_parity.py --levels 600000 --genes 17 --alleles 1000 --haps 8 --pairs 200000 --check-pairs 300 --dir /tmp/sp > gpurun_out/r02j_sp.json 2> gpurun_out/r02j_sp.err; echo "rc $?"
tail -30 gpurun_out/r02j_sp.err | cut -c1-300
python -X faulthandler tools/scale_parity.py --levels 100000 --genes 4 --alleles 100 --haps 8 --pairs 20000 --check-pairs 300 --dir /tmp/sp2 > gpurun_out/r02j_sp2.json 2> gpurun_out/r02j_sp2.err; echo "rc $?"
tail -30 gpurun_out/r02j_sp2.err | cut -c1-300; cat gpurun_out/r02j_sp2.json | cut -c1-600
