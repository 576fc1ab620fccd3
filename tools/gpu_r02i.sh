#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02i_pytest.log; tail -5 gpurun_out/r02i_pytest.log
timeout 1500 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc $?"
grep "\[bench\]" gpurun_out/r02i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench.json'))
print(d['value'], d['e2e']['value'], d['check'])
print(json.dumps(d['stages']['typing'], indent=1)[:3000])
print(json.dumps(d['cpu_baseline'], indent=1)[:2500])
print(json.dumps(d['roofline']['chain_kernel'], indent=1), d['roofline']['aligned_chains'], d['roofline']['whole_path_achieved'])
PY
