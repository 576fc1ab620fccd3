#!/bin/bash
mkdir -p gpurun_out
HLALA_TYPING_PROFILE=1 timeout 1500 python bench.py --steps 1 --warmup 1 --e2e-steps 0 --stages 0 --cpu-pairs 400 --strong-single 1 --strong-pairs 500000 > gpurun_out/r02l.json 2> gpurun_out/r02l.err; echo "rc $?"
grep -E "typing-profile|\[bench\]" gpurun_out/r02l.err | tail -24
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l.json').read().strip().splitlines()[-1]); print(json.dumps(d['strong_scaling_config3'], indent=1))
PY
timeout 900 python -m pytest "tests/test_gpu_typing.py" -x -q 2>&1 | tail -30 | cut -c1-300
