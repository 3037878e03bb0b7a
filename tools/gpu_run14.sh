cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refapi.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --config 3 --seconds 20 --steps 2 --warmup 1 --cpu-seconds 0.5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['state_machine'], d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['kernels'])"
