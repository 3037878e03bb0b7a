# after the FIR staging fixes: every GPU test, then config 3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --config 3 2>&1 | tail -1 > gpurun_out/r2d_bench_cfg3.json; cut -c1-300 gpurun_out/r2d_bench_cfg3.json
python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_cfg3.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['raw_frames'], d['config']['bits_matching_ground_truth'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()})"
