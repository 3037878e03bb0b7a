# dense traffic through the generic segment walker: the detector stress cases, then configs 4 and 3 with counters
# and (config 3) the host / wave timeline of one step
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_detector_stress.py -x -q -k "dense or 12mhz or squelch" 2>&1 | tail -4
IR_SCAN_DEBUG=1 timeout 600 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/gen_cfg4.json 2> gpurun_out/gen_cfg4.err
python -c "
import json; d=json.loads(open('gpurun_out/gen_cfg4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()})"
grep "seg scan" gpurun_out/gen_cfg4.err | tail -1
IR_CHUNK_DEBUG=1 IR_SCAN_DEBUG=1 timeout 900 python bench.py --config 3 --steps 1 --warmup 3 > gpurun_out/tl_cfg3.json 2> gpurun_out/tl_cfg3.err
python -c "
import json; d=json.loads(open('gpurun_out/tl_cfg3.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()})"
grep "^host:" gpurun_out/tl_cfg3.err | tail -4
grep -c "^wave" gpurun_out/tl_cfg3.err
