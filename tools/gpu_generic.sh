# dense traffic / integer formats: parity cases, then configs 4 and 3 (bench lines for profiles/)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_detector_stress.py -x -q -k "not full_size" 2>&1 | tail -4
for c in 4 3; do
  IR_SCAN_DEBUG=1 timeout 900 python bench.py --config $c > gpurun_out/r2d_bench_cfg$c.json 2> gpurun_out/r2d_bench_cfg$c.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_cfg$c.json').read().strip().splitlines()[-1]); print($c, d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()}, d['roofline'].get('fp32'))"
  grep "seg scan" gpurun_out/r2d_bench_cfg$c.err | tail -1
done
