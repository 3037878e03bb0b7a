# dense traffic through the generic segment walker: the detector stress cases, then configs 3 and 4 with counters
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_detector_stress.py -x -q 2>&1 | tail -8
for c in 4 3; do
  IR_SCAN_DEBUG=1 timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/gen_cfg$c.json 2> gpurun_out/gen_cfg$c.err
  tail -c 1500 gpurun_out/gen_cfg$c.json; grep "seg scan" gpurun_out/gen_cfg$c.err | tail -2
done
