# quick post-change check: path parity + the device-resident timeline
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refapi.py -x -q 2>&1 | tail -3
timeout 300 python tools/dev_timeline.py 60 4 2>&1 | grep "^wave\|ms_total\|^host: every" | tail -9
