set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_detector_stress.py -x -q 2>&1 | tail -4
timeout 300 python tools/dev_timeline.py 60 3 2>&1 | grep -v "^scan \|^stream scan" | tail -14
