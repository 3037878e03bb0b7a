set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refapi.py -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_detector_stress.py -x -q -k "dense or full_size or ragged" 2>&1 | tail -8
timeout 300 python tools/dev_timeline.py 60 3 2>&1 | grep -v "^scan \|^stream scan" | tail -22
IR_FIR_LEGACY=1 timeout 300 python tools/dev_timeline.py 60 3 2>&1 | grep -v "^scan \|^stream scan" | tail -3
