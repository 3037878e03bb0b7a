cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_detector_stress.py -x -q -k "12mhz" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_zz_gpu_classify.py -x -q -s -k "plugin or linked" 2>&1 | grep -E "^\[|passed|failed|Error|assert" | tail -8
