cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refapi.py -x -q 2>&1 | tail -3
timeout 300 python tools/dev_timeline.py 60 3 2>&1 | grep -v "^scan \|^stream scan\|^host" | tail -13
IR_SCAN_DEBUG=1 timeout 300 python tools/chunk_timeline.py 2>&1 | grep -v "^scan \|^stream scan" | tail -56
IR_SCAN_DEBUG=1 timeout 600 python bench.py --config 3 --seconds 10 --steps 1 --warmup 1 --cpu-seconds 0.5 2>&1 | grep "seg scan" | tail -2
