set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -2
timeout 900 python -m pytest tests/test_gpu_detector_stress.py -x -q -k "not full_size" 2>&1 | tail -25
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -8
IR_SCAN_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-seconds 1 2>&1 | tail -30
