cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refapi.py -x -q 2>&1 | tail -3
for v in "" "IR_FIR_P128=1" "IR_FIR_SCALAR=1"; do echo "== $v"; env $v timeout 300 python tools/dev_timeline.py 60 3 2>&1 | grep "^wave  [12]\|ms_total" | tail -3; done
