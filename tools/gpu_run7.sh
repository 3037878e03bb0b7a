set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu -k regex:"k_fir" --launch-skip 1 -c 2 --set full --clock-control none --import-source on -o gpurun_out/r2_fir_ws -f python tools/dev_timeline.py 60 1 > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_c.log
python tools/ncu_digest.py gpurun_out/r2_fir_ws.ncu-rep gpurun_out/r2_fir_ws_summary.csv
cat gpurun_out/r2_fir_ws_summary.csv
ncu -i gpurun_out/r2_fir_ws.ncu-rep --page source --csv --print-source sass 2>/dev/null > gpurun_out/r2_fir_ws_source.csv
wc -l gpurun_out/r2_fir_ws_source.csv
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refapi.py -x -q 2>&1 | tail -4
