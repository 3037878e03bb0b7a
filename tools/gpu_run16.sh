cd $GRAFT_REPO_ROOT
timeout 900 ncu -k regex:"k_fir_ws" --launch-skip 1 -c 1 --set full --clock-control none --import-source on -o gpurun_out/r2_fir_ws2 -f python tools/dev_timeline.py 60 1 > gpurun_out/ncu_d.log 2>&1
python tools/ncu_digest.py gpurun_out/r2_fir_ws2.ncu-rep gpurun_out/r2_fir_ws2_summary.csv
cat gpurun_out/r2_fir_ws2_summary.csv | tail -1
ncu -i gpurun_out/r2_fir_ws2.ncu-rep --page source --csv --print-source sass 2>/dev/null > gpurun_out/r2_fir_ws2_source.csv
