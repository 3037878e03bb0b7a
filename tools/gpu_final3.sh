# after the FIR sample-buffer pitch fix: every GPU test, config 3 again, a full capture of the FIR kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --config 3 2>&1 | tail -1 > gpurun_out/r2d_bench_cfg3.json; cut -c1-300 gpurun_out/r2d_bench_cfg3.json
timeout 600 ncu -k regex:"k_fir_ws" --launch-skip 1 -c 2 --set full --clock-control none --import-source on -o gpurun_out/r2d_fir_full -f python tools/dev_timeline.py 60 1 > gpurun_out/ncu_full4.log 2>&1
python tools/ncu_digest.py gpurun_out/r2d_fir_full.ncu-rep gpurun_out/r2d_fir_full_summary.csv && cut -c1-330 gpurun_out/r2d_fir_full_summary.csv
