cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c_bench_n1.json; cut -c1-1200 gpurun_out/r2c_bench_n1.json
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c_bench_cfg3.json; cut -c1-700 gpurun_out/r2c_bench_cfg3.json
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c_bench_cfg4.json; cut -c1-700 gpurun_out/r2c_bench_cfg4.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/r2c_bench_ref.json; cut -c1-400 gpurun_out/r2c_bench_ref.json
timeout 900 ncu -k regex:"k_" --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r2c_bench_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0.5 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu -k regex:"k_chain|k_demod|k_fir_ws" --launch-skip 3 -c 6 --set full --clock-control none --import-source on -o gpurun_out/r2c_burst_full -f python tools/dev_timeline.py 60 1 > gpurun_out/ncu_full2.log 2>&1
python tools/ncu_digest.py gpurun_out/r2c_burst_full.ncu-rep gpurun_out/r2c_burst_full_summary.csv
cut -c1-260 gpurun_out/r2c_burst_full_summary.csv
