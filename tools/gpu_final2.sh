# end of round 2: every GPU test, the bench lines for profiles/, the launch list of one bench step, full captures of
# the FFT and FIR kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2d_bench_n1.json; cut -c1-900 gpurun_out/r2d_bench_n1.json
timeout 600 python bench.py --config 3 2>&1 | tail -1 > gpurun_out/r2d_bench_cfg3.json; cut -c1-300 gpurun_out/r2d_bench_cfg3.json
timeout 600 ncu -k regex:"k_detect_fft|k_fir_ws" --launch-skip 2 -c 4 --set full --clock-control none --import-source on -o gpurun_out/r2d_fft_fir_full -f python tools/dev_timeline.py 60 1 > gpurun_out/ncu_full3.log 2>&1
python tools/ncu_digest.py gpurun_out/r2d_fft_fir_full.ncu-rep gpurun_out/r2d_fft_fir_full_summary.csv && cut -c1-330 gpurun_out/r2d_fft_fir_full_summary.csv
timeout 600 ncu -k regex:"k_" --metrics gpu__time_duration.sum --clock-control none --launch-skip 1242 -c 850 --csv --log-file gpurun_out/r2d_bench_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0.5 > gpurun_out/ncu_bench.log 2>&1
grep -c "k_" gpurun_out/r2d_bench_launches.csv
