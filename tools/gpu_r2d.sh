# after the FFT occupancy fix: parity, the device timeline of the bench recording, a full capture of the FFT kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python tools/dev_timeline.py 60 4 > gpurun_out/tl_r2d.log 2>&1
grep "ms_total\|ms_detect" gpurun_out/tl_r2d.log | tail -2 | cut -c1-700
grep "^host: every" gpurun_out/tl_r2d.log | tail -1
timeout 600 ncu -k regex:"k_detect_fft|k_fir_ws" --launch-skip 4 -c 6 --set full --clock-control none --import-source on -o gpurun_out/r2d_fft_full -f python tools/dev_timeline.py 20 1 > gpurun_out/ncu_full3.log 2>&1
python tools/ncu_digest.py gpurun_out/r2d_fft_full.ncu-rep gpurun_out/r2d_fft_full_summary.csv
cut -c1-330 gpurun_out/r2d_fft_full_summary.csv
