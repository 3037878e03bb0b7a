# round-2 profile set: launch list of the bench command, DRAM traffic per kernel group, full captures of the hot kernels
cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0.5 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_seg|k_detect|k_fir|k_chain|k_demod" -c 1200 --csv --log-file gpurun_out/r2_traffic_launches.csv python tools/dev_timeline.py 60 1 > gpurun_out/ncu_traffic.log 2>&1
grep "^RUNJSON" gpurun_out/ncu_traffic.log | tail -1 | sed 's/^RUNJSON //' > gpurun_out/r2_traffic_run.json
python tools/ncu_traffic.py gpurun_out/r2_traffic_launches.csv gpurun_out/r2_traffic_run.json gpurun_out/r2_kernel_traffic.json | tail -40
timeout 1200 ncu -k regex:"k_detect_fft|k_detect_classify|k_seg_walk|k_seg_base|k_seg_gather|k_fir_ws|k_chain|k_demod" --launch-skip 6 -c 30 --set full --clock-control none --import-source on -o gpurun_out/r2_hot_full -f python tools/dev_timeline.py 60 1 > gpurun_out/ncu_full.log 2>&1
python tools/ncu_digest.py gpurun_out/r2_hot_full.ncu-rep gpurun_out/r2_hot_full_summary.csv
cut -c1-230 gpurun_out/r2_hot_full_summary.csv | head -40
