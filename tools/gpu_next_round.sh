# first GPU call of the next round: what the final build of round 2 has not been measured on
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# 1. the case written after round 2's last GPU minute (a chunk boundary with 112 bursts alive)
timeout 300 python -m pytest tests/gpu_crowded_boundary_cases.py tests/test_gpu_detector_stress.py -x -q -m gpu -k "crowded or pieces" 2>&1 | tail -4
# 2. bench lines of the final build (config 2 was measured; 3 and 4 predate the last FIR staging fix by one commit)
for c in 3 4; do timeout 600 python bench.py --config $c 2>&1 | tail -1 > gpurun_out/r3_bench_cfg$c.json; cut -c1-200 gpurun_out/r3_bench_cfg$c.json; done
# 3. FIR at DEC = 48 alone (never captured): 240-output tiles, integer staging
timeout 600 ncu -k regex:"k_fir_ws" --launch-skip 2 -c 2 --set full --clock-control none --import-source on -o gpurun_out/r3_fir48_full -f python bench.py --config 3 --seconds 10 --steps 1 --warmup 1 --cpu-seconds 0.5 > gpurun_out/ncu_fir48.log 2>&1
python tools/ncu_digest.py gpurun_out/r3_fir48_full.ncu-rep gpurun_out/r3_fir48_full_summary.csv && cut -c1-330 gpurun_out/r3_fir48_full_summary.csv
# (multi-GPU: `gpurun --gpus 8 -- bash tools/gpu_run_multi.sh` -- the r2c weak / strong scaling tables predate the FIR peeling,
#  the crowded-segment walker and the 12 MHz fixes)
