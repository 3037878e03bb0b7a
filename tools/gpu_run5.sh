set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu -k regex:"k_seg_base|k_seg_walk" --launch-skip 2 -c 4 --set full --clock-control none --import-source on -o gpurun_out/r2_seg_full -f python tools/dev_timeline.py 60 1 > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
python tools/ncu_digest.py gpurun_out/r2_seg_full.ncu-rep gpurun_out/r2_seg_full_summary.csv
cat gpurun_out/r2_seg_full_summary.csv
