cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_bench_n1.json; cut -c1-3000 gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_bench_cfg3.json; cut -c1-2500 gpurun_out/r2_bench_cfg3.json
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_bench_cfg4.json; cut -c1-2500 gpurun_out/r2_bench_cfg4.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/r2_bench_ref.json; cut -c1-1500 gpurun_out/r2_bench_ref.json
