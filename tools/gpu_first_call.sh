#!/bin/bash
# First GPU call of a round: everything that was committed without a B200 run, then the standing checks.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# Outputs land in gpurun_out/ (copy what is worth keeping into profiles/).
mkdir -p gpurun_out
R=${ROUND:-r2}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1
timeout 900 python -m pytest tests/test_zz_gpu_classify.py -q -m gpu > gpurun_out/${R}_pytest_classify.log 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
timeout 600 python tools/bench_classify.py > gpurun_out/${R}_bench_classify.json 2> gpurun_out/${R}_bench_classify.err
timeout 600 python tools/bench_blocks.py --seconds 30 > gpurun_out/${R}_bench_blocks_n1.json 2> gpurun_out/${R}_bench_blocks_n1.err
# (with --gpus 2: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_blocks.py)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_classify_frames -c 2 \
    -o gpurun_out/${R}_classify python tools/bench_classify.py 4 > gpurun_out/${R}_ncu_classify.log 2>&1
tail -5 gpurun_out/${R}_pytest_classify.log gpurun_out/${R}_pytest_gpu.log
cat gpurun_out/${R}_bench.json gpurun_out/${R}_bench_classify.json
