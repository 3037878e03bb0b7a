# multi-GPU session: two-device cases, strong-scaling by time blocks, weak-scaling bench with the concurrent-copy floor
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
nproc; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)"; free -g | head -2
timeout 900 python -m pytest tests/test_zz_gpu_classify.py -x -q -s -k "two_devices" 2>&1 | grep -E "^\[|passed|failed|Error" | tail -4
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python tools/bench_blocks.py --seconds 60 --steps 3 --warmup 2 2>&1 | tail -1 > gpurun_out/r2c_blocks_n$n.json
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/bench_blocks.py --seconds 60 --steps 3 --warmup 2 2>&1 | tail -1 > gpurun_out/r2c_blocks_n$n.json
    fi
    cut -c1-900 gpurun_out/r2c_blocks_n$n.json
  fi
done
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 --cpu-seconds 1 2>&1 | tail -1 > gpurun_out/r2c_bench_n$n.json
    python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_bench_n$n.json').read())
print($n, d['value'], d['ms_per_step'], json.dumps(d['e2e']))
PY
  fi
done
