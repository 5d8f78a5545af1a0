# usage: tools/run_multi.sh N TAG   (under gpurun --gpus N)
N=$1; TAG=$2
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multi_nccl.py -x -q -m gpu > gpurun_out/pytest_nccl_$TAG.log 2>&1; tail -n 3 gpurun_out/pytest_nccl_$TAG.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("lm"), indent=0)[:1500])
print(json.dumps(d.get("sharded"), indent=0)[:2500])
PY
