#!/bin/bash
# Multi-GPU evidence on one box (run through `gpurun --gpus N -- bash scripts/evidence_multi.sh N`): BASELINE configs 3 / 4 / 5 at N ranks
# (torchrun, one rank per GPU, NCCL), the bare H2D copy ceiling of the host, and the 2-rank gathered-metadata test.
N=${1:-8}
PORT=29517
run() {   # config, extra args
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --config $1 --steps 10 --warmup 3 \
        > gpurun_out/r2_cfg$1_${N}gpu.json 2> gpurun_out/r2_cfg$1_${N}gpu.err
    tail -c 300 gpurun_out/r2_cfg$1_${N}gpu.err
    PORT=$((PORT + 1))
}
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
for c in 3 4 5; do run $c; done
if [ "$N" -ge 8 ]; then timeout 300 python scripts/h2d_bw.py > gpurun_out/r2_h2d_bw_8gpu.jsonl 2> gpurun_out/r2_h2d_bw_8gpu.err; fi
if [ "$N" -ge 2 ]; then timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2 > gpurun_out/r2_pytest_multi_${N}gpu.txt; cat gpurun_out/r2_pytest_multi_${N}gpu.txt; fi
python - <<PY
import json
for c in (3, 4, 5):
    try:
        d = json.loads(open("gpurun_out/r2_cfg%d_${N}gpu.json" % c).read().strip().splitlines()[-1])
        print("config", c, "N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "verified", d["verified"], "meta", d["meta_verified"], "h2d", d["e2e"]["h2d_gbs_per_rank"], d["clocks"]["reasons"])
    except Exception as e:
        print("config", c, "failed", e)
PY
