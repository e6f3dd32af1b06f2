#!/bin/bash
# usage (under gpurun --gpus 8): tools/gpu_scale_final.sh TAG - parity check at N = 4, 8 and the bench line at N = 8, 4, 2, 1 of one box
set -u
TAG=$1
mkdir -p gpurun_out
for N in 8 4; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 \
      | grep "multi-gpu" | tee gpurun_out/multi_gpu_check_${TAG}_n${N}.log | tail -1
done
for N in 8 4 2; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 400 --warmup 10 2>&1 \
      | grep "^{" > gpurun_out/bench_${TAG}_n${N}_fused.json
done
timeout 100 python bench.py --steps 400 --warmup 10 --no-extras --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/bench_${TAG}_n1_quick.json
python - <<PY
import json
for n in (1, 2, 4, 8):
    f = "gpurun_out/bench_${TAG}_n%d_%s.json" % (n, "quick" if n == 1 else "fused")
    try:
        d = json.load(open(f))
        ex = d.get("exchange", {})
        ar = (ex.get("nccl_allreduce") or {})
        ue = (ex.get("unpack_every_step") or {})
        print(n, "value %.0f GB/s  ms/step %.4f  kernel %.4f  gap %.1f us  e2e %.0f  unpack-each %.0f  allreduce %.0f GB/s (%s)" % (
            d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], ex.get("step_minus_kernel_us", 0), d["e2e"]["value"], ue.get("value", 0), ar.get("value", 0), ar.get("same_bytes_as_fused")))
    except Exception as e:
        print(n, "failed", e)
PY
