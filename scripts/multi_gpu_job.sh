#!/bin/bash
# One 8-GPU box: the headline at 8 GPUs, BASELINE configs[3] (chunk mode) and one configs[4] cell at 8
# (again with the per-rank batches rotated by one device if some rank is more than 3 % slower than the
# fastest: slow device or slow batch?), and the sharded decode test.  Lines go to gpurun_out/r2_multi_*.
cd "$(dirname "$0")/.." || exit 1
out=gpurun_out
run() {  # run <n> <devices> <port> <tag> <bench args...>
  n=$1; dev=$2; port=$3; tag=$4; shift 4
  CUDA_VISIBLE_DEVICES=$dev python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
      --master-port "$port" bench.py --gpus "$n" "$@" 2> "$out/r2_multi_${tag}_n$n.err" | tail -1 > "$out/r2_multi_${tag}_n$n.json"
}
QUICK="--steps 3 --warmup 3 --no-cpu --no-e2e --check-reads 8"
run 8 0,1,2,3,4,5,6,7 29527 c3 --steps 3 --warmup 3 --no-cpu --check-reads 64
run 8 0,1,2,3,4,5,6,7 29525 c4chunk --workload c4-chunk $QUICK
run 8 0,1,2,3,4,5,6,7 29526 c5 --workload c5:16:12:1 $QUICK
python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -3 > $out/r2_multi_gpu_test.txt
python - <<'PY'
import json, subprocess, sys
d = json.load(open("gpurun_out/r2_multi_c3_n8.json"))
ms = [r["kernel_ms"] for r in d["per_rank"]]
print("c3 N=8 per-rank kernel ms:", ["%.1f" % x for x in ms])
open("gpurun_out/r2_multi_need_rotate", "w").write("1" if max(ms) / min(ms) > 1.03 else "0")
PY
if [ "$(cat $out/r2_multi_need_rotate)" = "1" ]; then
  run 8 0,1,2,3,4,5,6,7 29528 c3rot --rotate 1 --steps 3 --warmup 3 --no-cpu --no-e2e --check-reads 8
fi
for f in $out/r2_multi_*_n*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e") or {}
    print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"], "ms %.1f" % d["ms_per_step"], "e2e %s" % ("%.4g" % e["value"] if e else "-"),
          "per-rank ms", ["%.1f" % r["kernel_ms"] for r in d["per_rank"]], d["parity_check"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
cat $out/r2_multi_gpu_test.txt
