#!/bin/bash
# Both bench arms at N GPUs of one box, for cover@1080p (BASELINE configs[1]) and the synthetic 1e5-shape scene at 8K
# (configs[4]; SKIP_SYN=1 leaves it out); results under gpurun_out/.   benchmarks/scaling_run.sh <N> [tag]
N=$1; TAG=${2:-r2}
run() {  # run <outfile> <bench args...>
  out=$1; shift
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; fi
  [ -s gpurun_out/$out.json ] || tail -c 1500 gpurun_out/$out.err
}
mkdir -p gpurun_out
run scale_${TAG}_ref_n$N --impl reference --steps 2 --warmup 1
run scale_${TAG}_n$N --steps 20 --warmup 3 --no-cold
[ -n "$SKIP_SYN" ] || run scale_${TAG}_syn1e5_n$N --scene synthetic:100000 --width 7680 --height 4320 --steps 5 --warmup 3 --no-cold
python - <<PY
import json
for name in ("scale_${TAG}_ref_n$N", "scale_${TAG}_n$N", "scale_${TAG}_syn1e5_n$N"):
    try:
        d = json.load(open(f"gpurun_out/{name}.json"))
        e = d.get("e2e") or {}
        print(name, "n", d["n_gpus"], "ms", round(d["ms_per_step"], 4), "Mrays/s", round(d["value"], 1), "e2e ms", e.get("ms_per_frame"), "family", d.get("family"),
              "e2e family", e.get("family"), "match", (d.get("frame") or {}).get("frame_matches_n1"), "cores", (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as exc:
        print(name, "FAILED", exc)
PY
