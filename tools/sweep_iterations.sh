#!/bin/bash
# BASELINE.json configs[4]: Sinkhorn iteration / epsilon sweep at batch 512 over G GPUs (run on the GPU box).
#   gpurun --gpus 2 -- './tools/sweep_iterations.sh 2 r01'
G=${1:-2}; R=${2:-r01}
mkdir -p gpurun_out
OUT=gpurun_out/${R}_sweep_g${G}.jsonl; : > $OUT
run() {  # workload images_per_gpu steps extra...
  local wl=$1 img=$2 steps=$3; shift 3
  if [ "$G" = 1 ]; then python bench.py --workload $wl --images $img --steps $steps --warmup 3 --no-cpu-baseline --no-dense "$@" >> $OUT 2>/dev/null
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 \
         bench.py --gpus $G --workload $wl --images $img --steps $steps --warmup 3 --no-cpu-baseline --no-dense "$@" >> $OUT 2>/dev/null; fi
}
for s in 0.5 0.7 0.9 0.95 0.975; do run ape_b64 $((512 / G)) 30 --scaling $s; done
for b in 0.01 0.05; do run ape_b64 $((512 / G)) 30 --blur $b; done
for s in 0.5 0.9 0.975; do run dense_b32 $((32 / G > 4 ? 32 / G : 4)) 3 --scaling $s; done
python - <<PY
import json
for l in open("$OUT"):
    if not l.startswith("{"): continue
    d = json.loads(l); c = d["config"]
    print(f'{c["workload"]:10s} G={d["n_gpus"]} img/gpu={c["images_per_gpu"]:4d} scaling={c["scaling"]:<6} blur={c["blur"]:<6} rounds={c["softmin_rounds_per_image"]:4d} '
          f'{d["ms_per_step"]:9.4f} ms/step {d["value"]:12.1f} img/s  fp32 frac {d["roofline"]["frac"]:.3f}  e2e {d["e2e"]["value"]:.0f}')
PY
