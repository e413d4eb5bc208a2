#!/bin/bash
# Round-2 GPU pass Y (2 GPUs): per-step probe of both LET modes after the cell-array headroom fix.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 60 $TR --master-port 29531 tools/let_step_probe.py 5000000 let 10 1 > gpurun_out/y_probe_let.log 2> gpurun_out/y_probe_let.err
timeout 60 $TR --master-port 29532 tools/let_step_probe.py 5000000 let_halo 10 1 > gpurun_out/y_probe_halo.log 2> gpurun_out/y_probe_halo.err
python - <<'PY'
import json
for f in ("y_probe_let", "y_probe_halo"):
    rows = [json.loads(l) for l in open(f"gpurun_out/{f}.log") if l.startswith("{")]
    print(f, [r["dev_ms"] for r in rows if r["rank"] == 0], "phase3 r0", [r["phases"].get("3") for r in rows if r["rank"] == 0],
          "r1", [r["phases"].get("3") for r in rows if r["rank"] == 1], "mem", rows[-1]["mem_gb"])
PY
tail -2 gpurun_out/y_probe_halo.err
