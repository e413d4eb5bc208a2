#!/usr/bin/env python
"""Diagnostic (run under torchrun, one rank per GPU): per-step device time of the RK3 + dynamic-SFS UJ_fmm step in a LET mode,
with every rank's owned count, the phase walls of each step and the device memory in use — to see whether a slow bench line is
one stalled step (allocation growth, list overflow + restart) or every step.

    python -m torch.distributed.run --nproc-per-node 2 ... tools/let_step_probe.py [particles] [mode] [steps] [balance 0/1]"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

import flowunsteady_b200 as fb
from flowunsteady_b200 import fields
from flowunsteady_b200.dist import ShardedField, partition

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "let"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
balance = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
x, g, s = fields.vortex_rings(n)
P = fb.new_particles(x, fields.floor_gamma(g), s)
lo, hi = partition(n, world)[rank]
sch = fb.default_schemes(kernel="gaussianerf", integration="rungekutta3", relaxation="pedrizzetti", uj="fmm", sfs="dynamic",
                         alpha=0.999, force_positive=1, clippings=1)
eng = fb.Engine(hi - lo, device=lr, schemes=sch)
eng.upload(P[lo:hi].copy())
del P, x, g, s
field = ShardedField(eng, max_local=hi - lo, device=f"cuda:{lr}", fmm=mode)
field.let_balance = balance
ext = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", lr))
rows = []
for k in range(steps):
    timing = k >= steps // 2          # second half: with the (synchronising) phase timer, to see WHERE a slow step spends it
    field.let_timing = {} if timing else None
    eng.synchronize(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(ext)
    field.nextstep(1.0e-3, (0.0, 0.0, 0.0), relax=True)
    e1.record(ext)
    eng.synchronize(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    free_b, total_b = torch.cuda.mem_get_info(lr)
    L = field._let or {}
    rows.append({"step": k, "rank": rank, "dev_ms": round(e0.elapsed_time(e1), 1), "wall_ms": round(wall, 1),
                 "n_own": int(sum(L.get("recv", []))), "mem_gb": round((total_b - free_b) / 1e9, 2),
                 "bufs_mb": round(sum(t.numel() * t.element_size() for t in field._bufs.values()) / 1e6),
                 "tree": eng.fmm_stats(), "phases": {k2[:1]: round(v, 1) for k2, v in (field.let_timing or {}).items()}})
    field.let_timing = None
allr = [None] * world
dist.all_gather_object(allr, rows)
if rank == 0:
    for k in range(steps):
        for r in range(world):
            print(json.dumps(allr[r][k]))
dist.barrier()
dist.destroy_process_group()
