#!/usr/bin/env python
"""Workload sweep over the BASELINE.json configurations that bench.py does not headline:

  configs[0]  examples/wing stand-in      N = 10,100 / 20,200 / 40,401   direct FP64, RK3 + pedrizzetti
  configs[1]  examples/rotorhover stand-in N = 7e4 / 2e5 / 1e6            RK3 + dynamic SFS + pedrizzetti, direct and UJ_fmm
  configs[3]  examples/vahana stand-in     N = 5e6                         UJ_fmm p=4 ncrit=50 theta=0.4, dynamic SFS + controls
  configs[4]  random field sweep           N = 1e5 .. 5e7                  direct (while it stays under ~3 s) vs UJ_fmm

One JSON object per line goes to stdout and to gpurun_out/sweep.jsonl as soon as a case finishes.  Times are CUDA-event
times on the engine's stream (median of `reps` after one warm-up); errors are relative L2 / max-norm against the direct
kernel on 2048 sampled particles (vpmb200_uj_probe).  One GPU; the multi-GPU figures come from bench.py under torchrun.

    python tools/sweep.py [--cases wing,rotor,vahana,random] [--max-n 20000000] [--reps 3]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import flowunsteady_b200 as fb  # noqa: E402
from flowunsteady_b200 import engine as E, fields  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "sweep.jsonl")


def emit(rec):
    line = json.dumps(rec)
    print(line, flush=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def timed_ms(eng, fn, reps):
    """Median device time of fn() over `reps` calls (CUDA events on the engine's stream), after one warm-up call."""
    import torch
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", 0))
    fn()
    eng.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        fn()
        e1.record(ext)
        eng.synchronize()
        out.append(e0.elapsed_time(e1))
    return statistics.median(out)


def err_vs_direct(eng, x, n, nsamp=2048):
    """Relative error of the resident U, J against the direct kernel on sampled particles."""
    idx = np.random.default_rng(1234).choice(n, min(nsamp, n), replace=False)
    Ud, Jd = eng.uj_probe(np.ascontiguousarray(x[idx]), want_J=True)
    ptr_u = np.zeros((n, 43)) if n <= 2_000_000 else None
    if ptr_u is not None:
        eng.download(ptr_u, field_mask=E.FM_U | E.FM_J)
        Uf, Jf = ptr_u[idx, 9:12], ptr_u[idx, 15:24]
    else:   # large N: read the sampled rows straight from the device SoA (avoids a 17 GB host matrix)
        import torch
        p, ld = eng.device_field(0)

        class _A:
            __cuda_array_interface__ = {"shape": (43, ld), "typestr": "<f8", "data": (int(p), False), "version": 3}
        st = torch.as_tensor(_A(), device="cuda:0")
        ti = torch.as_tensor(idx, device="cuda:0")
        Uf = st[9:12][:, ti].T.cpu().numpy()
        Jf = st[15:24][:, ti].T.cpu().numpy()
    def l2(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
    def mx(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
    return {"U_l2": l2(Uf, Ud), "U_max": mx(Uf, Ud), "J_l2": l2(Jf, Jd), "J_max": mx(Jf, Jd), "samples": int(idx.size)}


def upload_state(eng, x, g, s):
    """Upload X, Gamma, sigma without materialising the full 43-column host matrix for very large N."""
    n = x.shape[0]
    if n <= 2_000_000:
        eng.upload(fb.new_particles(x, g, s))
        return
    chunk = 2_000_000
    first = True
    for lo in range(0, n, chunk):
        P = fb.new_particles(x[lo:lo + chunk], g[lo:lo + chunk], s[lo:lo + chunk])
        if first:
            eng.upload(P)
            first = False
        else:
            eng.add_particles(P)


DYN = dict(sfs="dynamic", alpha=0.999, force_positive=1, clippings=1)   # SFS_Cd_twolevel_nobackscatter (rotorhover.jl:53-55)


def case_wing(args):
    for rows in (100, 200, 400):
        x, g, s = fields.wing_wake(rows=rows)
        n = x.shape[0]
        if rows == 400:
            x, g, s = x[:40401], g[:40401], s[:40401]
            n = 40401
        with fb.Engine(n, schemes=fb.default_schemes(uj="direct", integration="rungekutta3", relaxation="pedrizzetti")) as eng:
            upload_state(eng, x, g, s)
            ms_uj = timed_ms(eng, lambda: eng.uj(True, False, False), args.reps)
            ms_step = timed_ms(eng, lambda: eng.nextstep(1e-4, (49.7, 0.0, 0.0), True), args.reps)
        emit({"config": "wing", "field": "wing_wake", "particles": n, "uj": "direct", "sfs": "none",
              "ms_per_evaluation": ms_uj, "ms_per_step": ms_step, "interactions_per_s": n * n / (ms_uj * 1e-3)})


def case_rotor(args):
    for n_req, nfil, spr in ((70_000, 41, 36), (200_000, 101, 72), (1_000_000, 101, 360)):
        x, g, s = fields.rotor_wake(n_req, nfil=nfil, nsteps_per_rev=spr, p_per_step=4 if spr == 36 else 2)
        n = x.shape[0]
        for uj in ("direct", "fmm"):
            if uj == "direct" and n > 300_000:
                continue
            sch = fb.default_schemes(uj=uj, integration="rungekutta3", relaxation="pedrizzetti", **DYN)
            with fb.Engine(n, schemes=sch) as eng:
                upload_state(eng, x, g, s)
                ms_uj = timed_ms(eng, lambda: eng.uj(True, True, True), args.reps)
                err = err_vs_direct(eng, x, n) if uj == "fmm" else None
                ms_step = timed_ms(eng, lambda: eng.nextstep(1e-5, (0.0, 0.0, 0.0), True), args.reps)
                stats = eng.fmm_stats() if uj == "fmm" else None
                bad = eng.count_nonfinite()
            emit({"config": "rotorhover", "field": "rotor_wake", "particles": n, "sigma": float(s[0]), "uj": uj,
                  "sfs": "dynamic (pseudo3level_positive, alpha=0.999, clipping_backscatter)",
                  "ms_per_evaluation_with_estr": ms_uj, "ms_per_step": ms_step, "evaluations_per_step": 5,
                  "err_vs_direct": err, "fmm_tree": stats, "nonfinite": bad})


def case_vahana(args):
    n_req = min(5_000_000, args.max_n)
    x, g, s = fields.vahana_wake(n_req)
    n = x.shape[0]
    # vahana.jl:147-150: DynamicSFS(Estr_fmm, pseudo3level_positive; alpha=0.999, clippings, controls=(directional, magnitude))
    sch = fb.default_schemes(uj="fmm", integration="rungekutta3", relaxation="pedrizzetti", controls=3, **DYN)
    with fb.Engine(n, schemes=sch) as eng:
        upload_state(eng, x, g, s)
        ms_uj = timed_ms(eng, lambda: eng.uj(True, True, False), args.reps)
        ms_ujs = timed_ms(eng, lambda: eng.uj(True, True, True), args.reps)
        err = err_vs_direct(eng, x, n)
        ms_step = timed_ms(eng, lambda: eng.nextstep(30.0 / 21600, (0.0, 0.0, 0.0), True), args.reps)
        stats = eng.fmm_stats()
        bad = eng.count_nonfinite()
    emit({"config": "vahana", "field": "vahana_wake", "particles": n, "sigma": float(s[0]), "uj": "fmm p=4 ncrit=50 theta=0.4",
          "sfs": "dynamic + control_directional + control_magnitude", "ms_per_evaluation": ms_uj,
          "ms_per_evaluation_with_estr": ms_ujs, "ms_per_step": ms_step, "evaluations_per_step": 5, "err_vs_direct": err,
          "fmm_tree": stats, "nonfinite": bad})


def case_random(args):
    for n in (100_000, 200_000, 500_000, 1_000_000, 2_000_000, 5_000_000, 10_000_000, 20_000_000, 50_000_000):
        if n > args.max_n:
            break
        t0 = time.perf_counter()
        x, g, s = fields.random_field(n)
        rec = {"config": "random", "field": "random_field", "particles": n, "sigma": float(s[0])}
        with fb.Engine(n, schemes=fb.default_schemes(uj="fmm")) as eng:
            upload_state(eng, x, g, s)
            rec["fmm_ms_per_evaluation"] = timed_ms(eng, lambda: eng.uj(True, False, False), args.reps)
            rec["fmm_err_vs_direct"] = err_vs_direct(eng, x, n)
            rec["fmm_tree"] = eng.fmm_stats()
            if n <= 1_000_000:
                eng.set_schemes(fb.default_schemes(uj="direct"))
                ms = timed_ms(eng, lambda: eng.uj(True, False, False), 1 if n >= 1_000_000 else args.reps)
                rec["direct_ms_per_evaluation"] = ms
                rec["direct_interactions_per_s"] = float(n) * n / (ms * 1e-3)
        rec["wall_s_total"] = time.perf_counter() - t0
        emit(rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="wing,rotor,vahana,random")
    ap.add_argument("--max-n", type=int, default=20_000_000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    for c in args.cases.split(","):
        {"wing": case_wing, "rotor": case_rotor, "vahana": case_vahana, "random": case_random}[c](args)


if __name__ == "__main__":
    main()
