#!/usr/bin/env python
"""Sweep the launch knobs of k_scan_fused (producer mode, producer warps, ring depth) on a synthetic alignment that
stays resident; one JSON line per setting with the kernel's CUDA-event time and the whole step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edsparser_b200 as E  # noqa: E402


def run(lib, view, rows, env, reps=5):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        c = lib.context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        c.msa_transform_device(view, 10)
        c.msa_transform_device(view, 10)
        c.set_profiling(True)
        acc = {}
        for _ in range(reps):
            c.msa_transform_device(view, 10)
            for n, t in c.kernel_times():
                acc[n] = acc.get(n, 0.0) + t / reps
        return acc
    finally:
        c.close()


def main():
    rows, cols = int(sys.argv[1]), int(sys.argv[2])
    lib = E.load()
    base = lib.context(0)
    view = base.msa_synth(rows, cols, 80, seed=1, variable_ppm=10_000)
    ring = {"EDSB_FUSED": 1, "EDSB_FUSED_L2": 0}  # k_scan_fused (TMA ring), the default
    settings = [{"EDSB_FUSED": 0}, ring]
    what = sys.argv[3:]
    if "none" in what:
        pass
    elif "probe" in what:
        # where the time goes: the same kernel with a phase left out (results are not valid)
        settings += [dict(ring, EDSB_FUSED_PROBE=p) for p in (1, 2, 3, 4)]
    elif "split" in what:
        settings += [dict(ring, EDSB_FUSED_SPLIT=1), dict(ring, EDSB_FUSED_PROBE=3), dict(ring, EDSB_FUSED_SPLIT=1, EDSB_FUSED_PROBE=3),
                     dict(ring, EDSB_FUSED_SPLIT=1, EDSB_FUSED_STAGES=3)]
    elif "direct" in what:
        for nd in (8, 16, 24, 32):
            settings.append(dict(ring, EDSB_FUSED_DIRECT=nd))
        settings.append(dict(ring, EDSB_FUSED_DIRECT=32, EDSB_FUSED_PW=2))
        settings.append(dict(ring, EDSB_FUSED_DIRECT=24, EDSB_FUSED_PROBE=3))
    elif "l2" in what:
        settings += [{"EDSB_FUSED": 1, "EDSB_FUSED_L2": 1, "EDSB_FUSED_PROBE": 2}]
        for cw, dw in ((8, 2), (4, 2), (6, 2), (8, 1), (2, 2)):
            settings.append({"EDSB_FUSED": 1, "EDSB_FUSED_L2": 1, "EDSB_FUSED_CW": cw, "EDSB_FUSED_DW": dw})
    else:
        settings += [dict(ring, EDSB_FUSED_MODE=1), dict(ring, EDSB_FUSED_PW=6), dict(ring, EDSB_FUSED_PAIR=1)]
        # both copy engines at once (mode 2): PW producer warps, the first PWB issue bulk copies for PCT % of the rows
        for pw, pwb, pct in ((4, 2, 50), (6, 4, 70), (8, 6, 70), (8, 5, 60)):
            settings.append(dict(ring, EDSB_FUSED_MODE=2, EDSB_FUSED_PW=pw, EDSB_FUSED_PWB=pwb, EDSB_FUSED_BULK_PCT=pct))
    for env in settings:
        try:
            k = run(lib, view, rows, env)
            scan = k.get("k_scan_l2", k.get("k_scan_fused", k.get("k_scan", 0.0)))
            gbs = rows * cols * 81 / 80 / (scan / 1e3) / 1e9 if scan else 0
            print(json.dumps({"rows": rows, "cols": cols, **env, "scan_ms": round(scan, 4), "scan_GBs": round(gbs, 1),
                              "sum_ms": round(sum(k.values()), 4), "kernels": {a: round(b, 4) for a, b in k.items()}}), flush=True)
        except Exception as err:
            print(json.dumps({"rows": rows, "cols": cols, **env, "error": str(err)[:200]}), flush=True)
    base.close()


if __name__ == "__main__":
    main()
