"""GPU: sweep k_scan's grid (blocks per SM) on the config-2 window; prints per-kernel CUDA-event times."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edsparser_b200 as E

R, C = int(os.environ.get("R", 100)), int(os.environ.get("C", 10_000_000))
lib = E.Library(os.environ["EDSB_LIB"]) if os.environ.get("EDSB_LIB") else E.load()
ctx = lib.context(0)
view = ctx.msa_synth(R, C, 80, seed=1, variable_ppm=10_000)
ctx.set_profiling(True)
for bps in [int(x) for x in os.environ.get("BPS", "2,4,6,8,12,16").split(",")]:
    ctx.set_tuning(0, bps)
    acc = {}
    n = 5
    for i in range(n + 2):
        ctx.msa_transform_device(view, 10)
        if i >= 2:
            for k, t in ctx.kernel_times():
                acc[k] = acc.get(k, 0) + t / n
    tot = sum(acc.values())
    import time
    ctx.set_profiling(False)
    for _ in range(3):
        ctx.msa_transform_device(view, 10)
    t0 = time.perf_counter()
    for _ in range(20):
        e, s_, st = ctx.msa_transform_device(view, 10)
    step_ms = (time.perf_counter() - t0) / 20 * 1e3
    ctx.set_profiling(True)
    print(json.dumps({"blocks_per_sm": bps, "k_scan_ms": round(acc["k_scan"], 4),
                      "scan_GBps": round(R * C * 81 / 80 / acc["k_scan"] / 1e6, 1), "sum_ms": round(tot, 4), "step_ms": round(step_ms, 4), "out_MB": round((e.bytes + s_.bytes) / 1e6, 1),
                      "kernels": {k: round(v, 4) for k, v in acc.items()}}))
