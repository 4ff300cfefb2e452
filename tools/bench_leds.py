"""GPU: eds2leds -l 10 LINEAR on genrandomeds-shaped input (BASELINE config 3 shape: 10 % variant sites, 2-4
alternatives, 4 paths, commons {0}); bytes checked against the oracle port at every size, the unmodified reference
library (oracle/_ref/ref_driver, --threads = host cores) timed beside it at the sizes it can finish.
Prints one JSON line per size.   python tools/bench_leds.py [sizes_bp ...]"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import edsparser_b200 as E  # noqa: E402
import oracle_lib  # noqa: E402

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def genrandomeds_like(n_bp, variability=0.10, paths=4, seed=1):
    rng = np.random.default_rng(seed)
    ref = BASES[rng.integers(0, 4, n_bp)]
    n_sites = int(n_bp * variability)
    sites = np.sort(rng.choice(n_bp, n_sites, replace=False))
    eds, seds = [], []
    prev = 0
    refb = ref.tobytes()
    for pos in sites.tolist():
        if pos > prev:
            eds.append(b"{" + refb[prev:pos] + b"}")
            seds.append(b"{0}")
        k = int(rng.integers(2, 5))
        alts = [refb[pos:pos + 1]]
        for _ in range(k - 1):
            u = rng.random()
            if u < 0.7:
                alts.append(bytes(BASES[rng.integers(0, 4, 1)]))
            elif u < 0.85:
                alts.append(refb[pos:pos + 1] + bytes(BASES[rng.integers(0, 4, int(rng.integers(1, 11)))]))
            else:
                alts.append(b"")
        owner = np.concatenate([np.arange(k), rng.integers(0, k, max(0, paths - k))])[:paths]
        eds.append(b"{" + b",".join(alts) + b"}")
        for a in range(k):
            ids = [str(p + 1) for p in np.nonzero(owner == a)[0]]
            seds.append(("{" + ",".join(ids) + "}").encode() if ids else b"{%d}" % (a % paths + 1))
        prev = pos + 1
    if prev < n_bp:
        eds.append(b"{" + refb[prev:] + b"}")
        seds.append(b"{0}")
    return b"".join(eds), b"".join(seds)


def reference_seconds(eds, seds, l, threads):
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(ref):
        return None
    with tempfile.TemporaryDirectory() as d:
        pe, ps = os.path.join(d, "i.eds"), os.path.join(d, "i.seds")
        open(pe, "wb").write(eds)
        open(ps, "wb").write(seds)
        out = subprocess.run([ref, "eds2leds", pe, ps, str(l), os.path.join(d, "o.leds"), os.path.join(d, "o.seds"),
                              str(threads), "1"], check=True, capture_output=True, text=True).stdout
        got = open(os.path.join(d, "o.leds"), "rb").read(), open(os.path.join(d, "o.seds"), "rb").read()
    return float(re.search(r"seconds=([0-9.eE+-]+)", out).group(1)), got


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [20_000, 100_000, 1_000_000]
    ctx = E.load().context(0)
    L = 10
    for n in sizes:
        eds, seds = genrandomeds_like(n)
        exp = oracle_lib.eds2leds(eds, seds, L)
        out = ctx.leds_merge_host(eds, seds, L)  # warm-up
        assert out[:2] == exp, f"parity failure at {n} bp"
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            out = ctx.leds_merge_host(eds, seds, L)
            ts.append(time.perf_counter() - t0)
        ctx.set_profiling(True)
        ctx.leds_merge_host(eds, seds, L)
        kt = ctx.kernel_times()
        ctx.set_profiling(False)
        alg = len(eds) + len(seds) + len(out[0]) + len(out[1])
        line = {"workload": f"genrandomeds-shaped {n} bp, 10% sites, 4 paths, eds2leds -l {L} LINEAR", "bp": n,
                "in_bytes": len(eds) + len(seds), "out_bytes": len(out[0]) + len(out[1]), "rounds": out[2],
                "gpu_host_to_host_ms": round(min(ts) * 1e3, 3), "bp_per_s": n / min(ts), "algorithmic_GBps": alg / min(ts) / 1e9,
                "device_kernel_ms": round(sum(t for _, t in kt), 3), "n_launches": len(kt), "parity": "byte-identical to oracle"}
        if n <= 50_000:
            r = reference_seconds(eds, seds, L, os.cpu_count())
            if r:
                assert r[1] == exp, "reference and oracle disagree"
                line["reference_seconds"] = r[0]
                line["reference_threads"] = os.cpu_count()
                line["speedup_vs_reference"] = r[0] / min(ts)
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
