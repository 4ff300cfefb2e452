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
    """genrandomeds-shaped EDS + SEDS, fully vectorised."""
    rng = np.random.default_rng(seed)
    ref = BASES[rng.integers(0, 4, n_bp)]
    n_sites = int(n_bp * variability)
    sites = np.sort(rng.choice(n_bp, n_sites, replace=False)).astype(np.int64)
    k = rng.integers(2, 5, n_sites)                       # alternatives per site
    A = int(k.sum())
    site_of = np.repeat(np.arange(n_sites), k)
    first = np.cumsum(k) - k                              # flat index of each site's alt 0
    j = np.arange(A) - first[site_of]                     # alt index within its site
    u = rng.random(A)
    ins_extra = rng.integers(1, 11, A)
    alen = np.where(j == 0, 1, np.where(u < 0.7, 1, np.where(u < 0.85, 1 + ins_extra, 0))).astype(np.int64)
    # --- EDS layout: [common_s][site_s] ... [tail common]
    prev = np.concatenate(([0], sites[:-1] + 1))
    clen = sites - prev
    cbytes = np.where(clen > 0, clen + 2, 0)
    sum_alen = np.add.reduceat(alen, first)
    sbytes = 2 + sum_alen + (k - 1)
    tail_len = n_bp - (sites[-1] + 1) if n_sites else n_bp
    entry = cbytes + sbytes
    off_c = np.cumsum(entry) - entry                      # start of common_s
    off_s = off_c + cbytes                                # start of site_s
    total = int(entry.sum() + (tail_len + 2 if tail_len > 0 else 0))
    out = np.empty(total, dtype=np.uint8)
    # commons: every non-site reference position
    is_site = np.zeros(n_bp, dtype=bool); is_site[sites] = True
    seg = np.cumsum(is_site) - is_site                    # number of sites strictly before q  (segment index)
    q = np.nonzero(~is_site)[0]
    sq = seg[q]
    seg_start_off = np.concatenate((off_c, [int(entry.sum())]))   # segment s starts at off_c[s]; tail at end
    seg_prev = np.concatenate((prev, [sites[-1] + 1 if n_sites else 0]))
    out[seg_start_off[sq] + 1 + (q - seg_prev[sq])] = ref[q]
    has_c = clen > 0
    out[off_c[has_c]] = ord("{"); out[off_c[has_c] + 1 + clen[has_c]] = ord("}")
    if tail_len > 0:
        out[int(entry.sum())] = ord("{"); out[total - 1] = ord("}")
    # sites
    out[off_s] = ord("{"); out[off_s + sbytes - 1] = ord("}")
    # alt offsets within site: 1 + cumulative (alen + 1 separator)
    step = alen + 1
    cs = np.cumsum(step) - step
    a_off = off_s[site_of] + 1 + (cs - cs[first][site_of])
    # separators after every alt except the last of a site
    last = (j == k[site_of] - 1)
    out[(a_off + alen)[~last]] = ord(",")
    # alt characters: first char = ref base for alt 0 and insertions, random base for SNPs; insertion tails random
    has1 = alen >= 1
    firstch = np.where((j == 0) | (u >= 0.7), ref[sites[site_of]], BASES[rng.integers(0, 4, A)])
    out[a_off[has1]] = firstch[has1]
    extra = np.where(alen > 1, alen - 1, 0)
    E = int(extra.sum())
    if E:
        src = np.repeat(np.arange(A), extra)
        within = np.arange(E) - np.repeat(np.cumsum(extra) - extra, extra)
        out[a_off[src] + 1 + within] = BASES[rng.integers(0, 4, E)]
    # --- SEDS: commons "{0}", alternative a of a site: the paths that own it (path p < k owns alt p, the rest random)
    owner = rng.integers(0, 1 << 30, (n_sites, paths)) % k[:, None]
    owner[:, :2] = np.arange(2)[None, :]
    for p in range(2, paths):
        owner[:, p] = np.where(k > p, p, owner[:, p])
    # per alt: mask of paths
    masks = np.zeros(A, dtype=np.int64)
    for p in range(paths):
        np.add.at(masks, first + owner[:, p], 1 << p)
    # text per mask via lookup table (paths <= 8)
    texts = []
    for m in range(1 << paths):
        ids = [str(p + 1) for p in range(paths) if m >> p & 1]
        texts.append(("{" + ",".join(ids) + "}").encode())
    tlen = np.array([len(t) for t in texts], dtype=np.int64)
    maxl = int(tlen.max())
    tab = np.zeros((1 << paths, maxl), dtype=np.uint8)
    for m, t in enumerate(texts):
        tab[m, :len(t)] = np.frombuffer(t, dtype=np.uint8)
    s_alt = tlen[masks]
    s_site = np.add.reduceat(s_alt, first)
    s_entry = np.where(has_c, 3, 0) + s_site
    s_off_c = np.cumsum(s_entry) - s_entry
    s_total = int(s_entry.sum() + (3 if tail_len > 0 else 0))
    sout = np.empty(s_total, dtype=np.uint8)
    c3 = np.frombuffer(b"{0}", dtype=np.uint8)
    for d in range(3):
        sout[s_off_c[has_c] + d] = c3[d]
        if tail_len > 0:
            sout[s_total - 3 + d] = c3[d]
    s_cs = np.cumsum(s_alt) - s_alt
    sa_off = (s_off_c + np.where(has_c, 3, 0))[site_of] + (s_cs - s_cs[first][site_of])
    for d in range(maxl):
        sel = s_alt > d
        sout[sa_off[sel] + d] = tab[masks[sel], d]
    return out.tobytes(), sout.tobytes()


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


def timed(ctx, eds, seds, l, max_out=0, reps=3):
    out = ctx.leds_merge_host(eds, seds, l, max_output_bytes=max_out)  # warm-up: buffers grow once
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = ctx.leds_merge_host(eds, seds, l, max_output_bytes=max_out)
        ts.append(time.perf_counter() - t0)
    ctx.set_profiling(True)
    ctx.leds_merge_host(eds, seds, l, max_output_bytes=max_out)
    kt = ctx.kernel_times()
    ctx.set_profiling(False)
    agg = {}
    for name, t in kt:
        agg[name] = agg.get(name, 0.0) + t
    top = {k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]}
    return out, min(ts), sum(t for _, t in kt), len(kt), top


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [20_000, 100_000, 1_000_000]
    check_up_to = int(os.environ.get("EDSB_LEDS_CHECK_BP", "20000000"))  # the oracle port is linear: ~0.5 s per Mbp
    ctx = E.load().context(0)
    if os.environ.get("EDSB_PARTITIONS"):  # A/B of the scan partition count (default: 4 per SM)
        ctx.set_tuning(int(os.environ["EDSB_PARTITIONS"]), 0)
    L = 10
    for n in sizes:
        eds, seds = genrandomeds_like(n)
        out, best, dev_ms, n_launch, top = timed(ctx, eds, seds, L)
        parity = "not checked at this size (EDSB_LEDS_CHECK_BP)"
        if n <= check_up_to:
            exp = oracle_lib.eds2leds(eds, seds, L)
            assert out[:2] == exp, f"parity failure at {n} bp"
            parity = "byte-identical to oracle"
        else:  # size-independent properties: the result is an l-EDS and merging it again changes nothing
            assert ctx.is_leds(out[0], L)
            again = ctx.leds_merge_host(out[0], out[1], L)
            assert again[:2] == out[:2] and again[2] == 0
            parity = "is_leds holds and the merge is idempotent on its own output"
        alg = len(eds) + len(seds) + len(out[0]) + len(out[1])
        # the C ABI alone: pinned input, results as pinned views (no Python copies inside the timed region)
        import torch

        pe = torch.empty(len(eds), dtype=torch.uint8).pin_memory()
        pe.copy_(torch.frombuffer(bytearray(eds), dtype=torch.uint8))
        ps = torch.empty(len(seds), dtype=torch.uint8).pin_memory()
        ps.copy_(torch.frombuffer(bytearray(seds), dtype=torch.uint8))
        assert ctx.leds_merge_host_view(eds, seds, L)[:2] == out[:2]
        ctx.leds_merge_host_view_raw(pe.data_ptr(), len(eds), ps.data_ptr(), len(seds), L)
        tv = []
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.leds_merge_host_view_raw(pe.data_ptr(), len(eds), ps.data_ptr(), len(seds), L)
            tv.append(time.perf_counter() - t0)
        del pe, ps
        line = {"workload": f"config 3 shape: genrandomeds-like {n} bp, 10% sites, 4 paths, eds2leds -l {L} LINEAR", "bp": n,
                "in_bytes": len(eds) + len(seds), "out_bytes": len(out[0]) + len(out[1]), "rounds": out[2],
                "gpu_host_to_host_ms": round(best * 1e3, 3), "c_abi_pinned_view_ms": round(min(tv) * 1e3, 3), "bp_per_s": n / best, "algorithmic_GBps": alg / best / 1e9,
                "device_kernel_ms": round(dev_ms, 3), "device_algorithmic_GBps": round(alg / dev_ms / 1e6, 2),
                "n_launches": n_launch, "top_kernels_ms": top, "parity": parity}
        if n <= 50_000:
            r = reference_seconds(eds, seds, L, os.cpu_count())
            if r:
                assert r[1] == oracle_lib.eds2leds(eds, seds, L), "reference and oracle disagree"
                line["reference_seconds"] = r[0]
                line["reference_threads"] = os.cpu_count()
                line["speedup_vs_reference"] = r[0] / best
        print(json.dumps(line), flush=True)
        # LINEAR vs CARTESIAN on the same generator at variability 0.01 (CARTESIAN cannot run at 0.10: SURVEY 6)
        eds1, seds1 = genrandomeds_like(n, variability=0.01, seed=2)
        lin, t_lin, d_lin, _, _ = timed(ctx, eds1, seds1, L)
        budget = 8 << 30
        try:
            car, t_car, d_car, _, _ = timed(ctx, eds1, None, L, max_out=budget)
            cart = {"host_to_host_ms": round(t_car * 1e3, 3), "device_kernel_ms": round(d_car, 3), "out_bytes": len(car[0]), "rounds": car[2]}
            if n <= check_up_to and len(car[0]) < (1 << 28):
                assert car[0] == oracle_lib.eds2leds(eds1, None, L, max_out_bytes=1 << 30)[0]
                cart["parity"] = "byte-identical to oracle"
        except E.EdsError as err:
            cart = {"refused": err.message}
        try:
            ctx.leds_merge_host(eds, None, L, max_output_bytes=budget)
            refused = "NOT refused"
        except E.EdsError as err:
            refused = f"status {err.status}: {err.message}"
        print(json.dumps({"workload": f"genrandomeds-like {n} bp, 1% sites: LINEAR vs CARTESIAN, -l {L}", "bp": n,
                          "in_bytes": len(eds1) + len(seds1),
                          "linear": {"host_to_host_ms": round(t_lin * 1e3, 3), "device_kernel_ms": round(d_lin, 3),
                                     "out_bytes": len(lin[0]) + len(lin[1]), "rounds": lin[2]},
                          "cartesian": cart, "cartesian_on_the_10pct_input_with_8GiB_budget": refused}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
