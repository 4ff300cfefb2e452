"""GPU box: FILE-TO-FILE wall time of the three tools (timed scope ii): process start, CUDA context, read,
H2D, kernels, D2H, write — next to the unmodified reference library making the same calls on the same files
(oracle/_ref/ref_driver; msa2eds / vcf2eds are single-threaded by construction, eds2leds gets every host core) at the
sizes it can finish, each labelled with its size. Output bytes are compared where both ran. One JSON line per tool.
    python tools/bench_cli.py [--gpus N] [--dir /dev/shm]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import edsparser_b200 as E  # noqa: E402

BIN = os.path.join(ROOT, "edsparser_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def run_tool(args):
    t0 = time.perf_counter()
    r = subprocess.run(args, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("%s failed: %s" % (args[0], r.stderr[-400:]))
    m = re.search(r"Runtime: ([0-9.]+) ?(ms|s)", r.stderr)
    own = None
    if m:
        own = float(m.group(1)) * (1e-3 if m.group(2) == "ms" else 1.0)
    return dt, own


def run_ref(args):
    if not os.path.exists(REF):
        return None, None
    t0 = time.perf_counter()
    r = subprocess.run([REF] + args, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    m = re.search(r"seconds=([0-9.eE+-]+)", r.stdout)
    return dt, (float(m.group(1)) if m else None)


def same(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ap.add_argument("--msa-cols", type=int, default=10_000_000)
    ap.add_argument("--leds-bp", type=int, default=100_000_000)
    ap.add_argument("--vcf-sites", type=int, default=100_000)
    ap.add_argument("--only", default="", help="msa | leds | vcf")
    a = ap.parse_args()
    gp = ["--gpus", str(a.gpus)] if a.gpus > 1 else []
    ctx = E.load().context(0)
    cores = os.cpu_count()
    with tempfile.TemporaryDirectory(dir=a.dir) as d:
        if a.only in ("", "msa"):
            bench_msa(a, ctx, d, gp)
        if a.only in ("", "leds"):
            bench_leds(a, ctx, d, gp, cores)
        if a.only in ("", "vcf"):
            bench_vcf(a, ctx, d, gp)
    ctx.close()


def clear(d):
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))


def bench_msa(a, ctx, d, gp):
    if True:
        # ---- msa2eds -l 10: BASELINE config 2 (100 x 10 Mbp); the reference runs the same file in full
        v = ctx.msa_synth(100, a.msa_cols, 80, seed=1, variable_ppm=10_000)
        text = ctx.download(E.Buffer(v.text, v.text_bytes))
        ctx.msa_synth_free()
        msa = os.path.join(d, "c2.msa")
        open(msa, "wb").write(text)
        n_bytes = len(text)
        del text
        run_tool([os.path.join(BIN, "msa2eds"), "-i", msa, "-l", "10", "-o", os.path.join(d, "warm.leds"), "-s", os.path.join(d, "warm.seds")])  # page cache, driver
        wall, own = run_tool([os.path.join(BIN, "msa2eds"), "-i", msa, "-l", "10", "-o", os.path.join(d, "a.leds"), "-s", os.path.join(d, "a.seds")] + gp)
        rw, rs = run_ref(["msa2eds", msa, "10", os.path.join(d, "r.leds"), os.path.join(d, "r.seds")])
        line = {"tool": "msa2eds -l 10", "workload": "config 2: 100 seq x %d columns, %d bytes of .msa" % (a.msa_cols, n_bytes),
                "gpus": a.gpus, "wall_s": round(wall, 3), "tool_reported_s": own,
                "reference": {"wall_s": rw and round(rw, 3), "library_call_s": rs, "threads": 1, "same_file": True},
                "byte_equal": bool(rw) and same(os.path.join(d, "a.leds"), os.path.join(d, "r.leds")) and
                same(os.path.join(d, "a.seds"), os.path.join(d, "r.seds"))}
        print(json.dumps(line), flush=True)
        clear(d)


def bench_leds(a, ctx, d, gp, cores):
    if True:
        # ---- eds2leds -l 10 LINEAR: config 3 shape; the reference at 50 kbp (it is quadratic: SURVEY 3.2)
        def gen(bp, name):
            e, s = ctx.genrandomeds_device(bp, 100_000, 4, 1)
            pe, ps = os.path.join(d, name + ".eds"), os.path.join(d, name + ".seds")
            open(pe, "wb").write(ctx.download(e))
            open(ps, "wb").write(ctx.download(s))
            return pe, ps

        pe, ps = gen(a.leds_bp, "c3")
        wall, own = run_tool([os.path.join(BIN, "eds2leds"), "-i", pe, "-s", ps, "-l", "10", "-o", os.path.join(d, "a.leds")] + gp)
        se, ss = gen(50_000, "small")
        w2, _ = run_tool([os.path.join(BIN, "eds2leds"), "-i", se, "-s", ss, "-l", "10", "-o", os.path.join(d, "b.leds")])
        rw, rs = run_ref(["eds2leds", se, ss, "10", os.path.join(d, "r.leds"), os.path.join(d, "r.seds"), str(cores), "1"])
        line = {"tool": "eds2leds -l 10 (LINEAR)", "workload": "config 3 shape: %d bp, 10 %% sites, 4 paths; %d + %d bytes in" % (
            a.leds_bp, os.path.getsize(pe), os.path.getsize(ps)), "gpus": a.gpus, "wall_s": round(wall, 3), "tool_reported_s": own,
            "at_50kbp": {"wall_s": round(w2, 3), "reference_wall_s": rw and round(rw, 3), "reference_library_call_s": rs,
                         "reference_threads": cores,
                         "byte_equal": bool(rw) and same(os.path.join(d, "b.leds"), os.path.join(d, "r.leds")) and
                         same(os.path.join(d, "b.seds"), os.path.join(d, "r.seds"))}}
        print(json.dumps(line), flush=True)
        clear(d)


def bench_vcf(a, ctx, d, gp):
    if True:
        # ---- vcf2eds -l 10: config 5 shape at --vcf-sites; the reference at 2000 sites
        import vcf_checks

        def genv(sites, name):
            vcf, fa = vcf_checks.synth_vcf(n_bases=100 * sites, n_sites=sites, n_samples=2504, seed=1)
            pv, pf = os.path.join(d, name + ".vcf"), os.path.join(d, name + ".fa")
            open(pv, "wb").write(vcf)
            open(pf, "wb").write(fa)
            return pv, pf, len(vcf)

        pv, pf, nb = genv(a.vcf_sites, "c5")
        wall, own = run_tool([os.path.join(BIN, "vcf2eds"), "-i", pv, "-r", pf, "-l", "10", "-o", os.path.join(d, "a.leds"), "-s", os.path.join(d, "a.seds")] + gp)
        sv, sf, _ = genv(2000, "small")
        w2, _ = run_tool([os.path.join(BIN, "vcf2eds"), "-i", sv, "-r", sf, "-l", "10", "-o", os.path.join(d, "b.leds"), "-s", os.path.join(d, "b.seds")])
        rw, rs = run_ref(["vcf2eds", sv, sf, "10", os.path.join(d, "r.leds"), os.path.join(d, "r.seds")])
        line = {"tool": "vcf2eds -l 10", "workload": "config 5 shape: %d sites x 2504 samples, %d bytes of VCF" % (a.vcf_sites, nb),
                "gpus": a.gpus, "wall_s": round(wall, 3), "tool_reported_s": own,
                "at_2000_sites": {"wall_s": round(w2, 3), "reference_wall_s": rw and round(rw, 3), "reference_library_call_s": rs,
                                  "reference_threads": 1,
                                  "byte_equal": bool(rw) and same(os.path.join(d, "b.leds"), os.path.join(d, "r.leds")) and
                                  same(os.path.join(d, "b.seds"), os.path.join(d, "r.seds"))}}
        print(json.dumps(line), flush=True)
        clear(d)


if __name__ == "__main__":
    main()
