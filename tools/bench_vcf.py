"""GPU: vcf2eds on BASELINE config-5-shaped input (tests/vcf_checks.synth_vcf: random reference, SNP/indel sites, 1 %
overlapping companions, 2504 phased diploid samples). Bytes are checked against the oracle port at the smallest size;
the unmodified reference library (oracle/_ref/ref_driver vcf2eds, single-threaded by construction) is timed beside it
on a bounded sample. Prints one JSON line per size.
    python tools/bench_vcf.py [n_sites ...]   (reference length = 100 x n_sites)"""
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
import oracle_lib  # noqa: E402
import vcf_checks  # noqa: E402

N_SAMPLES = int(os.environ.get("EDSB_VCF_SAMPLES", "2504"))


def reference_seconds(vcf, fa, l):
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(ref):
        return None
    with tempfile.TemporaryDirectory() as d:
        pv, pf = os.path.join(d, "i.vcf"), os.path.join(d, "r.fa")
        open(pv, "wb").write(vcf)
        open(pf, "wb").write(fa)
        r = subprocess.run([ref, "vcf2eds", pv, pf, str(l), os.path.join(d, "o.eds"), os.path.join(d, "o.seds")],
                           capture_output=True, text=True)
        m = re.search(r"seconds=([0-9.eE+-]+)", r.stdout)
        return float(m.group(1)) if m else None


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [2000, 100000]
    ctx = E.load().context(0)
    for i, n_sites in enumerate(sizes):
        need_gb = 3.2e-5 * n_sites * N_SAMPLES / 2504 * 1000 / 1e3  # ~32 GB of host memory per 10^6 sites while generating
        avail_gb = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) / 1e6
        if need_gb > 0.6 * avail_gb:
            print(json.dumps({"skipped": n_sites, "why": "host memory: need ~%.0f GB, %.0f GB available" % (need_gb, avail_gb)}), flush=True)
            continue
        t0 = time.time()
        vcf, fa = vcf_checks.synth_vcf(n_bases=100 * n_sites, n_sites=n_sites, n_samples=N_SAMPLES, seed=1)
        gen_s = time.time() - t0
        line = {"workload": "config 5 shape: %d bp reference, %d sites x %d samples, vcf2eds" % (100 * n_sites, n_sites, N_SAMPLES),
                "vcf_bytes": len(vcf), "fasta_bytes": len(fa), "generator_s": round(gen_s, 2)}
        checked = None
        if len(vcf) <= 64 << 20:
            exp = oracle_lib.vcf2eds(vcf, fa, 0)
            got = ctx.vcf_transform_host(vcf, fa, 0)
            assert got[:2] == exp[:2], "vcf2eds differs from the oracle"
            exp10 = oracle_lib.vcf2eds(vcf, fa, 10)
            got10 = ctx.vcf_transform_host(vcf, fa, 10)
            assert got10[:2] == exp10[:2], "vcf2eds -l 10 differs from the oracle"
            checked = "byte-equal to the oracle at l = 0 and l = 10"
            t = reference_seconds(vcf, fa, 0)
            t10 = reference_seconds(vcf, fa, 10)
            line["reference_cpu"] = {"l0_s": t, "l10_s": t10, "threads": 1, "kind": "reference (oracle/_ref, unmodified)"}
        line["checked"] = checked
        dv, df = ctx.upload(vcf), ctx.upload(fa)
        try:
            for _ in range(2):
                ctx.vcf_transform_device(dv, df)
            best = None
            for _ in range(5):
                t0 = time.perf_counter()
                e, s, st = ctx.vcf_transform_device(dv, df)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            ctx.set_profiling(True)
            ctx.vcf_transform_device(dv, df)
            times = ctx.kernel_times()
            ctx.set_profiling(False)
            algo = len(vcf) + len(fa) + st["eds_bytes"] + st["seds_bytes"]
            agg = {}
            for name, ms in times:
                agg[name] = round(agg.get(name, 0.0) + ms, 4)
            line.update({"device_resident_ms": round(best * 1e3, 3), "algorithmic_bytes": algo,
                         "gb_per_s": round(algo / best / 1e9, 1), "stats": st, "kernels_ms": agg})
        finally:
            ctx.device_free(dv)
            ctx.device_free(df)
        # host to host through the C ABI (H2D from pinned memory + kernels + D2H into malloc'd strings), l = 0 and 10
        if True:
            import torch

            pv = torch.empty(len(vcf), dtype=torch.uint8).pin_memory()
            pv.copy_(torch.frombuffer(bytearray(vcf), dtype=torch.uint8))
            pf = torch.empty(len(fa), dtype=torch.uint8).pin_memory()
            pf.copy_(torch.frombuffer(bytearray(fa), dtype=torch.uint8))
            big = len(vcf) > (4 << 30)
            for l in (0, 10):
                if not big:  # the malloc'd form: a fresh gigabyte-sized result per call
                    ctx.vcf_transform_host_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), l)
                    t0 = time.perf_counter()
                    ctx.vcf_transform_host_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), l)
                    line["host_to_host_l%d_ms" % l] = round((time.perf_counter() - t0) * 1e3, 2)
                sizes_out, st_l = ctx.vcf_transform_host_view_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), l)
                line["out_bytes_l%d" % l] = list(sizes_out)
                if l:
                    line["leds_rounds"] = st_l.get("leds_rounds")
                ts = []
                for _ in range(1 if big else 3):
                    t0 = time.perf_counter()
                    ctx.vcf_transform_host_view_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), l)
                    ts.append(time.perf_counter() - t0)
                line["host_to_host_view_l%d_ms" % l] = round(min(ts) * 1e3, 2)
            n_gpus = int(os.environ.get("EDSB_VCF_GPUS", "0"))
            if n_gpus > 1:
                # slices of the record lines over n_gpus devices (eds_group_vcf_transform_host): same bytes, host to host
                import hashlib

                g = E.load().group(list(range(n_gpus)))
                try:
                    for l in (0, 10):
                        one = ctx.vcf_transform_host_view(pv, pf, l)
                        sz, st_g, used, ge, gs = g.vcf_transform_host_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), l, keep=True)
                        same = hashlib.sha256(ge).digest() == hashlib.sha256(one[0]).digest() and hashlib.sha256(gs).digest() == hashlib.sha256(one[1]).digest()
                        del one, ge, gs
                        ts, tv = [], []
                        g.vcf_transform_host_view_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), l)
                        for _ in range(1 if big else 3):
                            t0 = time.perf_counter()
                            g.vcf_transform_host_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), l)
                            ts.append(time.perf_counter() - t0)
                            t0 = time.perf_counter()
                            g.vcf_transform_host_view_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), l)
                            tv.append(time.perf_counter() - t0)
                        line["group%d_l%d" % (n_gpus, l)] = {"host_to_host_ms": round(min(ts) * 1e3, 2),
                                                              "host_to_host_view_ms": round(min(tv) * 1e3, 2), "shards_used": used,
                                                              "byte_equal_to_one_device": bool(same)}
                finally:
                    g.close()
            line["host_to_host"] = ("eds_vcf_transform_host: pinned input, H2D + kernels + D2H into fresh malloc'd strings; "
                                    "_view: eds_vcf_transform_host_view, results in pinned memory kept by the context")
            del pv, pf
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
