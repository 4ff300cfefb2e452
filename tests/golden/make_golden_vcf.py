#!/usr/bin/env python3
"""Generate tests/golden/vcf.json from the UNMODIFIED reference (oracle/_ref/ref_driver vcf2eds).

Build container only (needs `make -C oracle ref`). The JSON is committed; tests never run this script.
Cases: the reference's own data/vcf fixtures (small, test_overlaps, test_samepos) and seeded random
VCF + FASTA pairs: SNPs / insertions / deletions / multi-allelic sites, <DEL> <INS> and unsupported <DUP>,
overlapping and same-position records, unsorted input (more than 16 records so std::sort leaves the
insertion-sort regime), missing and unphased genotypes, FORMAT suffixes, allele indices past the ALT list,
no sample columns, malformed lines; each at l = 0 and at some l > 0.
"""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
REF = os.environ.get("EDS_REFERENCE", "/root/reference")


def run_vcf(vcf: bytes, fa: bytes, l: int):
    with tempfile.TemporaryDirectory() as d:
        pv, pf = os.path.join(d, "in.vcf"), os.path.join(d, "ref.fa")
        open(pv, "wb").write(vcf)
        open(pf, "wb").write(fa)
        e, s = os.path.join(d, "o.eds"), os.path.join(d, "o.seds")
        r = subprocess.run([DRIVER, "vcf2eds", pv, pf, str(l), e, s], capture_output=True)
        if r.returncode != 0:
            return {"error": r.stderr.decode("latin-1").strip().splitlines()[-1]}
        stats = [x for x in r.stdout.decode().splitlines() if x.startswith("stats ")][0]
        return {"eds": open(e, "rb").read().decode("latin-1"), "seds": open(s, "rb").read().decode("latin-1"), "stats": stats}


def random_pair(rng, n_ref=None, n_sites=None, n_samples=None):
    n_ref = n_ref or rng.randint(30, 400)
    ref = "".join(rng.choice("ACGT") for _ in range(n_ref))
    wrap = rng.choice([10, 17, 60, 80, n_ref])
    fa = ">chr1 test\n" + "\n".join(ref[i:i + wrap] for i in range(0, n_ref, wrap)) + "\n"
    n_samples = rng.choice([0, 1, 3, 5, 40, 70]) if n_samples is None else n_samples
    n_sites = n_sites if n_sites is not None else rng.randint(0, 40)
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT" + "".join("\tS%d" % i for i in range(n_samples))]
    sites = []
    for _ in range(n_sites):
        pos = rng.randint(1, n_ref)
        kind = rng.random()
        if kind < 0.55:
            reflen = 1
        elif kind < 0.8:
            reflen = rng.randint(2, 5)
        else:
            reflen = 1
        reflen = min(reflen, n_ref - pos + 1)
        refa = ref[pos - 1:pos - 1 + reflen]
        if rng.random() < 0.1:
            refa = "".join(rng.choice("ACGT") for _ in range(reflen))  # REF is not checked against the FASTA
        nalt = rng.choice([1, 1, 1, 2, 3])
        alts = []
        for _a in range(nalt):
            u = rng.random()
            if u < 0.5:
                alts.append(rng.choice("ACGT"))
            elif u < 0.7:
                alts.append(refa[0] + "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 5))))
            elif u < 0.8:
                alts.append(refa[0])
            elif u < 0.87:
                alts.append("<DEL>")
            elif u < 0.93:
                alts.append("<INS>")
            elif u < 0.96:
                alts.append("<DUP>")
            else:
                alts.append(refa)
        gts = []
        for _s in range(n_samples):
            u = rng.random()
            sep = "|" if rng.random() < 0.8 else "/"
            if u < 0.03:
                g = "."
            elif u < 0.06:
                g = "." + sep + "."
            elif u < 0.1:
                g = str(rng.randint(0, nalt))
            elif u < 0.13:
                g = str(rng.randint(0, nalt + 2)) + sep + str(rng.randint(0, nalt))
            else:
                g = str(rng.choice([0, 0, 0, rng.randint(0, nalt)])) + sep + str(rng.choice([0, 0, rng.randint(0, nalt)]))
            if rng.random() < 0.1:
                g += ":%d:PASS" % rng.randint(1, 99)
            gts.append(g)
        fmt = "GT" if rng.random() < 0.8 else "GT:DP:FT"
        sites.append((pos, "chr1\t%d\t.\t%s\t%s\t99\tPASS\t.\t%s%s" % (pos, refa, ",".join(alts), fmt, "".join("\t" + g for g in gts))))
    if rng.random() < 0.7:
        sites.sort(key=lambda t: t[0])
    lines += [t[1] for t in sites]
    if rng.random() < 0.15:
        lines.insert(rng.randint(2, len(lines)), "chr1\t12")  # malformed
    if rng.random() < 0.1:
        lines.insert(rng.randint(2, len(lines)), "chr1\tabc\t.\tA\tC\t.\t.\t.\tGT")  # invalid POS
    text = "\n".join(lines) + ("\n" if rng.random() < 0.9 else "")
    return text.encode(), fa.encode()


def main():
    rng = random.Random(20261018)
    cases = []
    for name in ("small", "test_overlaps", "test_samepos"):
        vcf = open(os.path.join(REF, "data", "vcf", name + ".vcf"), "rb").read()
        fa = open(os.path.join(REF, "data", "vcf", name + ".fa"), "rb").read()
        for l in (0, 3, 10):
            c = {"name": name, "vcf": vcf.decode("latin-1"), "fa": fa.decode("latin-1"), "l": l}
            c.update(run_vcf(vcf, fa, l))
            cases.append(c)
    for i in range(150):
        vcf, fa = random_pair(rng)
        for l in (0, rng.choice([1, 2, 4, 10])):
            c = {"vcf": vcf.decode("latin-1"), "fa": fa.decode("latin-1"), "l": l}
            c.update(run_vcf(vcf, fa, l))
            cases.append(c)
    for i in range(6):  # many records: std::sort beyond insertion sort, wide sample matrix
        vcf, fa = random_pair(rng, n_ref=3000, n_sites=rng.randint(60, 300), n_samples=rng.choice([5, 130]))
        c = {"vcf": vcf.decode("latin-1"), "fa": fa.decode("latin-1"), "l": 0}
        c.update(run_vcf(vcf, fa, 0))
        cases.append(c)
    with open(os.path.join(HERE, "vcf.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden_vcf.py", "cases": cases}, f)
    n_err = sum("error" in c for c in cases)
    print(len(cases), "cases,", n_err, "errors")
    for c in cases:
        if "error" in c:
            print("  ", c["error"][:100])


if __name__ == "__main__":
    main()
