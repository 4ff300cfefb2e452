#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/ref_driver).

Run in the build container only (needs /root/reference to have been compiled by
`make -C oracle ref`). The JSON files are committed; tests never call this script.

Cases:
  msa.json       seeded random alignments inside the reference's well-defined input
                 domain plus the reference's own test inputs
                 (tests/cpp/test_msa.cpp:20-229) and data/msa/small.msa, each at several l.
  leds.json      seeded random EDS (+SEDS) texts through eds_to_leds_linear /
                 eds_to_leds_cartesian, error cases included, plus data/eds/*.eds.
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


def run_msa(text: bytes, l: int):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "in.msa")
        with open(p, "wb") as f:
            f.write(text)
        e, s = os.path.join(d, "o.eds"), os.path.join(d, "o.seds")
        r = subprocess.run([DRIVER, "msa2eds", p, str(l), e, s], capture_output=True)
        if r.returncode != 0:
            return {"error": r.stderr.decode("latin-1").strip()}
        return {"eds": open(e, "rb").read().decode("latin-1"), "seds": open(s, "rb").read().decode("latin-1")}


def run_leds(eds: bytes, seds, l: int, compact: bool, threads: int = 1):
    with tempfile.TemporaryDirectory() as d:
        pe = os.path.join(d, "in.eds")
        with open(pe, "wb") as f:
            f.write(eds)
        ps = "-"
        if seds is not None:
            ps = os.path.join(d, "in.seds")
            with open(ps, "wb") as f:
                f.write(seds)
        oe, os_ = os.path.join(d, "o.leds"), os.path.join(d, "o.seds")
        r = subprocess.run(
            [DRIVER, "eds2leds", pe, ps, str(l), oe, os_ if seds is not None else "-", str(threads), "1" if compact else "0"],
            capture_output=True,
        )
        if r.returncode != 0:
            return {"error": r.stderr.decode("latin-1").strip()}
        out = {"eds": open(oe, "rb").read().decode("latin-1")}
        if seds is not None and os.path.exists(os_):
            out["seds"] = open(os_, "rb").read().decode("latin-1")
        return out


def fasta(rows, wrap, final_newline=True, names=None):
    parts = []
    for i, row in enumerate(rows):
        parts.append(">" + (names[i] if names else "seq%d" % (i + 1)))
        for k in range(0, len(row), wrap):
            parts.append(row[k : k + wrap])
    text = "\n".join(parts)
    return (text + "\n" if final_newline else text).encode()


def random_alignment(rng: random.Random):
    R = rng.choice([2, 2, 3, 3, 4, 5, 7, 12, 33, 40])
    C = rng.choice([1, 2, 3, 5, 8, 13, 24, 31, 32, 33, 47, 64, 65, 100, 150])
    wrap = rng.choice([1, 2, 3, 7, 12, 16, 17, 60, 80, C, C + 5])
    wrap = max(1, wrap)
    p_var = rng.choice([0.0, 0.02, 0.1, 0.3, 0.8])
    alphabet = rng.choice(["ACGT", "ACGT", "ACGTN", "acgtACGT"])
    first = [rng.choice(alphabet) for _ in range(C)]
    for c in range(C):
        if rng.random() < p_var * 0.3:
            first[c] = "-"
    rows = ["".join(first)]
    var_cols = [rng.random() < p_var for _ in range(C)]
    for _ in range(1, R):
        row = list(first)
        for c in range(C):
            if var_cols[c]:
                u = rng.random()
                if u < 0.35:
                    row[c] = rng.choice(alphabet)
                elif u < 0.55:
                    row[c] = "-"
        rows.append("".join(row))
    names = ["s" * rng.randint(1, 9) + str(i) for i in range(R)] if rng.random() < 0.5 else None
    return fasta(rows, wrap, final_newline=rng.random() < 0.8, names=names)


def msa_cases():
    rng = random.Random(20261018)
    cases = []
    fixed = [
        fasta(["AGTC--TCTATA", "AGTCCCTATATA", "AGTC--TATATA"], 60),  # test_msa.cpp:20-103
        fasta(["AGTCTA", "AGTCTA", "AGTCTA"], 60),  # :106-138
        fasta(["AGTCTA", "AGTGTA"], 60),  # :141-171 (single SNP shape)
        fasta(["--AGTC", "CCAGTC"], 60),  # leading gap
        fasta(["AGTC--", "AGTCGG"], 60),  # trailing gap
        open(os.path.join(REF, "data", "msa", "small.msa"), "rb").read(),
        fasta(["ACGT", "ACGA"], 60),
        fasta(["ACGT", "ACGA"], 2, final_newline=False),
        fasta(["A-GT", "A-GT"], 60),
        fasta(["A-CT", "AC-T"], 60),
        fasta(["acgt", "ACGT"], 60),
        fasta(["TCGTACGT", "AGGTACTT"], 60),
        fasta(["----", "----", "----"], 3),
        fasta(["A", "C"], 1),
        fasta(["A", "A"], 1),
    ]
    for text in fixed:
        for l in (0, 1, 2, 3, 4, 5, 10):
            cases.append({"msa": text.decode("latin-1"), "l": l, **run_msa(text, l)})
    for _ in range(260):
        text = random_alignment(rng)
        for l in rng.sample([0, 1, 2, 3, 4, 6, 10, 25], 2):
            cases.append({"msa": text.decode("latin-1"), "l": l, **run_msa(text, l)})
    return cases


def random_eds(rng: random.Random, with_sources: bool):
    n = rng.randint(1, 9)
    P = rng.randint(1, 5)
    loose = rng.random() < 0.15  # arbitrary subsets: merges may end with no surviving combination
    syms, srcs = [], []
    for _ in range(n):
        if rng.random() < 0.5:
            length = rng.choice([0, 1, 2, 3, 5, 8, 12])
            syms.append(["".join(rng.choice("ACGT") for _ in range(length))])
            srcs.append([[0]] if rng.random() < 0.85 else [sorted(rng.sample(range(1, P + 1), rng.randint(1, P)))])
        else:
            k = rng.randint(2, 3)
            alts = ["".join(rng.choice("ACGT") for _ in range(rng.randint(0, 3))) for _ in range(k)]
            sets = [set() for _ in range(k)]
            for p in range(1, P + 1):
                if loose and rng.random() < 0.5:
                    continue
                sets[rng.randrange(k)].add(p)
                if rng.random() < 0.2:
                    sets[rng.randrange(k)].add(p)
            for s in sets:
                if not s:
                    s.add(rng.randint(1, P))
                if rng.random() < 0.05:
                    s.add(0)
            syms.append(alts)
            srcs.append([sorted(s) for s in sets])
    compact_in = rng.random() < 0.3
    eds = ""
    for alts in syms:
        if compact_in and len(alts) == 1 and alts[0]:
            eds += alts[0]
        else:
            eds += "{" + ",".join(alts) + "}"
    if compact_in:
        # adjacent bare runs fuse into one symbol when parsed; rebuild sources to match
        fused_syms, fused_srcs = [], []
        prev_bare = False
        for alts, ss in zip(syms, srcs):
            bare = len(alts) == 1 and alts[0] != ""
            if bare and prev_bare:
                fused_syms[-1] = [fused_syms[-1][0] + alts[0]]
            else:
                fused_syms.append(list(alts))
                fused_srcs.append(ss)
            prev_bare = bare
        syms, srcs = fused_syms, fused_srcs
    if rng.random() < 0.2:
        eds = eds.replace("}", "}\n", 1) + "\n"
    seds = None
    if with_sources:
        seds = "".join("{" + ",".join(map(str, s)) + "}" for ss in srcs for s in ss)
        if rng.random() < 0.15:
            seds += "\n"
    return eds, seds


def leds_cases():
    rng = random.Random(1018)
    cases = []
    named = [
        ("{AAAA}{}{C,G}{TTTT}", None, 2),
        ("{AAAA}{A,A}{C}{G,G}{TTTT}", None, 2),
        ("{AAAA}{A,C}{G}{T,G}{TTTT}", "{0}{1,2}{2,3}{0}{1,2}{2,3}{0}", 2),
        ("{AAAA}{A,C}{G}{T,G}{TTTT}", "{0}{0,7}{2}{0}{0,9}{3}{0}", 2),
        ("{AAAA}{A,C}{G}{T,G}{TTTT}", "{0}{1}{2}{0}{1}{3}{0}", 2),
        ("{AAAA}{A,C}{G}{T,G}{TTTT}", "{0}{1}{2}{0}{3}{4}{0}", 2),
        ("{A,C}{G}{T,G}", None, 5),
        ("{AA}{C}{GG}{T}{AA}", None, 2),
        ("{AAAA}{CCCC}{G}{T,A}{TTTT}", None, 2),
        ("{AGTC}{,CC}{T}{C,A}{TATAAAT}{AA,GG}{ATA}{,GGGG}", "{0}{1,3}{2}{0}{1}{2,3}{0}{1,2}{3}{0}{1,3}{2}", 10),
        ("{AGTC}{,CC}{T}{C,A}{TATAAAT}{AA,GG}{ATA}{,GGGG}", "{0}{1,3}{2}{0}{1}{2,3}{0}{1,2}{3}{0}{1,3}{2}", 4),
        ("{AGTC}{,CC}{T}{C,A}{TATAAAT}{AA,GG}{ATA}{,GGGG}", "{0}{1,3}{2}{0}{1}{2,3}{0}{1,2}{3}{0}{1,3}{2}", 2),
        ("{AGTC}{,CC}{T}{C,A}{TATAAAT}{AA,GG}{ATA}{,GGGG}", None, 10),
        ("", None, 3),
        ("ACGT", None, 3),
        ("{A}{C,G}", "{0}{1}", 3),
        ("{A}{C,G}", "{0}{1}{x}", 3),
        ("{A}{C,G}", "{0}{1}{}", 3),
        ("{A}{C,G", None, 3),
        ("A}{C,G}", None, 3),
        ("{A,C}{G,T}", "{1}{2}{1}{2}", 0),
    ]
    for eds, seds, l in named:
        for compact in (True, False):
            cases.append(
                {"eds_in": eds, "seds_in": seds, "l": l, "compact": compact,
                 **run_leds(eds.encode(), None if seds is None else seds.encode(), l, compact)}
            )
    eds_dir = os.path.join(REF, "data", "eds")
    for name in sorted(os.listdir(eds_dir)):
        if not name.endswith(".eds") or "_l" in name:
            continue
        text = open(os.path.join(eds_dir, name), "rb").read()
        for l in (1, 4, 5):
            cases.append({"eds_in": text.decode("latin-1"), "seds_in": None, "l": l, "compact": True,
                          "name": name, **run_leds(text, None, l, True)})
    for i in range(420):
        with_sources = i % 3 != 0
        eds, seds = random_eds(rng, with_sources)
        l = rng.choice([1, 2, 3, 4, 6, 10])
        compact = rng.random() < 0.6
        threads = rng.choice([1, 1, 4])
        cases.append(
            {"eds_in": eds, "seds_in": seds, "l": l, "compact": compact,
             **run_leds(eds.encode(), None if seds is None else seds.encode(), l, compact, threads)}
        )
    return cases


def main():
    if not os.path.exists(DRIVER):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    for name, fn in (("msa.json", msa_cases), ("leds.json", leds_cases)):
        cases = fn()
        with open(os.path.join(HERE, name), "w") as f:
            json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/ref_driver (unmodified reference)",
                       "cases": cases}, f, indent=0)
        errs = sum(1 for c in cases if "error" in c)
        print(name, len(cases), "cases,", errs, "error cases,", os.path.getsize(os.path.join(HERE, name)), "bytes")


if __name__ == "__main__":
    main()
