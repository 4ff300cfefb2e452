"""VCF front-end parity checks shared by the GPU tier (product library) and the CPU tier (emulator build).

Everything goes through the C ABI (eds_vcf_transform_host / _device) and is compared byte for byte with the golden
outputs of the unmodified reference (tests/golden/vcf.json) or with the oracle port on seeded random inputs."""
import json
import os
import random
import sys

import pytest

import oracle_lib
from edsparser_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_vcf  # noqa: E402  (random_pair only; never runs the reference here)

STAT_MAP = (("total", "total_variants"), ("processed", "processed_variants"), ("malformed", "skipped_malformed"),
            ("sv", "skipped_unsupported_sv"), ("groups", "variant_groups"))


def golden_cases():
    with open(os.path.join(HERE, "golden", "vcf.json")) as f:
        return json.load(f)["cases"]


def stats_line(st):
    return "stats total=%d processed=%d malformed=%d sv=%d groups=%d" % tuple(st[k] for _, k in STAT_MAP)


def check_case(ctx, vcf, fa, l, exp):
    """exp: dict with eds/seds/stats or error (golden entry), or None -> compare with the oracle."""
    if exp is None:
        try:
            e, s, st, warn = oracle_lib.vcf2eds(vcf, fa, l)
            exp = {"eds": e, "seds": s, "stats": "stats total=%d processed=%d malformed=%d sv=%d groups=%d" % tuple(
                st[k] for k in oracle_lib.VCF_STAT_KEYS), "n_warn": len(warn)}
        except oracle_lib.OracleError as err:
            exp = {"error": "Error: " + err.message}
    else:
        exp = dict(exp)
        for k in ("eds", "seds"):
            if k in exp and isinstance(exp[k], str):
                exp[k] = exp[k].encode("latin-1")
    if "error" in exp:
        with pytest.raises(capi.EdsError) as ei:
            ctx.vcf_transform_host(vcf, fa, l)
        assert "Error: " + ei.value.message == exp["error"], (ei.value.message, exp["error"])
        return "error"
    eds, seds, st, sv_lines = ctx.vcf_transform_host(vcf, fa, l)
    assert ctx.vcf_transform_host_view(vcf, fa, l)[:2] == (eds, seds)  # pinned-view form of the same call
    assert eds == exp["eds"], (l, eds[:300], exp["eds"][:300])
    assert seds == exp["seds"], (l, seds[:300], exp["seds"][:300])
    assert stats_line(st) == exp["stats"]
    assert len(sv_lines) == st["skipped_unsupported_sv"]
    return "ok"


def check_golden(ctx, stride=1, offset=0, max_vcf_bytes=None):
    n = n_err = 0
    for c in golden_cases()[offset::stride]:
        if max_vcf_bytes and len(c["vcf"]) > max_vcf_bytes:
            continue
        r = check_case(ctx, c["vcf"].encode("latin-1"), c["fa"].encode("latin-1"), c["l"], c)
        n += 1
        n_err += r == "error"
    return n, n_err


def check_shipped(ctx):
    """data/vcf/{small,test_overlaps,test_samepos} as shipped by the reference (embedded in vcf.json by name)."""
    named = [c for c in golden_cases() if c.get("name")]
    assert len(named) == 9
    for c in named:
        check_case(ctx, c["vcf"].encode("latin-1"), c["fa"].encode("latin-1"), c["l"], c)


def check_random(ctx, seed, n_cases, ls=(0, 3), **kw):
    rng = random.Random(seed)
    for _ in range(n_cases):
        vcf, fa = make_golden_vcf.random_pair(rng, **kw)
        for l in ls:
            check_case(ctx, vcf, fa, l, None)


FA = b">chr1 demo\nACGTACGTAC\nGTACGTACGT\nACGT\n"
HDR = b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3\n"


def check_edges(ctx):
    def run(body, fa=FA, l=0, hdr=HDR):
        return check_case(ctx, hdr + body, fa, l, None)

    run(b"")                                                      # no records: the reference as one symbol
    run(b"", fa=b">x\nACGT")                                      # FASTA without a final newline
    run(b"chr1\t3\t.\tG\tT\t.\t.\t.\tGT\t0|1\t1|1\t.|.\n")
    run(b"chr1\t3\t.\tG\tT\t.\t.\t.\tGT\t0|1\t1|1\t.|.")       # last line without a newline
    run(b"chr1\t1\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1|1\t0|0\n")       # first base
    run(b"chr1\t24\t.\tT\tC,<DEL>\t.\t.\t.\tGT\t0|1\t2|1\t0|0\n")  # last base, deletion allele
    run(b"chr1\t5\t.\tA\t<INS>,<DEL>,AGG\t.\t.\t.\tGT\t1|2\t3|0\t0/3\n")
    run(b"chr1\t5\t.\tA\t<DUP>\t.\t.\t.\tGT\t1|0\t0|0\t0|0\nchr1\t7\t.\tG\t<>\t.\t.\t.\tGT\t1|0\t0|0\t0|0\n")
    run(b"chr1 5 . A T . . . GT 0|1 1|1 0|0\n")                   # blanks instead of tabs (:269-279)
    run(b"chr1\t5\t.\tA T\t. . . . GT 0|1 1|1\n")                # fewer than five tab fields -> blank split
    run(b"chr1\t\t5\t\t.\tA\tT\t.\t.\t.\tGT\t\t0|1\t1|1\t0|0\n")  # empty tab fields are dropped
    run(b"chr1\t5\n\nchr1\tx\t.\tA\tT\n#note\nchr1\t+6\t.\tC\tG\t.\t.\t.\tGT\t1\t0\t.\n")
    run(b"chr1\t99999999999999999999999\t.\tA\tT\n")             # stoull overflow -> malformed
    run(b"chr1\t5\t.\tA\tT,,G,\t.\t.\t.\tGT\t2|0\t3|1\t0|0\n")     # empty ALT piece, trailing comma
    run(b"chr1\t5\t.\tA\tT,G\t.\t.\t.\tGT:DP\t0|7:3\t-1|2:4\t99999999999|1:5\n")  # out-of-list, negative, overflow
    run(b"chr1\t5\t.\tA\tT\t.\t.\t.\tGT\t0/1|1\t|1\t1|\n")
    run(b"chr1\t5\t.\tA\tT\t.\t.\t.\tGT\t:0|1\tx\t 1| 0\n")
    run(b"chr1\t5\t.\tA\tT\t.\t.\t.\tGT\n")                      # nine fields: no genotypes -> {0}
    run(b"chr1\t5\t.\tA\tT\n", l=2)                               # ... which LINEAR then rejects
    run(b"chr1\t5\t.\tA\tT\t.\t.\t.\tGT\t0|1\t1|1\t0|0\t1|0\t0|1\n")  # more sample columns than the header
    run(b"chr1\t5\t.\tACG\tA\t.\t.\t.\tGT\t0|1\t1|1\t0|0\nchr1\t6\t.\tC\tT\t.\t.\t.\tGT\t1|0\t0|0\n"
        b"chr1\t7\t.\tG\tGA,C\t.\t.\t.\tGT\t0|2\t1|1\t0|0\t1|1\n")   # overlapping, differing sample counts
    run(b"chr1\t9\t.\tA\tT\t.\t.\t.\tGT\t0|1\t1|1\t0|0\nchr1\t3\t.\tG\tC\t.\t.\t.\tGT\t1|1\t0|0\t0|1\n")  # unsorted
    run(b"chr1\t5\t.\tA\tT\t.\t.\t.\tGT\t0|1\t1|1\t0|0\nchr1\t5\t.\tA\tG\t.\t.\t.\tGT\t1|0\t0|0\t0|1\n", l=4)
    run(b"chr1\t5\t.\tA\tA\t.\t.\t.\tGT\t0|1\t1|1\t0|0\n")       # ALT equals the reference span
    run(b"chr1\t5\t.\tG\tT,T\t.\t.\t.\tGT\t0|1\t2|2\t0|0\n")     # REF field differs from the FASTA, duplicate ALTs
    # degenerate VCF texts
    check_case(ctx, b"", FA, 0, None)                              # empty file: the reference as one symbol
    check_case(ctx, b"\n\n", FA, 0, None)
    check_case(ctx, b"##only a header", FA, 0, None)
    check_case(ctx, HDR, FA, 3, None)                             # no records, through the merge
    # a later record with far more sample columns than the first one: the genotype kernel is re-run with wider bitsets
    wide = b"\t".join(b"%d|%d" % (i % 2, (i // 2) % 2) for i in range(70))
    run(b"chr1\t3\t.\tG\tT\t.\t.\t.\tGT\t0|1\t1|1\n" + b"chr1\t9\t.\tA\tC,G\t.\t.\t.\tGT\t" + wide + b"\n" +
        b"chr1\t15\t.\tG\tT\t.\t.\t.\tGT\t1|1\n")
    # a line exactly one 16-byte vector / one 512-byte tile long around the genotype columns
    for pad in (0, 1, 15, 16, 17):
        info = b"X" * (460 + pad)
        run(b"chr1\t3\t.\tG\tT\t.\t.\t" + info + b"\tGT\t0|1\t1|1\t0|0\n")
    # FASTA edge cases
    run(b"chr1\t3\t.\tG\tT\t.\t.\t.\tGT\t0|1\t1|1\t0|0\n", fa=b">a\nACGTAC\n>b\nGGGGGG\n")  # second record ignored
    run(b"chr1\t3\t.\tG\tT\t.\t.\t.\tGT\t0|1\t1|1\t0|0\n", fa=b">a\nACGT\nAC\n\n\n")            # trailing blank lines
    for bad_fa, msg in ((b"", "Invalid FASTA format"), (b"ACGT\n", "Invalid FASTA format"), (b">only", "FASTA file is empty"),
                        (b">only\n", "FASTA file is empty")):
        with pytest.raises(capi.EdsError) as ei:
            ctx.vcf_transform_host(HDR, bad_fa, 0)
        assert msg in ei.value.message and ei.value.status == capi.EDS_ERR_RUNTIME
        with pytest.raises(oracle_lib.OracleError) as oi:
            oracle_lib.vcf2eds(HDR, bad_fa, 0)
        assert oi.value.message == ei.value.message
    # inputs the reference is undefined on are refused
    for body, fa in ((b"chr1\t0\t.\tA\tT\n", FA), (b"chr1\t24\t.\tTA\tT\n", FA), (b"chr1\t30\t.\tT\tA\n", FA),
                     (b"chr1\t3\t.\tG\tT\n", b">a\nACGT\nAC\nACGT\n"), (b"chr1\t3\t.\tG\tT\n", b">a\r\nACGT\r\n"),
                     (b"chr1\t3\t.\tG\tT\n", b">a\n\nACGT\n")):
        with pytest.raises(capi.EdsError) as ei:
            ctx.vcf_transform_host(HDR + body, fa, 0)
        assert ei.value.status == capi.EDS_ERR_BAD_VCF, ei.value.message


def check_wide(ctx, n_samples=300, n_sites=12, seed=5):
    """Sample matrix wider than one 512-byte tile and than 32 ids: multi-word bitsets, id widths 1-3."""
    rng = random.Random(seed)
    vcf, fa = make_golden_vcf.random_pair(rng, n_ref=400, n_sites=n_sites, n_samples=n_samples)
    for l in (0, 5):
        check_case(ctx, vcf, fa, l, None)


def synth_vcf(n_bases, n_sites, n_samples, seed=1, wrap=60, overlap_frac=0.01, same_pos=True):
    """BASELINE config 5 shape at any size, numpy: random ACGT reference (FASTA wrap 60); sorted distinct
    sites, 80 % SNP / 10 % insertion (1-5 bp) / 10 % deletion (2-5 bp REF); a 1 % sub-population of same-position /
    adjacent companion records (grouping + the unstable sort); phased diploid genotypes, per-site ALT frequency
    ~ Beta(0.3, 2.0). same_pos=False keeps the companions at the next position only (no ties for the unstable sort).
    Returns (vcf bytes, fasta bytes)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    letters = "ACGT"
    ref = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n_bases)]
    refs = ref.tobytes()
    fa = b">chr1 synthetic\n" + b"\n".join(refs[i:i + wrap] for i in range(0, n_bases, wrap)) + b"\n"
    # distinct positions at least 8 apart, so that only the companions overlap
    n_sites = min(n_sites, (n_bases - 16) // 8)
    pos = np.sort(rng.choice((n_bases - 16) // 8, size=n_sites, replace=False)) * 8 + 1 + rng.integers(0, 2, n_sites)
    kinds = rng.random(n_sites)
    head = []
    for i, p in enumerate(pos.tolist()):
        k = kinds[i]
        if k < 0.8:
            r = refs[p - 1:p].decode()
            a = letters[(letters.index(r) + 1 + int(rng.integers(0, 3))) % 4]
        elif k < 0.9:
            r = refs[p - 1:p].decode()
            a = r + "".join(letters[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 6))))
        else:
            r = refs[p - 1:p - 1 + int(rng.integers(2, 6))].decode()
            a = r[0]
        head.append("chr1\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT" % (p, r, a))
        if rng.random() < overlap_frac:
            q = p + (int(rng.integers(0, 2)) if same_pos else 1)
            r2 = refs[q - 1:q].decode()
            a2 = letters[(letters.index(r2) + 1) % 4] + ("" if rng.random() < 0.5 else "T")
            head.append("chr1\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT" % (q, r2, a2))
    n_rec = len(head)
    out = [b"##fileformat=VCFv4.2\n", b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT" +
           b"".join(b"\tS%d" % i for i in range(n_samples)) + b"\n"]
    block = 4096
    for b0 in range(0, n_rec, block):
        nb = min(block, n_rec - b0)
        freq = rng.beta(0.3, 2.0, nb)
        gt = rng.random((nb, n_samples, 2)) < freq[:, None, None]
        cell = np.empty((nb, n_samples, 4), dtype=np.uint8)
        cell[:, :, 0] = ord("\t")
        cell[:, :, 1] = ord("0") + gt[:, :, 0]
        cell[:, :, 2] = ord("|")
        cell[:, :, 3] = ord("0") + gt[:, :, 1]
        rows = cell.reshape(nb, n_samples * 4)
        for i in range(nb):
            out.append(head[b0 + i].encode())
            out.append(rows[i].tobytes())
            out.append(b"\n")
    return b"".join(out), fa


def conserved_text(eds, seds):
    """Concatenation of the symbols whose only source set is {0} (the common text between variant groups)."""
    syms = eds[1:-1].split(b"}{") if eds else []
    sets = seds[1:-1].split(b"}{") if seds else []
    out, k = [], 0
    for sym in syms:
        n = sym.count(b",") + 1
        if n == 1 and sets[k] == b"0":
            out.append(sym)
        k += n
    assert k == len(sets)
    return b"".join(out)


def expected_conserved_text(vcf, fa):
    """The reference outside the groups of overlapping records, from the record heads alone (sorted input)."""
    ref = fa.split(b"\n", 1)[1].replace(b"\n", b"")
    spans = []
    for line in vcf.split(b"\n"):
        if not line or line[:1] == b"#":
            continue
        f = line.split(b"\t", 5)
        spans.append((int(f[1]) - 1, int(f[1]) - 1 + len(f[3])))
    spans.sort(key=lambda t: t[0])
    out, cursor, i = [], 0, 0
    while i < len(spans):
        lo, hi = spans[i]
        j = i + 1
        while j < len(spans) and spans[j][0] < hi:
            hi = max(hi, spans[j][1])
            j += 1
        out.append(ref[cursor:lo])
        cursor = hi
        i = j
    out.append(ref[cursor:])
    return b"".join(out)
