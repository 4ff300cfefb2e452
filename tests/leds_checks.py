"""Parity checks of the l-EDS merge path (eds_leds_merge_host), shared by the GPU tier and the emulator tier."""
import json
import os

import numpy as np

import gen
import oracle_lib
from edsparser_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))

STATUS_OF_ORACLE = {1: capi.EDS_ERR_RUNTIME, 2: capi.EDS_ERR_INVALID_ARGUMENT, 3: capi.EDS_ERR_OUT_OF_RANGE}


def golden_cases(stride=1):
    with open(os.path.join(HERE, "golden", "leds.json")) as f:
        return json.load(f)["cases"][::stride]


def check_golden(ctx, stride=1):
    n = n_err = 0
    for c in golden_cases(stride):
        eds = c["eds_in"].encode("latin-1")
        seds = None if c["seds_in"] is None else c["seds_in"].encode("latin-1")
        n += 1
        if "error" in c:
            n_err += 1
            try:
                ctx.leds_merge_host(eds, seds, c["l"], c["compact"])
            except capi.EdsError as e:
                assert "Error: " + e.message == c["error"], (c, e.message)
            else:
                raise AssertionError(("no error", c))
            continue
        out, sout, _ = ctx.leds_merge_host(eds, seds, c["l"], c["compact"])
        assert ctx.leds_merge_host_view(eds, seds, c["l"], c["compact"])[:2] == (out, sout)  # pinned-view form
        assert out == c["eds"].encode("latin-1"), (c, out)
        if seds is not None:
            assert sout == c["seds"].encode("latin-1"), (c, sout)
    return n, n_err


BUDGET = 1 << 22  # keeps a runaway cross product from exhausting host memory in the oracle


def same_as_oracle(ctx, eds, seds, l, compact=True):
    try:
        exp = oracle_lib.eds2leds(eds, seds, l, compact, max_out_bytes=BUDGET)
    except oracle_lib.OracleError as oe:
        if oe.status == 4:  # the oracle's own size guard: ours must refuse or finish, never crash
            try:
                ctx.leds_merge_host(eds, seds, l, compact, max_output_bytes=BUDGET)
            except capi.EdsError as e:
                assert e.status == capi.EDS_ERR_BUDGET, e
            return "budget"
        try:
            ctx.leds_merge_host(eds, seds, l, compact)
        except capi.EdsError as e:
            assert e.status == STATUS_OF_ORACLE[oe.status], (eds, seds, l, e, oe.message)
            assert e.message == oe.message, (eds, seds, l, e.message, oe.message)
            return "error"
        raise AssertionError(("expected error", oe.message, eds, seds, l))
    out, sout, rounds = ctx.leds_merge_host(eds, seds, l, compact)
    assert out == exp[0], (eds, seds, l, compact, out, exp[0])
    assert sout == exp[1], (eds, seds, l, compact, sout, exp[1])
    return rounds


def check_survey_vectors(ctx):
    # SURVEY.md Appendix B
    e0 = b"{AGTC}{,CC}{T}{C,A}{TATAAAT}{AA,GG}{ATA}{,GGGG}"
    s0 = b"{0}{1,3}{2}{0}{1}{2,3}{0}{1,2}{3}{0}{1,3}{2}"
    assert ctx.leds_merge_host(e0, s0, 2)[:2] == (b"AGTC{TC,TA,CCTA}TATAAAT{AA,GG}ATA{,GGGG}\n", b"{0}{1}{3}{2}{0}{1,2}{3}{0}{1,3}{2}\n")
    assert ctx.leds_merge_host(e0, s0, 4)[:2] == (b"AGTC{TC,TA,CCTA}TATAAAT{AAATA,AAATAGGGG,GGATA}\n", b"{0}{1}{3}{2}{0}{1}{2}{3}\n")
    # config 1 (second half): eds2leds -l 10 LINEAR
    assert ctx.leds_merge_host(e0, s0, 10)[:2] == (b"AGTC{TCTATAAATAAATA,TATATAAATGGATA,CCTATATAAATAAATAGGGG}\n", b"{0}{1}{3}{2}\n")
    cart, none, _ = ctx.leds_merge_host(e0, None, 10)
    assert none is None and len(cart) == 294 and cart.startswith(b"AGTC{TCTATAAATAAATA,") and cart.endswith(b"CCTATATAAATGGATAGGGG}\n")
    assert ctx.leds_merge_host(b"{AAAA}{}{C,G}{TTTT}", None, 2)[0] == b"AAAA{C,G}TTTT\n"
    assert ctx.leds_merge_host(b"{AAAA}{A,A}{C}{G,G}{TTTT}", None, 2)[0] == b"AAAA{ACG,ACG,ACG,ACG}TTTT\n"
    assert ctx.leds_merge_host(b"{AAAA}{A,C}{G}{T,G}{TTTT}", b"{0}{1,2}{2,3}{0}{1,2}{2,3}{0}", 2)[:2] == (
        b"AAAA{AGT,AGG,CGT,CGG}TTTT\n", b"{0}{1,2}{2}{2}{2,3}{0}\n")
    assert ctx.leds_merge_host(b"{AAAA}{A,C}{G}{T,G}{TTTT}", b"{0}{0,7}{2}{0}{0,9}{3}{0}", 2)[:2] == (
        b"AAAA{AGT,AGG,CGT}TTTT\n", b"{0}{0}{3}{2}{0}\n")
    assert ctx.leds_merge_host(b"{AAAA}{A,C}{G}{T,G}{TTTT}", b"{0}{1}{2}{0}{1}{3}{0}", 2)[:2] == (b"AAAAAGTTTTT\n", b"{0}{1}{0}\n")
    assert ctx.leds_merge_host(b"{A,C}{G}{T,G}", None, 5)[0] == b"{AGT,AGG,CGT,CGG}\n"
    assert ctx.leds_merge_host(b"{AA}{C}{GG}{T}{AA}", None, 2)[0] == b"AACGGTAA\n"
    assert ctx.leds_merge_host(b"{AAAA}{CCCC}{G}{T,A}{TTTT}", None, 2)[0] == b"AAAACCCCG{T,A}TTTT\n"
    assert ctx.leds_merge_host(b"{AAAA}{CCCC}{G}{T,A}{TTTT}", None, 2, compact=False)[0] == b"{AAAA}{CCCCG}{T,A}{TTTT}\n"
    # whitespace anywhere is dropped (eds.cpp:46); compact input; empty EDS
    assert ctx.leds_merge_host(b" {AA}\n{C} {G G}{T}{A\tA}\n", None, 2)[0] == b"AACGGTAA\n"
    assert ctx.leds_merge_host(b"AA{C,G}T", None, 2)[0] == oracle_lib.eds2leds(b"AA{C,G}T", None, 2)[0]
    assert ctx.leds_merge_host(b"", None, 3)[0] == b"\n"


def check_errors(ctx):
    def err(eds, seds, l):
        try:
            ctx.leds_merge_host(eds, seds, l)
        except capi.EdsError as e:
            return e.status, e.message
        return 0, ""

    assert err(b"{A,C}{G,T}", b"{1}{2}{1}{2}", 0) == (capi.EDS_ERR_INVALID_ARGUMENT, "context_length must be > 0 for l-EDS transformation")
    assert err(b"{AAAA}{A,C}{G}{T,G}{TTTT}", b"{0}{1}{2}{0}{3}{4}{0}", 2) == (
        capi.EDS_ERR_RUNTIME, "Merging positions 1 and 2 results in empty set (no valid source intersections)")
    assert err(b"{A}{C,G}", b"{0}{1}", 3)[1] == "sEDS: Source count (2) does not match EDS cardinality (3)"
    assert err(b"{A}{C,G}", b"{0}{1}{x}", 3)[1] == "sEDS: Invalid character 'x' at position 7"
    assert err(b"{A}{C,G}", b"{0}{1}{}", 3)[1] == "sEDS: Empty path set at string 2"
    assert err(b"{A}{C,G", None, 3)[1] == "Expected '}' at position 7"
    assert err(b"A}{C,G}", None, 3)[1] == "Expected '{' at position 0"
    assert err(b"{A}{C,G}", b"", 3)[1] == "sEDS input is empty"
    assert err(b"{A}{C,G}", b"{0}{1}{99999999999}", 3)[0] == capi.EDS_ERR_OUT_OF_RANGE
    # CARTESIAN blow-up is refused, not truncated, when a budget is given
    big = b"".join(b"{A,C,G,T}" for _ in range(14))
    try:
        ctx.leds_merge_host(big, None, 5, max_output_bytes=1 << 20)
    except capi.EdsError as e:
        assert e.status == capi.EDS_ERR_BUDGET
    else:
        raise AssertionError("expected EDS_ERR_BUDGET")


def check_random(ctx, seed, n_cases, max_sym=30, paths=4):
    rng = np.random.default_rng(seed)
    n_err = 0
    for i in range(n_cases):
        linear = rng.random() < 0.7
        eds, seds = gen.random_eds(rng, n_sym=int(rng.integers(1, max_sym + 1)), paths=paths, with_sources=linear,
                                   p_deg=float(rng.choice([0.2, 0.5, 0.8])), compact_in=bool(rng.integers(0, 2)))
        l = int(rng.choice([1, 2, 3, 5, 10]))
        if not linear and eds.count(b",") > 14:
            l = min(l, 2)  # keep the cartesian product small
        r = same_as_oracle(ctx, eds, seds, l, compact=bool(rng.integers(0, 2)))
        n_err += r == "error"
    return n_err


def check_long_strings_and_many_paths(ctx):
    rng = np.random.default_rng(4)
    # leaves longer than the per-thread copy limit; 300 paths (10 words of bitset)
    a = bytes(gen.ALPHABET[rng.integers(0, 4, 700)])
    b = bytes(gen.ALPHABET[rng.integers(0, 4, 300)])
    eds = b"{" + a + b"}{A,C}{GT}{" + b + b",T}{ACGTACGTACGT}"
    seds = b"{0}{" + b",".join(b"%d" % i for i in range(1, 200)) + b"}{" + b",".join(b"%d" % i for i in range(150, 301)) + \
        b"}{0}{" + b",".join(b"%d" % i for i in range(1, 160)) + b"}{" + b",".join(b"%d" % i for i in range(100, 301)) + b"}{0}"
    for l in (1, 3, 400, 1000):
        same_as_oracle(ctx, eds, seds, l)
        same_as_oracle(ctx, eds, None, l, compact=False)


def check_msa_pipeline(ctx, seed=6, n_cases=5):
    """config 1 as a pipeline: msa2eds (l = 0) output fed to eds2leds LINEAR, both on the device path."""
    rng = np.random.default_rng(seed)
    for _ in range(n_cases):
        text, _, _ = gen.random_msa_text(rng, max_rows=6, max_cols=80)
        e0, s0, _ = ctx.msa_transform_host(text, 0)
        assert (e0, s0) == oracle_lib.msa2eds(text, 0)
        for l in (2, 10):
            same_as_oracle(ctx, e0, s0, l)
