"""CPU: pin the restatement (oracle/eds_oracle.cpp) to the reference.

Golden vectors in tests/golden/*.json were produced by the unmodified reference
library (oracle/_ref/ref_driver) through tests/golden/make_golden.py.
"""
import json
import os

import pytest

import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)["cases"]


def test_msa_golden_all():
    cases = _cases("msa.json")
    assert len(cases) > 500
    for c in cases:
        text = c["msa"].encode("latin-1")
        eds, seds = oracle_lib.msa2eds(text, c["l"])
        assert eds == c["eds"].encode("latin-1"), (c["msa"], c["l"])
        assert seds == c["seds"].encode("latin-1"), (c["msa"], c["l"])


def test_reference_unit_test_strings():
    # tests/cpp/test_msa.cpp:20-103 (golden strings held by the reference's own test)
    msa = b">seq1\nAGTC--TCTATA\n>seq2\nAGTCCCTATATA\n>seq3\nAGTC--TATATA\n"
    assert oracle_lib.msa2eds(msa, 0) == (b"{AGTC}{,CC}{T}{C,A}{TATA}", b"{0}{1,3}{2}{0}{1}{2,3}{0}")
    assert oracle_lib.msa2eds(msa, 4) == (b"{AGTC}{TC,CCTA,TA}{TATA}", b"{0}{1}{2}{3}{0}")


def test_survey_appendix_b_vectors():
    small = (b">seq1\nAGTC--TCTATA\nAATAAATA----\n>seq2\nAGTCCCTATATA\nAATAAATAGGGG\n"
             b">seq3\nAGTC--TATATA\nAATGGATA----\n")
    e0, s0 = oracle_lib.msa2eds(small, 0)
    assert e0 == b"{AGTC}{,CC}{T}{C,A}{TATAAAT}{AA,GG}{ATA}{,GGGG}"
    assert s0 == b"{0}{1,3}{2}{0}{1}{2,3}{0}{1,2}{3}{0}{1,3}{2}"
    assert oracle_lib.msa2eds(small, 10) == (
        b"{AGTC}{TCTATAAATAAATA,CCTATATAAATAAATAGGGG,TATATAAATGGATA}", b"{0}{1}{2}{3}")
    # config 1: msa2eds output -> eds2leds -l 10 LINEAR
    assert oracle_lib.eds2leds(e0, s0, 10) == (
        b"AGTC{TCTATAAATAAATA,TATATAAATGGATA,CCTATATAAATAAATAGGGG}\n", b"{0}{1}{3}{2}\n")
    cart, none = oracle_lib.eds2leds(e0, None, 10)
    assert none is None and len(cart) == 294 and cart.startswith(b"AGTC{TCTATAAATAAATA,")


def test_leds_golden_all():
    cases = _cases("leds.json")
    assert len(cases) > 400
    n_err = 0
    for c in cases:
        eds = c["eds_in"].encode("latin-1")
        seds = None if c["seds_in"] is None else c["seds_in"].encode("latin-1")
        if "error" in c:
            n_err += 1
            with pytest.raises(oracle_lib.OracleError) as ei:
                oracle_lib.eds2leds(eds, seds, c["l"], c["compact"])
            assert "Error: " + ei.value.message == c["error"], c
            continue
        out, sout = oracle_lib.eds2leds(eds, seds, c["l"], c["compact"])
        assert out == c["eds"].encode("latin-1"), c
        if seds is not None:
            assert sout == c["seds"].encode("latin-1"), c
    assert n_err >= 10


def test_reference_data_eds_pairs():
    # data/eds/X.eds -> data/eds/X_l<N>.eds pairs shipped by the reference (cartesian, compact, trailing \n);
    # inputs/outputs are embedded in leds.json under "name".
    named = [c for c in _cases("leds.json") if c.get("name")]
    assert len(named) >= 20


def _stats_line(st):
    return "stats total=%d processed=%d malformed=%d sv=%d groups=%d" % tuple(st[k] for k in oracle_lib.VCF_STAT_KEYS)


def test_vcf_golden_all():
    # outputs of the unmodified reference (tests/golden/make_golden_vcf.py), incl. data/vcf/{small,test_overlaps,
    # test_samepos} which the reference's own test_vcf.cpp reads
    cases = _cases("vcf.json")
    assert len(cases) > 300
    n_err = 0
    for c in cases:
        vcf, fa = c["vcf"].encode("latin-1"), c["fa"].encode("latin-1")
        if "error" in c:
            n_err += 1
            with pytest.raises(oracle_lib.OracleError) as ei:
                oracle_lib.vcf2eds(vcf, fa, c["l"])
            assert "Error: " + ei.value.message == c["error"], c["error"]
            continue
        eds, seds, st, _ = oracle_lib.vcf2eds(vcf, fa, c["l"])
        assert eds == c["eds"].encode("latin-1"), (c.get("name"), c["l"])
        assert seds == c["seds"].encode("latin-1"), (c.get("name"), c["l"])
        assert _stats_line(st) == c["stats"]
    assert n_err >= 10


def test_vcf_reference_shipped_outputs():
    # data/vcf/small.{eds,seds} and test_samepos as shipped by the reference
    named = {(c["name"], c["l"]): c for c in _cases("vcf.json") if c.get("name")}
    c = named[("test_samepos", 0)]
    assert (c["eds"], c["seds"]) == ("{ACGT}{A,C,G}{CGTACGT}", "{0}{1,2}{1}{2}{0}")
    assert named[("small", 0)]["eds"].startswith("{AGCT}{T,C}{AG}{C,G}{TAAGCTTACGA}{T}{CGATCG}")
