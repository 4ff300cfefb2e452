"""The msa2eds / eds2leds tools (edsparser_b200/host/tools): CLI contract of the reference tools
(src/cpp/tools/msa2eds.cpp, eds2leds.cpp). CPU tier: argument handling and the loud no-GPU failure.
GPU tier: files written, names, stdout text, bytes against the oracle."""
import os
import subprocess

import pytest

import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "edsparser_b200", "bin")
SMALL = (b">seq1\nAGTC--TCTATA\nAATAAATA----\n>seq2\nAGTCCCTATATA\nAATAAATAGGGG\n"
         b">seq3\nAGTC--TATATA\nAATGGATA----\n")


@pytest.fixture(scope="module", autouse=True)
def tools():
    if not (os.path.exists(os.path.join(BIN, "msa2eds")) and os.path.exists(os.path.join(BIN, "eds2leds"))
            and os.path.exists(os.path.join(BIN, "vcf2eds"))):
        subprocess.check_call(["make", "-C", ROOT, "lib", "host"], stdout=subprocess.DEVNULL)


def run(tool, *args):
    p = subprocess.run([os.path.join(BIN, tool), *args], capture_output=True)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


def test_help_and_validation(tmp_path):
    rc, out, err = run("msa2eds", "--help")
    assert rc == 0 and out.startswith("msa2eds - Transform MSA") and "[Performance] Runtime:" in err
    rc, out, err = run("eds2leds", "-h")
    assert rc == 0 and "LINEAR" in out and "[Performance] Runtime:" in err
    rc, out, err = run("msa2eds", "-i", str(tmp_path / "x.fasta"))
    assert rc == 1 and err.startswith("Error: Input file must be an MSA file (.msa)\nGot: ")
    rc, out, err = run("eds2leds", "-i", str(tmp_path / "x.eds"), "-l", "0")
    assert rc == 1 and err.startswith("Error: Context length must be > 0\n")
    rc, out, err = run("eds2leds", "-i", str(tmp_path / "x.eds"), "-l", "3", "-t", "0")
    assert rc == 1 and err.startswith("Error: Number of threads must be >= 1\n")
    rc, out, err = run("eds2leds", "-i", str(tmp_path / "x.txt"), "-l", "3")
    assert rc == 1 and err.startswith("Error: Input file must be an EDS file (.eds)\n")
    rc, out, err = run("eds2leds", "-l", "3")
    assert rc == 1 and "the option '--input' is required but missing" in err
    rc, out, err = run("msa2eds", "-i", str(tmp_path / "missing.msa"))
    assert rc == 1 and "Failed to open input file" in err
    rc, out, err = run("vcf2eds", "--help")
    assert rc == 0 and out.startswith("vcf2eds - Transform VCF") and "[Performance] Runtime:" in err
    rc, out, err = run("vcf2eds", "-i", str(tmp_path / "x.txt"), "-r", str(tmp_path / "r.fa"))
    assert rc == 1 and err.startswith("Error: Input file must be a VCF file (.vcf)\nGot: ")
    rc, out, err = run("vcf2eds", "-i", str(tmp_path / "x.vcf"), "-r", str(tmp_path / "r.fa"))
    assert rc == 1 and err.startswith("Error: Reference FASTA file not found: ")
    rc, out, err = run("vcf2eds", "-i", str(tmp_path / "x.vcf"))
    assert rc == 1 and "the option '--reference' is required but missing" in err


def test_fails_loudly_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    (tmp_path / "a.msa").write_bytes(SMALL)
    rc, out, err = run("msa2eds", "-i", str(tmp_path / "a.msa"))
    assert rc == 1 and "no CPU fallback" in err and not (tmp_path / "a.eds").exists()


@pytest.mark.gpu
def test_config1_through_the_tools(tmp_path):
    # BASELINE config 1: msa2eds on the reference's data/msa alignment, then eds2leds -l 10 LINEAR
    msa = tmp_path / "small.msa"
    msa.write_bytes(SMALL)
    rc, out, err = run("msa2eds", "-i", str(msa))
    assert rc == 0, err
    assert out == ("MSA → EDS transformation\n  Input: \"%s\"\nTransformation complete!\n  Output: \"%s\"\n  Sources: \"%s\"\n"
                   % (msa, tmp_path / "small.eds", tmp_path / "small.seds"))
    assert err.startswith("[Performance] Runtime: ")
    e0, s0 = (tmp_path / "small.eds").read_bytes(), (tmp_path / "small.seds").read_bytes()
    assert (e0, s0) == oracle_lib.msa2eds(SMALL, 0)
    rc, out, err = run("eds2leds", "-i", str(tmp_path / "small.eds"), "-s", str(tmp_path / "small.seds"), "-l", "10")
    assert rc == 0, err
    assert "  Output mode: compact\n  Threads: 1 (sequential)\n" in out and out.endswith("Transformation complete!\n")
    assert (tmp_path / "small_l10.leds").read_bytes() == b"AGTC{TCTATAAATAAATA,TATATAAATGGATA,CCTATATAAATAAATAGGGG}\n"
    assert (tmp_path / "small_l10.seds").read_bytes() == b"{0}{1}{3}{2}\n"
    # direct MSA -> l-EDS, default names <stem>_l<N>.leds / .seds
    rc, out, err = run("msa2eds", "-i", str(msa), "-l", "4")
    assert rc == 0 and out.startswith("MSA → l-EDS transformation (l=4)\n")
    assert ((tmp_path / "small_l4.leds").read_bytes(), (tmp_path / "small_l4.seds").read_bytes()) == oracle_lib.msa2eds(SMALL, 4)
    # CARTESIAN, full format, custom output, threads reported as given
    rc, out, err = run("eds2leds", "-i", str(tmp_path / "small.eds"), "-l", "2", "--full", "-o", str(tmp_path / "c.leds"), "-t", "8")
    assert rc == 0 and "  Output mode: full\n  Threads: 8 (parallel)\n" in out
    assert (tmp_path / "c.leds").read_bytes() == oracle_lib.eds2leds(e0, None, 2, compact=False)[0]
    # library errors surface as "Error: <what>" with exit code 1
    (tmp_path / "bad.eds").write_bytes(b"{AAAA}{A,C}{G}{T,G}{TTTT}")
    (tmp_path / "bad.seds").write_bytes(b"{0}{1}{2}{0}{3}{4}{0}")
    rc, out, err = run("eds2leds", "-i", str(tmp_path / "bad.eds"), "-s", str(tmp_path / "bad.seds"), "-l", "2")
    assert rc == 1 and err.startswith("Error: Merging positions 1 and 2 results in empty set (no valid source intersections)\n")


@pytest.mark.gpu
def test_vcf2eds_tool(tmp_path):
    import vcf_checks

    named = {(c["name"], c["l"]): c for c in vcf_checks.golden_cases() if c.get("name")}
    c0 = named[("small", 0)]
    (tmp_path / "small.vcf").write_bytes(c0["vcf"].encode("latin-1"))
    (tmp_path / "small.fa").write_bytes(c0["fa"].encode("latin-1"))
    rc, out, err = run("vcf2eds", "-i", str(tmp_path / "small.vcf"), "-r", str(tmp_path / "small.fa"))
    assert rc == 0, err
    assert out.startswith("VCF → EDS transformation\n  Input: \"%s\"\n  Reference: \"%s\"\nTransformation complete!\n"
                          % (tmp_path / "small.vcf", tmp_path / "small.fa"))
    assert ("Variant Processing Statistics:\n  Total variants read:        10\n  Successfully processed:     10\n"
            "  Skipped (malformed):        0\n  Skipped (unsupported SV):   0\n  Total skipped:              0\n"
            "  Variant groups created:     10\n  Success rate:               100.0%\n") in out
    assert (tmp_path / "small.eds").read_bytes() == c0["eds"].encode("latin-1")
    assert (tmp_path / "small.seds").read_bytes() == c0["seds"].encode("latin-1")
    c3 = named[("small", 3)]
    rc, out, err = run("vcf2eds", "-i", str(tmp_path / "small.vcf"), "-r", str(tmp_path / "small.fa"), "-l", "3")
    assert rc == 0 and out.startswith("VCF → l-EDS transformation (l=3)\n  Using two-stage pipeline: VCF→EDS→l-EDS\n")
    assert (tmp_path / "small_l3.leds").read_bytes() == c3["eds"].encode("latin-1")
    assert (tmp_path / "small_l3.seds").read_bytes() == c3["seds"].encode("latin-1")
    # unsupported symbolic allele: the reference's warning on stderr, the record counted and skipped
    (tmp_path / "sv.vcf").write_bytes(vcf_checks.HDR + b"chr1\t5\t.\tA\tT,<DUP>\t.\t.\t.\tGT\t1|0\t0|0\t0|0\n"
                                      b"chr1\t9\t.\tA\tT\t.\t.\t.\tGT\t1|0\t0|0\t0|0\n")
    (tmp_path / "sv.fa").write_bytes(vcf_checks.FA)
    rc, out, err = run("vcf2eds", "-i", str(tmp_path / "sv.vcf"), "-r", str(tmp_path / "sv.fa"), "-o", str(tmp_path / "o.eds"))
    assert rc == 0 and err.startswith("Warning: Skipping variant at chr1:5 - Unsupported structural variant type: DUP\n")
    assert "  Skipped (unsupported SV):   1\n" in out and "  Success rate:               50.0%\n" in out
    exp = oracle_lib.vcf2eds((tmp_path / "sv.vcf").read_bytes(), vcf_checks.FA, 0)
    assert ((tmp_path / "o.eds").read_bytes(), (tmp_path / "o.seds").read_bytes()) == exp[:2]
    assert exp[3] == ["Warning: Skipping variant at chr1:5 - Unsupported structural variant type: DUP"]
    # library errors: "Error: <what>", exit code 1
    (tmp_path / "bad.fa").write_bytes(b"ACGT\n")
    rc, out, err = run("vcf2eds", "-i", str(tmp_path / "sv.vcf"), "-r", str(tmp_path / "bad.fa"))
    assert rc == 1 and err.startswith("Error: Invalid FASTA format: expected header line starting with '>'\n")


@pytest.mark.gpu
def test_is_leds():
    import edsparser_b200

    ctx = edsparser_b200.load().context(0)
    try:
        assert ctx.is_leds(b"{AAAA}{C,G}{TTTT}", 3)
        assert not ctx.is_leds(b"{AAAA}{C,G}{TT}{A,C}{GGGG}", 3)
        assert ctx.is_leds(b"{AAAA}{C,G}{TT}{A,C}{GGGG}", 2)
        assert not ctx.is_leds(b"{A,C}{G,T}", 1)
        assert ctx.is_leds(b"{A,C}{G,T}", 0) and ctx.is_leds(b"{AC}", 5) and ctx.is_leds(b"{A}{C,G}", 9)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_msa2eds_gpus_option(tmp_path):
    """msa2eds --gpus N: one .eds/.seds pair written from N column shards == the oracle (N = the box's devices, at most 4;
    on a one-GPU box the option still goes through the group entry point with one device)."""
    import torch
    from edsparser_b200 import synth

    text = synth.fasta_window(60, 50_000, 70, seed=3, variable_ppm=20_000)
    src = tmp_path / "shape.msa"
    src.write_bytes(text)
    exp = oracle_lib.msa2eds(text, 10)
    for n in sorted({1, min(2, torch.cuda.device_count()), min(4, torch.cuda.device_count())}):
        out = tmp_path / f"o{n}.leds"
        rc, so, err = run("msa2eds", "-i", str(src), "-l", "10", "-o", str(out), "--gpus", str(n))
        assert rc == 0, err
        assert out.read_bytes() == exp[0]
        assert (tmp_path / "shape_l10.seds").read_bytes() == exp[1]
    rc, so, err = run("msa2eds", "-i", str(src), "--gpus", "0")
    assert rc == 1 and "--gpus" in err


@pytest.mark.gpu
def test_vcf2eds_gpus_option(tmp_path):
    """vcf2eds --gpus N: slices of the record lines over N devices == the one-device run (N = the box's devices, at most
    4; on a one-GPU box the option still goes through the group entry point with one device)."""
    import torch
    import vcf_checks

    vcf, fa = vcf_checks.synth_vcf(300_000, 8000, 50, seed=4)
    (tmp_path / "s.vcf").write_bytes(vcf)
    (tmp_path / "s.fa").write_bytes(fa)
    for l in (0, 5):
        ext = "leds" if l else "eds"
        rc, so, err = run("vcf2eds", "-i", str(tmp_path / "s.vcf"), "-r", str(tmp_path / "s.fa"), "-l", str(l), "-o", str(tmp_path / f"one.{ext}"))
        assert rc == 0, err
        for n in sorted({1, min(2, torch.cuda.device_count()), min(4, torch.cuda.device_count())}):
            rc, so, err = run("vcf2eds", "-i", str(tmp_path / "s.vcf"), "-r", str(tmp_path / "s.fa"), "-l", str(l), "-o",
                              str(tmp_path / f"g{n}.{ext}"), "--gpus", str(n))
            assert rc == 0, err
            assert (tmp_path / f"g{n}.{ext}").read_bytes() == (tmp_path / f"one.{ext}").read_bytes()
            assert (tmp_path / f"g{n}.seds").read_bytes() == (tmp_path / "one.seds").read_bytes()
    rc, so, err = run("vcf2eds", "-i", str(tmp_path / "s.vcf"), "-r", str(tmp_path / "s.fa"), "--gpus", "0")
    assert rc == 1 and "--gpus" in err
