"""The reference's own C++ tests (tests/cpp/test_msa.cpp, test_merge.cpp, test_sources.cpp, test_stats.cpp), compiled
UNMODIFIED against this repo's host layer (class EDS, transforms API) by `make reftests` — binaries only, under the
git-ignored tests/_refbin/, built where /root/reference is mounted and carried to the GPU box. Every test function that
passes against the unmodified reference library (tests/golden/ref_cpp_tests.json, made by
tests/golden/make_ref_cpp_golden.py) must pass here. CPU tier: over the kernel-logic emulator; GPU tier: over
libedsparser_b200.so."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpp_tests.json")))


def run_suite(flavour, names):
    checked = 0
    for name in names:
        exe = os.path.join(ROOT, "tests", "_refbin", f"{name}_{flavour}")
        if not os.path.exists(exe):
            pytest.skip(f"{exe} is not built (needs the reference sources at build time)")
        want = GOLDEN[name]["pass_on_reference"]
        p = subprocess.run([exe, *want], capture_output=True, text=True, cwd=os.path.join(ROOT, "tests"), timeout=1500)
        verdict = dict(line.split() for line in p.stdout.splitlines() if line.strip())
        failed = [f for f in want if verdict.get(f) != "PASS"]
        assert not failed, (name, failed, p.stderr[-500:])
        checked += len(want)
    return checked


def test_reference_cpp_tests_emulated():
    # test_msa and test_merge (the transforms and EDS::merge_adjacent) here; the GPU tier runs all four files
    assert run_suite("emu", ["test_msa", "test_merge"]) >= 25


@pytest.mark.gpu
def test_reference_cpp_tests_on_gpu():
    assert run_suite("gpu", ["test_msa", "test_merge", "test_sources", "test_stats"]) >= 49
