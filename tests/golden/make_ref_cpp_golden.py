#!/usr/bin/env python
"""Which test functions of the reference's own tests/cpp/*.cpp pass against the UNMODIFIED reference library
(tests/_refbin/*_ref, built by `make reftests` where /root/reference is mounted). Some of the reference's asserts do
not hold for the reference itself (its CMake Release build defines NDEBUG); the ones that do are what this repo's
host layer must pass too. Writes tests/golden/ref_cpp_tests.json."""
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/tests/cpp"
out = {}
for name in ("test_msa", "test_merge", "test_sources", "test_stats"):
    src = open(os.path.join(REF, name + ".cpp")).read()
    funcs = re.findall(r"^void (test_\w+)\(\)\s*\{", src, flags=re.M)
    p = subprocess.run([os.path.join(ROOT, "tests", "_refbin", name + "_ref"), *funcs], capture_output=True, text=True, cwd="/tmp")
    verdict = dict(line.split() for line in p.stdout.splitlines() if line.strip())
    out[name] = {"functions": funcs, "pass_on_reference": [f for f in funcs if verdict.get(f) == "PASS"]}
    print(name, len(funcs), "functions,", len(out[name]["pass_on_reference"]), "pass on the reference:", [f for f in funcs if verdict.get(f) != "PASS"], "do not")
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ref_cpp_tests.json"), "w"), indent=1)
