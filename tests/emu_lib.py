"""Loader of the TEST-ONLY emulator build (tests/emu/libedsparser_emu.so): the kernel sources of
edsparser_b200/csrc compiled with g++ over tests/emu/cuda_emu.h. Never used by the product."""
import glob
import os
import subprocess

from edsparser_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SO = os.path.join(ROOT, "tests", "emu", "libedsparser_emu.so")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        srcs = glob.glob(os.path.join(ROOT, "edsparser_b200", "csrc", "*")) + glob.glob(
            os.path.join(ROOT, "tests", "emu", "cuda_emu.*")) + [os.path.join(ROOT, "include", "edsparser_b200.h")]
        if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["make", "-C", ROOT, "emu"], stdout=subprocess.DEVNULL)
        _LIB = capi.Library(EMU_SO)
    return _LIB
