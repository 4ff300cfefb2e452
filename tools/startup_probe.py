"""Where a tool's fixed second goes: library load, first context, first transform (tiny input)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.perf_counter()
rt = ctypes.CDLL("libcudart.so")
rt.cudaFree(None)
t1 = time.perf_counter()
import edsparser_b200 as E
lib = E.load()
t2 = time.perf_counter()
c = lib.context(0)
t3 = time.perf_counter()
c2 = lib.context(0)
t4 = time.perf_counter()
text = b">a\nACGTACGTAC\n>b\nACGTACCTAC\n>c\nACGAACGTAC\n"
c.msa_transform_host(text, 3)
t5 = time.perf_counter()
c.msa_transform_host(text, 3)
t6 = time.perf_counter()
c.leds_merge_host(b"{ACGT}{A,C}{GT}", b"{0}{1}{2}{0}", 3)
t7 = time.perf_counter()
print("cudaFree(0) %.3f | import+dlopen %.3f | first ctx %.3f | second ctx %.3f | first msa %.3f | second msa %.3f | first leds %.3f" % (
    t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6))
