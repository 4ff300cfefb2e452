import os, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, edsparser_b200 as E, vcf_checks
vcf, fa = vcf_checks.synth_vcf(n_bases=5_000_000, n_sites=50_000, n_samples=2504, seed=1)
ctx = E.load().context(0)
pv = torch.empty(len(vcf), dtype=torch.uint8).pin_memory(); pv.copy_(torch.frombuffer(bytearray(vcf), dtype=torch.uint8))
pf = torch.empty(len(fa), dtype=torch.uint8).pin_memory(); pf.copy_(torch.frombuffer(bytearray(fa), dtype=torch.uint8))
for i in range(3):
    t=time.perf_counter(); ctx.vcf_transform_host_raw(pv.data_ptr(), len(vcf), pf.data_ptr(), len(fa), 0); print("call", (time.perf_counter()-t)*1e3, file=sys.stderr)
