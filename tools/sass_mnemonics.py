"""Instruction counts per kernel from `cuobjdump -sass libx3d2c.so` (stdin): which kernels use TMA, mbarriers, cp.async,
128-bit shared accesses. usage: cuobjdump -sass x3d2_b200/libx3d2c.so | python tools/sass_mnemonics.py"""
import collections
import re
import subprocess
import sys

WANT = ["UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "LDS.128", "STS.128", "LDS.64", "STS.64",
        "LDG", "STG", "BAR.SYNC", "MUFU.RCP64H"]
KEYS = ["transeq_m4_kernel", "transeq_m4i_kernel", "tds_m4_kernel", "tds_m4i_kernel", "tds_g_kernel",
        "spectral_010_kernel", "sum2_lincomb_kernel", "chunk_exchange_kernel", "halo_pack_kernel",
        "process_spectral_000"]

cur, cnt = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        for w in WANT:
            if m.group(1).startswith(w):
                cnt[cur][w] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass x3d2_b200/libx3d2c.so (sm_100a): instruction counts of the kernels named in DESIGN.md")
print("# UTMALDG / UTMASTG = tensor (TMA) loads / stores, SYNCS = mbarrier operations, LDGSTS = cp.async,")
print("# LDG / STG inside the *_m4i kernels = the system-scope carry polls / pushes of the in-kernel exchange")
for (f, c), d in zip(cnt.items(), names):
    if not any(k in d for k in KEYS):
        continue
    d = d.replace("(anonymous namespace)::", "")
    d = re.sub(r"_GLOBAL__N__\w+::", "", d)
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print("%-64s %s" % (d[:64], " ".join("%s=%d" % (k, v) for k, v in c.items() if v)))
