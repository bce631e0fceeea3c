"""`cuobjdump -sass libgccvae.so` -> instruction counts per kernel (static code) for the opcodes that prove what a kernel
is built on: python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "semi-supervised-gated-lt-vae_b200", "csrc", "libgccvae.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "HMMA", "LDGSTS", "RED", "ATOMG", "MUFU"]
kern, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("gccvae::", "").replace("void ", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[kern][k] += 1
print("# cuobjdump -sass %s (sm_100a): instruction counts per kernel (static code, not executed counts)" % os.path.relpath(lib, ROOT))
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = bulk copy (cp.async.bulk), UTCBAR = tcgen05.commit,")
print("# HMMA = mma.sync (the fused dense-chain kernels: 8 images per CTA are below tcgen05's 64-row tiles), LDGSTS = cp.async,")
print("# RED / ATOMG = global reductions")
print("%-58s %6s  %s" % ("kernel", "SASS", "opcodes"))
for k, n in sorted(total.items(), key=lambda kv: -kv[1]):
    print("%-58s %6d  %s" % (k[:58], n, " ".join("%s=%d" % (o, counts[k][o]) for o in KEYS if counts[k][o])))
