"""Static SASS mnemonic counts per kernel of libmmidx.so (sm_100a): python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "multimedia-indexing_b200", "libmmidx.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["HMMA", "LDSM", "LDGSTS", "UBLKCP", "IDP.4A", "FADD2", "DADD", "DMUL", "DFMA", "FFMA", "LDS", "ATOMS", "SYNCS", "LDG", "STG", "RED",
         "BAR.SYNC", "BAR.RED", "SHFL", "MATCH"]
print("# SASS mnemonic counts of libmmidx.so (sm_100a), `cuobjdump -sass`, per kernel (static instruction counts)")
print("# HMMA = tensor-core mma.sync (k_coarse_mma: the coarse-quantizer GEMM), LDSM = ldmatrix, LDGSTS = cp.async; UBLKCP = TMA bulk copy")
print("# (cp.async.bulk) of the T1 rows / ADC tables; IDP.4A + LDS + FADD2 = the ADC lookup / packed fp32 add of the fused scan;")
print("# DADD/DMUL = the binary64 reference arithmetic (DFMA only where the value is a filter input, never on a result path)")
print()
name, cnt = None, collections.Counter()


def flush():
    if name:
        short = re.sub(r"^_ZN5mmidx\d+", "", name)
        print(f"{short[:72]:72s} " + " ".join(f"{k}={cnt[k]}" for k in WATCH if cnt[k]))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, cnt = m.group(1), collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for k in WATCH:
            if op == k or op.startswith(k + "."):
                cnt[k] += 1
                break
flush()
