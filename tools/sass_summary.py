"""Blackwell-native evidence: per-kernel counts of the sm_100a-specific SASS opcodes in the shipped library.

    python tools/sass_summary.py [slime_b200/libslime_b200.so] > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = TMA tensor load
(cp.async.bulk.tensor), UTCBAR = tcgen05.commit, SYNCS = mbarrier, UBLKCP = cp.async.bulk, MUFU.EX2 / FFMA2 = the
softmax arithmetic, HMMA = mma.sync (decode-step kernels only), R2UR counts the register -> uniform-register moves in
front of the tensor-core instructions (the warp-uniform issue loops keep descriptors in uniform registers)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "slime_b200/libslime_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = [("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("UTCHMMA", r"UTCHMMA(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
       ("UTMALDG", r"UTMALDG"), ("UTMALDG.2CTA", r"UTMALDG[.\w]*\.2CTA"), ("UTMASTG", r"UTMASTG"), ("UTCBAR", r"UTCBAR"),
       ("UTCBAR.MULTICAST", r"UTCBAR[.\w]*MULTICAST"), ("SYNCS", r"\bSYNCS"), ("UBLKCP", r"UBLKCP"), ("MUFU.EX2", r"MUFU\.EX2"),
       ("FFMA2", r"\bFFMA2"), ("FMNMX3", r"FMNMX3"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"LDGSTS"), ("R2UR", r"\bR2UR")]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(CUtensorMap_st.*", "", cur)
        cur = re.sub(r"\(.*", "", cur)[:70]
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None or "/*" not in line:
        continue
    for name, pat in PAT:
        if re.search(pat, line):
            counts[cur][name] += 1
tot = collections.Counter()
print(f"# {lib}: sm_100a SASS opcode counts per kernel (cuobjdump -sass); kernels without any of them omitted\n")
for k, c in counts.items():
    if not c:
        continue
    tot.update(c)
    print(f"{k:72s} " + "  ".join(f"{n}={v}" for n, v in c.items()))
print("\nTOTAL " + "  ".join(f"{n}={v}" for n, v in tot.items()))
