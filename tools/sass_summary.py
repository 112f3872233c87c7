"""Opcode histogram per kernel of the built library (cuobjdump -sass): the instruction-level evidence that the hot kernels
use the Blackwell paths they claim (UTCHMMA / UTCBAR = tcgen05.mma / commit, UTMALDG / UTMASTG = TMA, UBLKCP = bulk copies,
SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed fp32).     python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio-to-motion-generation_b200", "liba2m_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op, mods = m.group(1), m.group(2)
        key = op
        if op in ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTCATOMSWS"):
            key = op + mods
        kernels[cur][key] += 1
demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
HOT = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDTM", "STTM", "FFMA2", "FADD2", "FMUL2", "HMMA", "MUFU",
       "BAR", "LDS", "STS", "LDG", "STG", "SHFL", "FFMA", "FADD", "FMUL", "DFMA", "DADD", "DMUL")
print("# cuobjdump -sass %s: static instruction counts per kernel (tools/sass_summary.py)" % os.path.relpath(LIB, ROOT))
print("# tensor core = UTCHMMA (tcgen05.mma kind::f16; .2CTA = cta_group::2) + UTCBAR (tcgen05.commit; .2CTA.MULTICAST); UTCATOMSWS = tcgen05.alloc /")
print("# dealloc; tensor memory = LDTM / STTM (tcgen05.ld / st); TMA = UTMALDG /")
print("# UTMASTG (tensor) and UBLKCP (bulk); mbarrier = SYNCS; packed fp32 = FFMA2 / FADD2 / FMUL2; HMMA would mean legacy mma.sync (none).")
for (name, ctr), dm in zip(kernels.items(), demangle):
    cut = dm.rfind(">(")
    short = dm[:cut + 1] if cut >= 0 else dm.split("(")[0]
    short = short.replace("(int)", "").replace("(bool)", "")
    short = re.sub(r"^void ", "", short)
    short = short.replace("a2m::(anonymous namespace)::", "").replace("a2m::", "")
    total = sum(ctr.values())
    hot = []
    for h in HOT:
        n = sum(v for k, v in ctr.items() if k == h or k.startswith(h + "."))
        if n:
            hot.append("%s=%d" % (h, n))
    print("%-70s %6d instr  %s" % (short[:70], total, " ".join(hot)))
    detail = ["%s=%d" % (k, v) for k, v in sorted(ctr.items()) if k.split(".")[0] in ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP") and "." in k]
    if detail:
        print("%-70s %s" % ("", " ".join(detail)))
