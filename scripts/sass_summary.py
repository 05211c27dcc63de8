"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the shipped library uses
(tcgen05 = UTCHMMA / LDTM / UTCBAR, TMA = UTMALDG / UBLKCP, packed FP32 = FFMA2, vector reductions = RED...128).
Needs cuobjdump only (no GPU):   python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "understanding_flow_robustness_b200", "libb200corr.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "LDS.128",
        "STG.E.ENL2.256", "LDG.E.NA.ENL2.256", "RED.E.ADD.F32", "REDG.E.ADD.F32x4", "LDGSTS", "SHFL", "BAR.SYNC"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            c = kernels[cur]
            c["total"] += 1
            for k in ("UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "LDGSTS", "SHFL", "BAR"):
                if op.startswith(k):
                    c[k] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                c["UTCHMMA.2CTA"] += 1
            if op == "FFMA" or op.startswith("FFMA."):
                c["FFMA"] += 1
            if op.startswith("LDS") and ".128" in op:
                c["LDS.128"] += 1
            if op.startswith("STG") and ".256" in op:
                c["STG.256"] += 1
            if op.startswith("LDG") and ".256" in op:
                c["LDG.256"] += 1
            if op.startswith("RED") or op.startswith("ATOM"):
                c["RED/ATOM"] += 1
                if ".128" in op or "x4" in op:
                    c["RED.128"] += 1
    cols = ["total", "UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "LDS.128", "LDG.256",
            "STG.256", "RED/ATOM", "RED.128", "LDGSTS", "SHFL", "BAR"]
    print("# cuobjdump -sass understanding_flow_robustness_b200/libb200corr.so  (sm_100a), instruction counts per kernel")
    print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA load),")
    print("# SYNCS = mbarrier ops, FFMA2 = packed fp32x2 FMA, RED.128 = red.global.add.v4.f32")
    print("kernel".ljust(64) + "".join(c.rjust(13) for c in cols))
    for name, c in kernels.items():
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", dem)
        dem = re.sub(r"\(.*", "", dem)[:62]
        print(dem.ljust(64) + "".join(str(c.get(k, 0)).rjust(13) for k in cols))


if __name__ == "__main__":
    sys.exit(main())
