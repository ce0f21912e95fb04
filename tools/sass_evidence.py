"""Count the SASS mnemonics that prove tcgen05 / TMEM / TMA / mbarrier use, per kernel of the built library, plus each
kernel's register count and static shared memory (cuobjdump; no GPU needed).
    python tools/sass_evidence.py > profiles/r02_sass_evidence.md"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "ct-clip-ut_b200" / "ctclip_b200" / "libctclip_b200.so"
PAT = OrderedDict([("UTC*MMA (tcgen05.mma)", r"\bUTC\w*MMA"), ("UTC*MMA.2CTA (cta_group::2)", r"\bUTC\w*MMA\.2CTA"),
                   ("UTMALDG.*.2CTA", r"\bUTMALDG\.\w+\.2CTA"), ("UTCBAR.2CTA.MULTICAST", r"\bUTCBAR\.2CTA\.MULTICAST"), ("UTCBAR / UTCCP (tcgen05 commit / cp)", r"\bUTC(BAR|CP)"),
                   ("LDTM / STTM (tcgen05.ld / st)", r"\b(LDTM|STTM)"), ("UTMALDG / UTMASTG / UBLKCP (TMA)", r"\b(UTMALDG|UTMASTG|UBLKCP)"),
                   ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("HMMA (mma.sync)", r"\bHMMA"),
                   ("MUFU.EX2", r"\bMUFU\.EX2")])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", n).replace("ctc::", "").replace("void ", "") for n in out]


sass = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True).stdout
kernels = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), {k: 0 for k in PAT})
        continue
    if cur is not None:
        for k, p in PAT.items():
            if re.search(p, line):
                cur[k] += 1
res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(SO)], capture_output=True, text=True).stdout
usage = {}
fn = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and fn:
        usage[fn] = (int(m.group(1)), int(m.group(2)))
names = list(kernels)
pretty = dict(zip(names, demangle(names)))
print("# SASS evidence (`cuobjdump -sass` / `--dump-resource-usage` of `libctclip_b200.so`, sm_100a)\n")
print("Counts of instruction mnemonics per kernel; `tcgen05.mma` shows up as `UTC*MMA`, `tcgen05.ld/st` as `LDTM/STTM`, TMA as\n"
      "`UTMALDG` / `UBLKCP`, mbarrier operations as `SYNCS` (B200_PROFILING.md).  Kernels without any of them are omitted from\n"
      "the first table.\n")
cols = list(PAT)
print("| kernel | regs | static smem | " + " | ".join(cols) + " |")
print("|---|---|---|" + "---|" * len(cols))
rest = []
for n in names:
    c = kernels[n]
    r, s = usage.get(n, ("?", "?"))
    if any(c[k] for k in cols[:9]):
        print(f"| `{pretty[n]}` | {r} | {s} | " + " | ".join(str(c[k]) for k in cols) + " |")
    else:
        rest.append((pretty[n], r, s, c["HMMA (mma.sync)"], c["MUFU.EX2"]))
print("\nOther kernels (no tcgen05 / TMA / cp.async): name, regs, static smem, HMMA, MUFU.EX2\n")
for p, r, s, h, e in sorted(rest):
    print(f"* `{p}` — {r} regs, {s} B, HMMA {h}, EX2 {e}")
