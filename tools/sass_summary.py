"""Per-kernel SASS evidence of the shipped library: python tools/sass_summary.py > profiles/sass_summary.txt
Counts of the Blackwell tensor / TMA mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP,
mma.sync -> HMMA, ldmatrix -> LDSM, cp.async -> LDGSTS) in every kernel of cultionet_b200/libcultionet_b200.so that has any."""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parent.parent / "cultionet_b200" / "libcultionet_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA\b"), ("LDTM", r"\bLDTM\b"), ("UTMALDG", r"\bUTMALDG\b"), ("UBLKCP", r"\bUBLKCP\b"),
                               ("HMMA", r"\bHMMA\b"), ("LDSM", r"\bLDSM\b"), ("LDGSTS", r"\bLDGSTS\b"), ("RED/ATOM", r"\b(RED|ATOM|ATOMG|ATOMS)\b")])
cur, rows = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur:
        for k, p in PAT.items():
            if re.search(p, line):
                rows[cur][k] += 1
arch = re.search(r"arch = (sm_\w+)", txt)
print(f"# cuobjdump -sass {so.name} ({arch.group(1) if arch else '?'}): kernels with tensor-core / TMA / async-copy instructions")
print(f"# {'kernel':100s} " + " ".join(f"{k:>8s}" for k in PAT))
for fn, c in rows.items():
    if not any(c[k] for k in ("UTC*MMA", "LDTM", "UTMALDG", "UBLKCP", "HMMA", "LDSM")):
        continue
    full = demangle(fn).replace("void ", "")
    cut = full.find(">(")
    name = (full[:cut + 1] if cut >= 0 else full.split("(")[0]).replace("(int)", "").replace("(bool)", "")
    print(f"{name[:102]:102s} " + " ".join(f"{c[k]:8d}" for k in PAT))
print(f"# {len(rows)} kernels in the library")
