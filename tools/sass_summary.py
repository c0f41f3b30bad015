"""tools/sass_summary.py [lib.so ...] -- per-kernel SASS evidence for the claims in DESIGN.md section 4 (run here, no GPU):
counts of the mnemonics that prove the Blackwell-native paths (UBLKCP = cp.async.bulk / TMA engine, SYNCS = mbarrier,
LDG/STG .ENL2.256 = 256-bit accesses, REDG = fire-and-forget atomics, REDUX = warp reduce, ATOM.*.128 = 128-bit CAS) plus
registers / shared memory per kernel.  Output is committed under profiles/."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

PATTERNS = OrderedDict([
    ("UBLKCP (bulk async copy, TMA engine)", r"\bUBLKCP"),
    ("UBLKPF / prefetch", r"\bUBLKPF"),
    ("SYNCS (mbarrier)", r"\bSYNCS"),
    ("LDG.*.ENL2.256", r"\bLDG\.\S*ENL2\.256"),
    ("STG.*.ENL2.256", r"\bSTG\.\S*ENL2\.256"),
    ("LDG.*.128", r"\bLDG\.\S*\.128"),
    ("LDS.128 / STS.128", r"\b(LDS|STS)\.128"),
    ("REDG (no-return atomics)", r"\bREDG?\.E"),
    ("ATOM/ATOMG (returning atomics)", r"\bATOMG?\.E"),
    ("ATOM*.CAS.128", r"\bATOMG?\.E\.CAS\.128"),
    ("REDUX (warp reduce)", r"\bREDUX"),
    ("SHFL", r"\bSHFL"),
    ("NANOSLEEP", r"\bNANOSLEEP"),
    ("BAR.SYNC", r"\bBAR\.SYNC"),
    ("DADD/DMUL/DFMA/DSETP", r"\b(DADD|DMUL|DFMA|DSETP)"),
    ("UTC*MMA (tensor core; none expected)", r"\bUTC\w*MMA"),
])


def demangle_short(name):
    out = subprocess.run(["c++filt", "-p", name], capture_output=True, text=True).stdout.strip() or name
    out = re.sub(r"\(anonymous namespace\)::", "", out)
    out = re.sub(r"kb200::Impl::", "", out)
    return out[:150]


def main():
    for so in sys.argv[1:]:
        sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
        usage = {}
        cur = None
        for line in res.splitlines():
            m = re.match(r"\s*Function (\S+):", line)
            if m:
                cur = m.group(1)
            elif cur and "REG:" in line:
                usage[cur] = " ".join(re.findall(r"(?:REG|STACK|SHARED):\d+", line))
        print(f"==== {so}")
        kernels = OrderedDict()
        name = None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                name = m.group(1)
                kernels[name] = Counter()
                continue
            if name is None:
                continue
            for label, pat in PATTERNS.items():
                if re.search(pat, line):
                    kernels[name][label] += 1
        for k, c in kernels.items():
            if not c:
                continue
            print(f"-- {demangle_short(k)}\n   [{usage.get(k, '')}]  " + ", ".join(f"{lab}: {n}" for lab, n in c.items()))


if __name__ == "__main__":
    main()
