"""Per-kernel SASS / resource summary of pecs_b200/lib/libpecs_b200.so (runs in the build container: cuobjdump only).
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
Counts the instructions that show what the kernels are built on: UBLKCP (bulk asynchronous copy engine, cp.async.bulk),
SYNCS (mbarrier), 256-bit global loads / stores, LDGSTS (cp.async), fp64 arithmetic, release atomics (MEMBAR + ATOM) / strong loads (ld.acquire.gpu and ld.global.cg: LDG.E...STRONG.GPU)
of the dataflow counters, griddepcontrol (ACQBULK / launch-dependents); registers, stack and spills from --dump-resource-usage.
There is no UTMALDG / UTC*MMA: the path is fp64, GEMV-shaped and HBM-bound, tensor cores do not apply."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pecs_b200", "lib", "libpecs_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731

usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        usage[cur] = dict(re.findall(r"(\w+):(\d+)", line))
        cur = None

PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDG.256", r"\bLDG\.E\.\S*256"), ("STG.256", r"\bSTG\.E\.\S*256"),
            ("LDG.nc", r"\bLDG\.E\.\S*CONSTANT"), ("LDG.strong", r"\bLDG\.E\.\S*STRONG\.GPU"), ("LDG.cg", r"\bLDG\.E\.\S*\.LTC"),
            ("LDGSTS", r"\bLDGSTS"), ("RED", r"\bRED(UX)?\.E"), ("ATOM", r"\bATOM"), ("MEMBAR", r"\bMEMBAR"),
            ("ACQBULK", r"\bACQBULK"), ("DFMA", r"\bDFMA"), ("DMUL", r"\bDMUL"), ("DADD", r"\bDADD"), ("MUFU", r"\bMUFU"),
            ("STL", r"\bSTL"), ("LDL", r"\bLDL"), ("UTMALDG", r"\bUTMALDG"), ("UTCMMA", r"\bUTC\w*MMA")]
kernels = {}
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        kernels[name] = {k: 0 for k, _ in PATTERNS}
        kernels[name]["instructions"] = 0
        continue
    if name and re.search(r"/\*[0-9a-f]{4}\*/", line):
        kernels[name]["instructions"] += 1
        for k, pat in PATTERNS:
            if re.search(pat, line):
                kernels[name][k] += 1

arch = re.search(r"arch = (sm_\w+)", sass)
print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, {arch.group(1) if arch else '?'}")
print("# columns: REG / STACK bytes / SHARED static bytes | instruction counts")
cols = ["instructions"] + [k for k, _ in PATTERNS]
total = {c: 0 for c in cols}
for name in sorted(kernels, key=demangle):
    d = kernels[name]
    short = re.sub(r"\(anonymous namespace\)::|pecs::", "", demangle(name))
    short = re.sub(r"\(.*", "", short)
    u = usage.get(name, {})
    counts = " ".join(f"{c}={d[c]}" for c in cols if d[c])
    print(f"{short:70s} REG={u.get('REG', '?'):>3s} STACK={u.get('STACK', '?'):>4s} SHARED={u.get('SHARED', '?'):>5s} | {counts}")
    for c in cols:
        total[c] += d[c]
print("# totals: " + " ".join(f"{c}={total[c]}" for c in cols))
sys.exit(0)
