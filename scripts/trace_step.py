"""GPU box, trace build only (python -m pecs_b200.build --variant trace; PECS_B200_LIB=pecs_b200/lib/libpecs_b200_trace.so):
per-block timestamps of the level kernels of (a) one Poisson solve alone, (b) one carrier solve alone, (c) one step.
Writes gpurun_out/trace_<tag>.npz with arrays [n, 4] = {tag << 32 | block, start, dependencies met, end} (ns, globaltimer)
and prints a per-launch summary.  Never a timing source for bench numbers."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import pecs_b200 as pecs  # noqa: E402
from pecs_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("-g", type=int, default=7)
ap.add_argument("--tag", default="trace")
a = ap.parse_args()
lib = _lib.load()
lib.pecs_trace_start.argtypes = [C.c_int]
lib.pecs_trace_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
CAP = 400000


def capture(fn):
    assert lib.pecs_trace_start(CAP) == 0
    fn()
    buf = np.zeros((CAP, 4), np.uint64)
    n = lib.pecs_trace_read(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), CAP)
    return buf[:n].copy()


def summary(name, rec):
    if len(rec) == 0:
        print(name, "no records")
        return
    tag = (rec[:, 0] >> np.uint64(32)).astype(np.int64)
    t0 = rec[:, 1].astype(np.int64)
    base = t0.min()
    print(f"== {name}: {len(rec)} blocks, span {(rec[:, 3].astype(np.int64).max() - base) / 1e3:.1f} us")
    order = sorted(set(tag.tolist()), key=lambda t: t0[tag == t].min())
    for t in order:
        m = tag == t
        s, d, e = t0[m] - base, rec[m, 2].astype(np.int64) - base, rec[m, 3].astype(np.int64) - base
        print(f"  tag {t:6d} blocks {m.sum():5d} first start {s.min() / 1e3:8.1f} last start {s.max() / 1e3:8.1f} "
              f"first end {e.min() / 1e3:8.1f} last end {e.max() / 1e3:8.1f} us | wait for deps: mean {(d - s).mean() / 1e3:6.2f} "
              f"max {(d - s).max() / 1e3:6.2f} | work: mean {(e - d).mean() / 1e3:6.2f} max {(e - d).max() / 1e3:6.2f}")


prob = pecs.SolarCellProblem(pecs.default_input_file(a.g, 1))
prob.setup_full_system()
prob.step(3)
prob.synchronize()
out = {}
for name, fn in (("poisson_solve", lambda: (prob.solve_Poisson(), prob.synchronize())),
                 ("electron_solve", lambda: (prob.solve_species(0), prob.synchronize())),
                 ("step", lambda: (prob.step(1), prob.synchronize()))):
    fn()  # warm
    rec = capture(fn)
    out[name] = rec
    summary(name, rec)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(f"gpurun_out/trace_{a.tag}.npz", **out)
prob.close()
