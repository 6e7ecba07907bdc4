"""GPU box: the reference's main() at cfg3 -- setup, 1000 IMEX steps, 101 time stamps written as .vtu (110 MB per stamp)
-- end to end, with the output on tmpfs and on the box's disk.

    python scripts/run_full_system_cfg3.py [--g 7] [--stamps 100] [--dir /dev/shm]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pecs_b200 as pecs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--g", type=int, default=7)
    ap.add_argument("--stamps", type=int, default=100)
    ap.add_argument("--dir", default="/dev/shm")
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="pecs_cfg3_", dir=a.dir)
    try:
        prob = pecs.SolarCellProblem(pecs.default_input_file(a.g, 1, computational__time_stamps=a.stamps))
        prob.set_output(tmp)
        t0 = time.perf_counter()
        prob.run_full_system()
        wall = time.perf_counter() - t0
        files = os.listdir(tmp)
        size = sum(os.path.getsize(os.path.join(tmp, f)) for f in files)
        ms = prob.step_timed(100)[0] / 100
        prob.close()
        print(json.dumps({"g": a.g, "time_stamps": a.stamps, "directory": a.dir, "run_full_system_wall_seconds": wall,
                          "files": len(files), "gigabytes_written": size / 1e9, "ms_per_step_time_loop_only": ms,
                          "seconds_1000_steps_alone": 1000 * ms / 1e3}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
