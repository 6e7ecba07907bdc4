"""CPU oracle of the PECS per-step IMEX path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
Nothing under pecs_b200/ imports, links or executes it.
"""
from .binding import Oracle, build, LIB_PATH  # noqa: F401
