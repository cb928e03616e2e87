"""Thread scaling of the CPU reference arm on this box (stage 4 only): which thread count is best?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from ngmlr_b200 import synth  # noqa: E402

g = synth.random_genome(5_000_000, 1)
pool = bench.make_pool(g, 512, seed=2)
kind = bench.cpu_impl_kind()
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "kind", kind)
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except OSError:
    pass
bench.cpu_reference_run(pool[:128], 128, kind, None)
for th in (1, 8, 16, 32, 64, 128):
    sub = pool[:max(16, min(len(pool), 4 * th))]
    t = bench.cpu_reference_run(sub, th, kind, None)
    c = sum(p.cells for p in sub)
    print(th, "threads %.2fs  total %.1f Mcells/s  per thread %.1f" % (t, c / t / 1e6, c / t / th / 1e6))
