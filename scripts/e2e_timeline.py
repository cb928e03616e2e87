"""Per-phase wall-clock timeline of the end-to-end step (host buffers -> results) for different ways
of driving the public API: S contexts x whole batch, or the batch split into S slices.
Usage: python scripts/e2e_timeline.py [reads] [steps]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ngmlr_b200 import B200Aligner, PackedBatch, PackedReads, refindex, split_read, synth  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = synth.random_genome(50_000_000, 1)
contigs = [g[i * 10_000_000:(i + 1) * 10_000_000] for i in range(5)]
ref = refindex.encode_reference(contigs)
idx = refindex.build_index(ref)
pool = synth.pacbio_problems(n_reads, seed=2, median=8000, genome=g)
bases = sum(len(p.qry) for p in pool)


def make(parts):
    out = []
    for j in range(parts):
        sl = pool[j::parts]
        out.append((PackedBatch.from_problems(sl), PackedReads([s for p in sl for s in split_read(p.qry)])))
    return out


def run(S, parts, label, grouped=False):
    als = []
    for _ in range(S):
        a = B200Aligner(0, stream=torch.cuda.Stream().cuda_stream)
        a.set_index(idx)
        a.set_reference(ref)
        als.append(a)
    data = make(parts)
    log = []

    def worker(j, items, rec):
        a = als[j]
        for (batch, subs) in items:
            t = [time.perf_counter()]
            if grouped:  # all H2D, then all kernels, then all D2H (names below keep the interleaved order)
                a.cs_upload(subs); t.append(time.perf_counter())
                a.upload(batch); t.append(time.perf_counter())
                a.cs_run(); t.append(time.perf_counter())
                a.run(); t.append(time.perf_counter())
                a.cs_fetch(); t.append(time.perf_counter())
                r = a.fetch(); t.append(time.perf_counter())
            else:
                a.cs_upload(subs); t.append(time.perf_counter())
                a.cs_run(); t.append(time.perf_counter())
                a.cs_fetch(); t.append(time.perf_counter())
                a.upload(batch); t.append(time.perf_counter())
                a.run(); t.append(time.perf_counter())
                r = a.fetch(); t.append(time.perf_counter())
            assert len(r) == batch.n
            if rec:
                log.append((j, t))

    def go(k, rec):
        # k steps; a step = every slice once. Work items are dealt round-robin to the S threads.
        items = [data[i % parts] for i in range(k * parts)]
        per = [items[j::S] for j in range(S)]
        ts = [threading.Thread(target=worker, args=(j, per[j], rec)) for j in range(S)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        return t0, time.perf_counter() - t0

    go(1, False)
    go(1, False)
    t0, dt = go(steps, True)
    print("%s grouped=%d: S=%d parts=%d  %.1f ms/step  %.3f Gbp/s" % (label, grouped, S, parts, dt * 1e3 / steps, bases * steps / dt / 1e9))
    names = ["cs_up", "upload", "cs_run", "run", "cs_fetch", "fetch"] if grouped else ["cs_up", "cs_run", "cs_fetch", "upload", "run", "fetch"]
    tot = np.zeros(6)
    for j, t in log:
        tot += np.diff(t)
    print("   mean per item (ms):", " ".join("%s %.1f" % (n, v * 1e3 / len(log)) for n, v in zip(names, tot)))
    if os.environ.get("TIMELINE"):
        for j, t in sorted(log, key=lambda x: x[1][0]):
            print("   thr %d start %.1f :" % (j, (t[0] - t0) * 1e3), " ".join("%.1f" % ((b - a) * 1e3) for a, b in zip(t, t[1:])))
    for a in als:
        a.close()


for S, parts, grouped in ((1, 1, False), (2, 1, False), (2, 1, True), (3, 1, True), (2, 2, True), (4, 4, True)):
    run(S, parts, "cfg", grouped)
