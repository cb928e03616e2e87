#!/usr/bin/env python
"""Times (and, under ncu, profiles) the two smaller kernels of the hot path on synthetic input:
the sub-read scorer (GCUPS) and the k-mer candidate search (sub-reads/s, index lookups/s)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ngmlr_b200 import B200Aligner, refindex, synth  # noqa: E402


def main():
    genome_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    rng = np.random.default_rng(3)
    g = synth.random_genome(int(genome_mb * 1e6), 7)
    t0 = time.time()
    ref = refindex.encode_reference([g])
    idx = refindex.build_index(ref)
    t_index = time.time() - t0
    al = B200Aligner(0)
    al.set_index(idx)
    # sub-reads: 256-bp pieces of 8 kb PacBio-like reads
    subs = []
    for _ in range(n_reads):
        L = 8000
        s = int(rng.integers(0, g.size - L - 1))
        read, _m = synth.mutate(g[s:s + L], rng, err=0.15)
        for k in range(0, read.size - 256, 256):
            subs.append(read[k:k + 256].tobytes())
    al.cs_search(subs[:1000])
    t0 = time.time()
    cands, mx = al.cs_search(subs)
    t_cs = time.time() - t0
    n_cand = sum(len(c) for c in cands)
    top_ok = sum(1 for c in cands if c)
    # candidate windows -> sub-read scoring, 1024-pair batches like ScoreBuffer
    refs, qrys = [], []
    start = ref.ref_start[0]
    for sub, cl in zip(subs[:20000], cands[:20000]):
        for score, loc, rev in cl[:2]:
            p = max(loc - 20 - start, 0)
            refs.append(g[p:p + 306].tobytes())
            qrys.append(sub if not rev else synth.revcomp(np.frombuffer(sub, np.uint8)).tobytes())
    al.BatchScore(refs[:1024], qrys[:1024])
    t0 = time.time()
    ms = 0.0
    for k in range(0, len(refs), 1024):
        al.BatchScore(refs[k:k + 1024], qrys[k:k + 1024])
        ms += al.sw_kernel_ms()
    t_sw = time.time() - t0
    cells = sum((len(r) + 1) * (len(q) + 1) for r, q in zip(refs, qrys))
    print(json.dumps({
        "genome_mb": genome_mb, "index_build_s": t_index, "index_positions": int(idx.pos.size),
        "cs": {"subreads": len(subs), "wall_s": t_cs, "subreads_per_s": len(subs) / t_cs,
               "kmer_lookups_per_s": len(subs) * 244 * 2 / t_cs, "candidates": n_cand,
               "subreads_with_candidates": top_ok},
        "sw": {"pairs": len(refs), "wall_s": t_sw, "kernel_ms": ms, "gcups_kernel": cells / (ms * 1e-3) / 1e9,
               "gcups_wall": cells / t_sw / 1e9}}))
    al.close()


if __name__ == "__main__":
    main()
