#!/usr/bin/env python
"""Targeted fuzz of the one corner the CUDA RAW fill does not replay literally: the
reference's SSE blocks first compute EVERY cell of a row -- including its last <= 12 columns -- and feed those
values to the best-cell tracking (src/ConvexAlignFast.cpp:1165-1170); the scalar tail then recomputes the last
<= 12 columns with the scalar rule and tracks again (:1179, :1270-1275). The oracle's rule 2 / the RAW kernel
track only the final value of every cell. A difference needs a scoring outside the default class (where SSE and
scalar values of a cell can differ at all) AND a tail cell whose SSE value beats every cell seen so far while
its final value does not.

Compares the COMPILED REFERENCE's fill (oracle/_ref/libngmlr_ref.so: directions, best score bits, best cell)
with the oracle's rule 2 on inputs built to hit the corner: corridors 13-48 columns wide (a third to all of every
row is "tail"), scorings with cheap gap opens / expensive mismatches, short periodic sequences that tie a lot.
Prints the number of cells and mismatches; exits non-zero on any mismatch."""
import argparse
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCORINGS = [(1.0, -20.0, -1.0, -1.0, -1.0, 0.15), (2.0, -8.0, -1.0, -2.0, -0.5, 0.3), (1.0, -3.0, -0.5, -0.5, -0.25, 0.05),
            (3.0, -30.0, -2.0, -1.0, -1.0, 0.0), (1.0, -1.0, -1.0, -1.0, -1.0, 0.15), (2.0, -5.0, -5.0, -5.0, -1.0, 0.15)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=float, default=1.2e8)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    from ngmlr_b200 import corridor, synth
    from oracle_lib import Oracle, Reference
    rng = np.random.default_rng(args.seed)
    orc = Oracle()
    cells = 0
    bad = 0
    n = 0
    per_scoring = {}
    n_scalar_probe = n_scalar_differs = 0
    while cells < args.cells:
        sc = SCORINGS[n % len(SCORINGS)]
        ref_eng = Reference(scoring=sc)
        for _ in range(200):
            H = int(rng.integers(20, 400))
            W = int(rng.integers(13, 49))
            kind = int(rng.integers(0, 3))
            if kind == 0:      # periodic: many equal-score paths
                unit = synth.random_genome(int(rng.integers(1, 5)), int(rng.integers(1 << 30)))
                r = np.tile(unit, H + 80)[:H + 60]
            else:
                r = synth.random_genome(H + 60, int(rng.integers(1 << 30)))
            q, _ = synth.mutate(r[10:10 + H], rng, err=float(rng.choice([0.05, 0.2, 0.4])))
            q = q[:H] if q.size >= H else np.concatenate([q, r[:H - q.size]])
            shift = int(rng.integers(-5, 15))
            o, l = corridor.corridor_linear(len(q), W)
            o = (o + 10 + shift).astype(np.int32)
            a = ref_eng.fill(r.tobytes(), q.tobytes(), o, l, 0)
            b = orc.fill(r.tobytes(), q.tobytes(), o, l, sc, 2)
            if n_scalar_probe < 3000:   # how often does this input separate the SSE semantics from the scalar rule at all?
                n_scalar_probe += 1
                s1 = orc.fill(r.tobytes(), q.tobytes(), o, l, sc, 1)
                n_scalar_differs += not (np.array_equal(a[0], s1[0]) and a[1:] == s1[1:])
            same = (np.array_equal(a[0], b[0]) and np.float32(a[1]).view(np.uint32) == np.float32(b[1]).view(np.uint32)
                    and a[2:] == b[2:])
            c = int(np.maximum(np.minimum(o + l, len(r)) - np.maximum(o, 0), 0).sum())
            cells += c
            per_scoring[sc] = per_scoring.get(sc, 0) + c
            if not same:
                bad += 1
                print("MISMATCH", sc, H, W, a[1:], b[1:], file=sys.stderr)
        ref_eng.close()
        n += 1
    print(f"{n_scalar_differs} of the first {n_scalar_probe} problems separate the as-coded SSE semantics from the pure "
          f"scalar rule (the fuzz is not vacuous)")
    print(f"cells {cells} problems {n * 200} mismatches {bad}; cells per scoring: "
          + ", ".join(f"{k}: {v}" for k, v in per_scoring.items()))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
