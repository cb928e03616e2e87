#!/usr/bin/env python
"""Differential fuzz of the SAM text (csrc/sam_text.cpp) against the UNMODIFIED reference writer (SAMWriter +
GenericReadWriter::WriteRead through oracle/_ref/libngmlr_full.so): seeded batches of made-up reads
(tests/sam_cases.make_reads: several alignments per read, skipped / all-skipped / empty / unmapped reads, positions
around 2^31 and 2^32, read groups, --bam-fix, reads without qualities) formatted by both, compared byte for byte.

  python scripts/fuzz_sam_text.py --seeds 200 --reads 120

Needs oracle/_ref (this container). The reference keeps singletons: the whole fuzz runs in this one process."""
import argparse
import ctypes as C
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--reads", type=int, default=120)
    ap.add_argument("--first-seed", type=int, default=1000)
    args = ap.parse_args()
    import oracle_lib
    import sam_cases
    from ngmlr_b200 import samtext as st
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libngmlr_full.so"))
    d = tempfile.mkdtemp(prefix="samfuzz_")
    fa = os.path.join(d, "ref.fa")
    open(fa, "wb").write(sam_cases.fasta_bytes())
    lib.ref_cs_init(fa.encode())
    ref = oracle_lib.SamReference(lib)
    total = records = 0
    for k in range(args.seeds):
        seed = args.first_seed + k
        reads = sam_cases.make_reads(seed, n_reads=args.reads, with_big_cigar=(k % 25 == 0))
        opts = dict(write_unmapped=bool(k % 3), bam_cigar_fix=bool(k % 2), rg_id=(b"rg%d" % k if k % 4 == 0 else None))
        want = ref.records(reads, **opts)
        got = st.sam_format(reads, ref.names, threads=1 + k % 5, **opts)
        if got != want:
            w, g = want.split(b"\n"), got.split(b"\n")
            for i, (a, b) in enumerate(zip(w, g)):
                if a != b:
                    print(f"seed {seed}: line {i} differs\n  reference: {a[:300]!r}\n  library:   {b[:300]!r}")
                    break
            sys.exit(1)
        total += len(want)
        records += want.count(b"\n")
    print(f"{args.seeds} batches, {records} records, {total} bytes of SAM text: 0 mismatches")


if __name__ == "__main__":
    main()
