#!/usr/bin/env python
"""Record the real SingleAlign / BatchScore / SingleScore stream of plain ngmlr (SURVEY.md section 7 step 0,
BASELINE.md section 3.2): simulate a configs[1]-shaped FASTQ (seeded), run oracle/_ref/ngmlr_rec -- the
unmodified reference with recording decorators around its own ConvexAlignFast / StrippedSW
(oracle/record_aligners.cpp) -- and leave the call log + a small JSON sidecar.

  python scripts/record_workload.py --reads 10000 --genome-mb 50 --threads 32 --out /tmp/ngmlr_calls

The log is large (about 80 KB per SingleAlign call) and is therefore generated where it is replayed
(scripts/replay_workload.py), not committed."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10000)
    ap.add_argument("--genome-mb", type=float, default=50.0)
    ap.add_argument("--contigs", type=int, default=5)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--median", type=int, default=8000)
    ap.add_argument("--out", default="/tmp/ngmlr_calls")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--sv", action="store_true", help="reads with one 1-50 kb insertion / deletion / inversion each "
                                                      "(configs[4] shape): split reads, retries, realignments")
    args = ap.parse_args()
    from ngmlr_b200 import synth
    os.makedirs(args.out, exist_ok=True)
    contig_len = int(args.genome_mb * 1e6) // args.contigs
    genome = synth.random_genome(contig_len * args.contigs, 1)
    ref = os.path.join(args.out, "ref.fa")
    with open(ref, "w") as f:
        for c in range(args.contigs):
            f.write(f">c{c}\n")
            s = genome[c * contig_len:(c + 1) * contig_len].tobytes().decode()
            f.write("\n".join(s[k:k + 80] for k in range(0, len(s), 80)) + "\n")
    reads, _ivs = synth.simulate_reads(args.reads, genome, contig_len, args.seed, median=args.median, sv=args.sv)
    fq = os.path.join(args.out, "reads.fq")
    bases = 0
    with open(fq, "w") as f:
        for i, r in enumerate(reads):
            s = r.decode()
            bases += len(s)
            f.write(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n")
    log = os.path.join(args.out, "calls.bin")
    exe = os.path.join(ROOT, "oracle", "_ref", "ngmlr_rec")
    t0 = time.time()
    r = subprocess.run([exe, "-r", ref, "-q", fq, "-o", os.path.join(args.out, "rec.sam"), "-t", str(args.threads),
                        "--no-progress"], capture_output=True, text=True, env=dict(os.environ, NGMLR_RECORD_FILE=log))
    wall = time.time() - t0
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-2000:])
        raise SystemExit(f"ngmlr_rec failed with {r.returncode}")
    done = [ln for ln in r.stderr.splitlines() if "Done" in ln]
    side = {"reads": args.reads, "read_bases": bases, "genome_mb": args.genome_mb, "threads": args.threads,
            "wall_s_incl_index_build": wall, "ngmlr_done_line": done[-1] if done else "", "log": log,
            "log_bytes": os.path.getsize(log),
            "recipe": "scripts/record_workload.py " + " ".join(sys.argv[1:])}
    json.dump(side, open(os.path.join(args.out, "calls.json"), "w"), indent=1)
    print(json.dumps(side))


if __name__ == "__main__":
    main()
