"""Synthetic reference + FASTQ reads for whole-pipeline ngmlr runs (plain reference binary vs the
same binary with its aligners swapped for the CUDA plugin)."""
import numpy as np

from ngmlr_b200 import synth


def write_dataset(dirpath, n_reads=24, seed=11):
    rng = np.random.default_rng(seed)
    contigs = [synth.random_genome(120_000, seed + 1), synth.random_genome(80_001, seed + 2)]
    ref = f"{dirpath}/ref.fa"
    with open(ref, "w") as f:
        for i, c in enumerate(contigs):
            f.write(f">chr{i + 1}\n")
            s = c.tobytes().decode()
            for k in range(0, len(s), 80):
                f.write(s[k:k + 80] + "\n")
    fq = f"{dirpath}/reads.fq"
    with open(fq, "w") as f:
        for i in range(n_reads):
            c = contigs[i % 2]
            L = int(rng.integers(1500, 7000))
            start = int(rng.integers(1000, c.size - L - 1000))
            window = c[start:start + L]
            kind = i % 6
            if kind == 3:      # 400-bp deletion in the read
                window = np.concatenate([window[:L // 2], window[L // 2 + 400:]])
            elif kind == 4:    # 300-bp insertion
                window = np.concatenate([window[:L // 2], synth.random_genome(300, seed + 100 + i), window[L // 2:]])
            elif kind == 5:    # inversion of a 600-bp segment
                a = L // 2
                window = np.concatenate([window[:a], synth.revcomp(window[a:a + 600]), window[a + 600:]])
            read, _ = synth.mutate(window, rng, err=0.12)
            if rng.integers(0, 2):
                read = synth.revcomp(read)
            s = read.tobytes().decode()
            f.write(f"@read{i}\n{s}\n+\n{'I' * len(s)}\n")
    return ref, fq


def sam_records(path):
    out = []
    for line in open(path):
        if line.startswith("@"):
            continue
        out.append(line.rstrip("\n"))
    return sorted(out)
