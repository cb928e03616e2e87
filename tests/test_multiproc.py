"""world_size-2 CPU (gloo) test of the multi-rank host logic bench.py uses: the reference is
broadcast once from rank 0, reads are sharded by rank (rank-seeded pools, no overlap, no per-step
collective), timings are max-reduced and totals sum-reduced."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json, hashlib
    sys.path.insert(0, os.environ["REPO_ROOT"])
    import numpy as np, torch, torch.distributed as dist
    from ngmlr_b200 import synth, PackedBatch
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 200_000
    g = torch.zeros(n, dtype=torch.uint8)
    if rank == 0:
        g.copy_(torch.from_numpy(synth.random_genome(n, 1)))
    dist.broadcast(g, src=0)                      # the one start-up collective
    genome = g.numpy()
    pool = synth.pacbio_problems(6, seed=2 + rank, median=1500, genome=genome)   # read sharding
    batch = PackedBatch.from_problems(pool)
    sig = hashlib.sha256(b"".join(p.qry for p in pool)).hexdigest()
    t = torch.tensor([10.0 + rank, 5.0 - rank], dtype=torch.float64)   # fake per-rank timings
    tot = torch.tensor([float(batch.read_bases)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    sigs = [None] * world
    dist.all_gather_object(sigs, (sig, batch.read_bases, hashlib.sha256(genome.tobytes()).hexdigest()))
    if rank == 0:
        print(json.dumps({"tmax": t.tolist(), "tot": tot.item(), "sigs": sigs}))
    dist.destroy_process_group()
''')


def test_two_rank_sharding_and_reductions(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    (s0, b0, g0), (s1, b1, g1) = d["sigs"]
    assert g0 == g1, "both ranks must hold the broadcast reference"
    assert s0 != s1, "ranks must align different reads"
    assert d["tot"] == b0 + b1
    assert d["tmax"] == [11.0, 5.0]


def test_bench_reference_arm_runs_on_rank0_only(tmp_path):
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints the line, rank 1 exits 0."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29732",
                          os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "4", "--genome-mb", "1"],
                         capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference"
    assert lines[0]["value"] > 0 and lines[0]["cpu_baseline"]["kind"] in ("reference", "port")
    assert lines[0]["e2e"]["h2d_bytes_per_step"] == 0
