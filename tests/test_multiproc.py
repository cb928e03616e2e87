"""world_size-2 CPU (gloo) test of the multi-rank host logic bench.py uses: the reference is
broadcast once from rank 0 as one packed buffer (ngmlr_b200.parallel.broadcast_reference: the 4-bit
encoding with its contig table; every rank decodes the identical flat genome from it), reads are sharded by
rank (rank-seeded pools, no overlap, no per-step collective), timings are max-reduced and totals
sum-reduced."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json, hashlib
    sys.path.insert(0, os.environ["REPO_ROOT"])
    import numpy as np, torch, torch.distributed as dist
    from ngmlr_b200 import parallel, refindex, synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, n_contigs = 200_000, 2
    genome = enc = None
    if rank == 0:
        genome = synth.random_genome(n, 1)
        enc = refindex.encode_reference([genome[:n // 2], genome[n // 2:]])
    # the one start-up collective (+ its header): the packed reference; the flat genome is decoded from it
    enc = parallel.broadcast_reference(enc, src=0)
    decoded = np.concatenate(refindex.decode_contigs(enc))
    assert genome is None or np.array_equal(genome, decoded)
    genome = decoded
    reads, ivs = synth.simulate_reads(6, genome, n // 2, 2 + rank, median=1500)   # read sharding: own reads per rank
    tasks = synth.interval_tasks(ivs, reads, lambda pos: enc.ref_start[pos // (n // 2)] + pos % (n // 2))
    assert all(t.read_index == k for k, t in enumerate(tasks))
    bases = sum(len(r) for r in reads)
    sig = hashlib.sha256(b"".join(reads)).hexdigest()
    t = torch.tensor([10.0 + rank, 5.0 - rank], dtype=torch.float64)   # fake per-rank timings
    tot = torch.tensor([float(bases)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ref_sig = hashlib.sha256(genome.tobytes() + enc.enc.tobytes()
                             + repr((enc.concat_len, list(enc.ref_start), list(enc.ref_len))).encode()).hexdigest()
    sigs = [None] * world
    dist.all_gather_object(sigs, (sig, bases, ref_sig))
    assert list(parallel.shard(10, rank, world)) == list(range(rank, 10, world))
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


SAM_WORKER = textwrap.dedent('''
    import os, sys, json, hashlib
    sys.path.insert(0, os.environ["REPO_ROOT"])
    sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
    import torch.distributed as dist
    import sam_cases
    from ngmlr_b200 import parallel, samtext as st
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    names = [b"c0", b"c1", b"c2", b"c3"]
    reads = sam_cases.make_reads(21, n_reads=90)
    mine = [reads[i] for i in parallel.shard(len(reads), rank, world)]     # reads i % world == rank
    body = st.sam_format(mine, names, threads=2)
    bodies = [None] * world
    dist.all_gather_object(bodies, body)                                   # the optional final gather (SURVEY 8(e))
    if rank == 0:
        header = st.sam_header(names, [1000, 2000, 3000, 4000])
        whole = header + b"".join(bodies)                                  # one header, per-rank bodies behind it
        single = st.sam_format(reads, names, threads=1)
        print(json.dumps({"lines": whole.count(b"\\n"), "header_lines": header.count(b"\\n"),
                          "same_records": sorted(b"".join(bodies).split(b"\\n")) == sorted(single.split(b"\\n")),
                          "rank_major": b"".join(bodies) == b"".join(st.sam_format([reads[i] for i in parallel.shard(
                              len(reads), r, world)], names) for r in range(world)),
                          "sha": hashlib.sha256(whole).hexdigest()}))
    dist.destroy_process_group()
''')


def test_two_rank_sam_output_is_one_header_plus_the_rank_bodies(tmp_path):
    """reads are sharded by index % nGPU, every rank formats the SAM body of its own reads
    (ngmlr_b200_sam_format), the output is one header followed by the per-rank bodies: the same records a single
    process writes for all reads."""
    script = tmp_path / "sam_worker.py"
    script.write_text(SAM_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29733", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["same_records"] and d["rank_major"] and d["header_lines"] == 6 and d["lines"] > 60
