#!/usr/bin/env python
"""bench.py -- aligned Gbp/s of the convex-gap banded alignment hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic PacBio-shaped reads: stage 0/2
(k-mer candidate search of every 256-bp sub-read, device-side window decode and StrippedSW scoring
of every candidate) followed by stage 4 (convex fill -> traceback -> binary CIGAR -> CIGAR/MD text of
the read's interval alignment): BASELINE.json configs[1] (synthetic 50 Mb
reference, ~8 kb reads, 15 % errors ins:del:sub 9:4:2, anchored corridor) sharded by read across
ranks (weak scaling: every rank aligns its own `--reads` reads per step; no per-step collective;
one NCCL broadcast of the reference at start-up).

  value      whole-job Gbp/s with the batch already resident in HBM (K x kernels only, CUDA events);
             the batch is dealt read by read to `--contexts` aligner contexts that run concurrently
  e2e        same metric through the public calls (B200Aligner -> C ABI) from HOST buffers:
             pack + H2D + kernels + D2H + CIGAR/MD text every step, same contexts
  roofline   fill kernel: algorithmic bytes per launch / mean launch time vs measured HBM peak
  cpu_baseline  the reference's own CPU ConvexAlignFast (oracle/_ref) or the oracle port, timed on
             this box's host cores on a bounded sample of the same workload

`--impl reference` times the CPU implementation instead (all host threads) and prints the same
line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "configs[1]: synthetic 50 Mb i.i.d. reference, PacBio-shaped reads (log-normal, median 8 kb, " \
           "15% errors ins:del:sub 9:4:2, strand 50/50), one interval alignment per read, corridor from 256-bp anchors"


def make_pool(genome, n_reads, seed):
    from ngmlr_b200 import synth
    return synth.pacbio_problems(n_reads, seed=seed, median=8000, genome=genome)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class CpuStage02:
    """Stage 0/2 on the CPU with the reference's own code (oracle/_ref/libngmlr_full.so: CS vote,
    DecodeRefSequence, StrippedSW), or with the oracle port when that library is absent."""

    def __init__(self, genome, n_contigs=5):
        import ctypes as C
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        self.C = C
        step_c = genome.size // n_contigs
        contigs = [genome[i * step_c:(i + 1) * step_c] for i in range(n_contigs)]
        self.kind = "port"
        if oracle_lib.CsReference.available():
            try:
                fasta = f"/tmp/ngmlr_b200_bench_{os.getpid()}.fa"
                with open(fasta, "w") as f:
                    for i, c in enumerate(contigs):
                        f.write(f">c{i}\n{c.tobytes().decode()}\n")
                lib = C.CDLL(oracle_lib.CsReference.PATH)
                lib.ref_cs_init(fasta.encode())
                lib.ref_cs_probe_create.restype = C.c_void_p
                lib.ref_full_ssw_create.restype = C.c_void_p
                lib.ref_full_ssw_score.restype = C.c_float
                lib.ref_full_ssw_score.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
                self.fast = hasattr(lib, "ref_stage02_read")   # the whole per-read loop in C (no GIL between calls)
                if self.fast:
                    lib.ref_stage02_read.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
                self.lib = lib
                self.kind = "reference"
                os.unlink(fasta)
            except OSError:
                self.kind = "port"
        if self.kind == "port":
            self.orc = oracle_lib.CsOracle([c.tobytes() for c in contigs])
            self.ssw = oracle_lib.Oracle()

    def worker_state(self):
        if self.kind == "reference":
            return (self.C.c_void_p(self.lib.ref_cs_probe_create()), self.C.c_void_p(self.lib.ref_full_ssw_create()))
        return None

    _CPL = bytes.maketrans(b"ACGT", b"TGCA")

    def read(self, st, qry):
        """All sub-reads of one read: vote, then score every candidate. Returns #candidates."""
        C = self.C
        if self.kind == "reference" and self.fast:
            return self.lib.ref_stage02_read(st[0], st[1], bytes(qry), len(qry), 256, 20, 308)
        n_c = 0
        for k in range(len(qry) // 256):
            sub = qry[k * 256:(k + 1) * 256]
            if self.kind == "reference":
                sc = (C.c_float * 512)()
                lo = (C.c_ulonglong * 512)()
                rv = (C.c_int * 512)()
                mh = C.c_float()
                n = self.lib.ref_cs_search_p(st[0], sub, len(sub), 16, sc, lo, rv, 512, C.byref(mh))
                buf = C.create_string_buffer(312)
                for j in range(max(0, min(n, 512))):
                    if not self.lib.ref_cs_decode(C.c_ulonglong(lo[j] - 20), C.c_ulonglong(308), buf):
                        buf.value = b"N" * 308
                    q = sub.translate(self._CPL)[::-1] if rv[j] else sub
                    self.lib.ref_full_ssw_score(st[1], buf.value, q)
                    n_c += 1
            else:
                cands, _ = self.orc.search(sub)
                for (_s, loc, rev) in cands:
                    w = self.orc.decode((loc - 20) % (1 << 64), 308) or b"N" * 308
                    self.ssw.ssw_score(w, sub.translate(self._CPL)[::-1] if rev else sub)
                    n_c += 1
        return n_c


def effective_cpus():
    """Host CPUs this process can really use: logical CPUs, limited by the affinity mask and by the
    container's CFS quota (cgroup v2 cpu.max / v1 cfs_quota_us) -- the GPU boxes expose 128 logical
    CPUs under a 16-CPU quota, and 128 busy threads there run slower than 32."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(per)
    except (OSError, ValueError):
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except (OSError, ValueError):
            pass
    if quota:
        n = max(1, min(n, int(quota * 2)))   # 2 threads per quota CPU measured best for the CPU arm
    return n


def tune_malloc_for_threads():
    """The reference allocates its multi-megabyte direction matrix per SingleAlign call; with glibc's
    defaults every such allocation is an mmap/munmap pair and the threads of one process serialise on
    the kernel's address-space lock (measured here: 4 threads = 1.0x one thread). Keeping large blocks
    in per-thread arenas lets the CPU arm scale with the host threads -- the faster, fairer baseline."""
    import ctypes as C
    try:
        libc = C.CDLL("libc.so.6")
        libc.mallopt(C.c_int(-3), C.c_int(1 << 30))    # M_MMAP_THRESHOLD
        libc.mallopt(C.c_int(-1), C.c_int(1 << 30))    # M_TRIM_THRESHOLD
        libc.mallopt(C.c_int(-8), C.c_int(256))        # M_ARENA_MAX
    except OSError:
        pass


def cpu_reference_run(problems, threads, impl, stage02=None):
    """Run problems through the CPU implementation on `threads` host threads (ctypes releases the
    GIL). impl: 'reference' = oracle/_ref (unmodified ConvexAlignFast), 'port' = oracle C port.
    stage02: CpuStage02 or None (stage 4 only)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    tune_malloc_for_threads()
    work = list(range(len(problems)))
    lock = threading.Lock()
    engines = [(oracle_lib.Reference() if impl == "reference" else oracle_lib.Oracle(),
                stage02.worker_state() if stage02 else None) for _ in range(threads)]

    def worker(t):
        eng, st = engines[t]
        while True:
            with lock:
                if not work:
                    break
                i = work.pop()
            p = problems[i]
            if stage02:
                stage02.read(st, p.qry)
            r = eng.single_align(p.ref, p.qry, p.offsets, p.lengths)
            assert r["ret"] == len(p.qry)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    if impl == "reference":
        for eng, _ in engines:
            eng.close()
    return dt


def cpu_impl_kind():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    if oracle_lib.Reference.available():
        try:
            oracle_lib.Reference().close()
            return "reference"
        except OSError:
            pass
    oracle_lib.Oracle()
    return "port"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=8192, help="reads (alignment problems) per step per GPU")
    ap.add_argument("--genome-mb", type=float, default=50.0)
    ap.add_argument("--dp-only", action="store_true", help="time stage 4 (convex alignment) alone")
    ap.add_argument("--contexts", type=int, default=4, help="aligner contexts (host threads/streams) per GPU")
    ap.add_argument("--fill-ctas", type=int, default=4,
                    help="fill CTAs per SM per launch in the concurrent phases (0 = full occupancy; the solo "
                         "roofline phase always runs at full occupancy)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = effective_cpus()

    # host worker threads of the library (packing / CIGAR text): share the logical CPUs among ranks x
    # contexts. Deliberately not limited to the container's CPU quota: the host phases are short bursts
    # and measured faster with 32 threads per context than with 8 (0.90 vs 0.84 Gbp/s end to end under a
    # 16-CPU quota), whereas the CPU reference arm, which is busy all the time, is fastest at 2 x quota.
    os.environ.setdefault("NGMLR_B200_HOST_THREADS",
                          str(max(4, min(32, (os.cpu_count() or 1) // max(1, world * max(1, args.contexts))))))

    from ngmlr_b200 import synth

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        kind = cpu_impl_kind()
        genome = synth.random_genome(int(args.genome_mb * 1e6), 1)
        threads = cores
        # bounded sample: ~8 reads per thread per step (dynamic scheduling evens out the read lengths) keeps a
        # K+W run within minutes
        n = args.cpu_sample or max(threads, min(args.reads, 8 * threads))
        pool = make_pool(genome, n, seed=2)
        bases = sum(len(p.qry) for p in pool)
        cells = sum(p.cells for p in pool)
        st02 = None if args.dp_only else CpuStage02(genome)
        for _ in range(max(1, min(args.warmup, 1))):   # grows the per-thread malloc arenas, warms the caches
            cpu_reference_run(pool, threads, kind, st02)
        times = [cpu_reference_run(pool, threads, kind, st02) for _ in range(args.steps)]
        t = float(np.sum(times))
        val = bases * args.steps / t / 1e9
        line = {"metric": "aligned_gbp_per_s", "value": val, "unit": "Gbp/s", "impl": "reference",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": val / 5.56e-4, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "reads_per_step": n, "read_bases_per_step": bases,
                           "dp_cells_per_step": cells},
                "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": threads, "kind": kind,
                                 "sample": f"{n} reads ({bases} bases, {cells} DP cells) per step, {threads} threads "
                                           f"(= usable CPUs: {os.cpu_count()} logical, container quota applied), "
                                           + ("ConvexAlignFast::SingleAlign only" if args.dp_only else
                                              f"CS vote + DecodeRefSequence + StrippedSW ({st02.kind}) then ConvexAlignFast::SingleAlign"),
                                 "mcells_per_s_per_core": cells * args.steps / t / threads / 1e6},
                "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # reference: genome + 4-bit encoding + k-mer index built on rank 0, then ONE NCCL broadcast of
    # the packed reference at start-up; no collective per step
    from ngmlr_b200 import refindex
    n_genome = int(args.genome_mb * 1e6)
    n_contigs = 5
    t_ref0 = time.perf_counter()
    if rank == 0:
        genome = synth.random_genome(n_genome, 1)
        step_c = n_genome // n_contigs
        enc_ref = refindex.encode_reference([genome[i * step_c:(i + 1) * step_c] for i in range(n_contigs)])
        kidx = refindex.build_index(enc_ref)
        meta = [enc_ref.enc.size, enc_ref.concat_len, kidx.tab.size, kidx.pos.size] + enc_ref.ref_start + enc_ref.ref_len
    if world > 1:
        mt = torch.zeros(4 + 2 * n_contigs, dtype=torch.int64, device=dev)
        if rank == 0:
            mt.copy_(torch.tensor(meta, dtype=torch.int64))
        dist.broadcast(mt, src=0)
        m = [int(x) for x in mt.cpu()]
        bufs = {"genome": (n_genome, torch.uint8), "enc": (m[0], torch.uint8), "tab": (m[2], torch.int32),
                "rci": (m[2], torch.int8), "pos": (m[3], torch.int32)}
        got = {}
        for name, (sz, dt) in bufs.items():
            t = torch.empty(sz, dtype=dt, device=dev)
            if rank == 0:
                src = {"genome": genome, "enc": enc_ref.enc, "tab": kidx.tab.view(np.int32),
                       "rci": kidx.rci, "pos": kidx.pos.view(np.int32)}[name]
                t.copy_(torch.from_numpy(np.ascontiguousarray(src)))
            dist.broadcast(t, src=0)
            got[name] = t.cpu().numpy()
            del t
        if rank != 0:
            genome = got["genome"]
            enc_ref = refindex.EncodedReference(got["enc"], m[1], m[4:4 + n_contigs], m[4 + n_contigs:4 + 2 * n_contigs])
            kidx = refindex.KmerIndex(13, 4, got["tab"].view(np.uint32), got["rci"], got["pos"].view(np.uint32))
    t_ref = time.perf_counter() - t_ref0

    from ngmlr_b200 import B200Aligner, PackedBatch, PackedReads, split_read
    pool = make_pool(genome, args.reads, seed=2 + rank)   # reads sharded by rank: own reads per rank
    batch = PackedBatch.from_problems(pool)
    bases = batch.read_bases
    subreads = PackedReads([s for p in pool for s in split_read(p.qry)])   # ReadProvider::splitRead
    # S independent aligner contexts (own stream, own device arenas), driven by S host threads --
    # the reference's model of one aligner object per worker thread. The step's batch is dealt to the
    # contexts read by read (context j aligns reads j, j+S, ...); their work overlaps on the GPU, which
    # hides the tail of each fill launch and the traceback behind another context's fill, and (end to
    # end) packing/H2D/D2H/text of one slice behind the kernels of the others.
    S = max(1, args.contexts)
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    als = [B200Aligner(local_rank, stream=st_.cuda_stream) for st_ in streams]
    al = als[0]
    for a_ in als:
        a_.set_index(kidx)
        a_.set_reference(enc_ref)
    slices = [pool[j::S] for j in range(S)]
    sl_batch = [PackedBatch.from_problems(sl) for sl in slices]
    sl_reads = [PackedReads([s for p in sl for s in split_read(p.qry)]) for sl in slices]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run_threads(fn):
        ts = [threading.Thread(target=fn, args=(j,)) for j in range(S)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    # ---- (a) one context alone on the whole batch: per-kernel durations for the roofline (the fill
    # kernel timed in isolation, inputs resident) ----
    al.upload(batch)
    al.cs_upload(subreads)
    for _ in range(args.warmup):
        if not args.dp_only:
            al.cs_run()
        al.run()
    barrier()
    fill_ms, tb_ms, cs_ms = [], [], []
    n_cand = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for _ in range(args.steps):
        n_cand, ms_ = (0, 0.0) if args.dp_only else al.cs_run()
        cs_ms.append(ms_)
        al.run()
        st = al.stats()
        fill_ms.append(st["fill_ms"])
        tb_ms.append(st["traceback_ms"])
    e1.record(streams[0])
    torch.cuda.synchronize(dev)
    solo_ms = e0.elapsed_time(e1)
    res = al.fetch()
    st = al.stats()
    assert all(res.ret(i) == len(p.qry) for i, p in enumerate(pool)), "bench: invalid alignment in the timed batch"
    cells = st["cells"]
    del res

    # ---- (b) device-resident timed region: every context keeps its slice in HBM, exactly K steps ----
    # Smaller persistent fill grids per launch so that the launches of the S contexts (and their
    # memory-bound candidate-search / traceback kernels) share the SMs instead of queueing.
    for a_ in als:
        a_.set_fill_ctas_per_sm(args.fill_ctas if S > 1 else 0)

    def resident_warm(j):
        als[j].upload(sl_batch[j])
        als[j].cs_upload(sl_reads[j])
        for _ in range(args.warmup):
            if not args.dp_only:
                als[j].cs_run()
            als[j].run()

    run_threads(resident_warm)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    cur = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for st_ in streams:
        st_.wait_event(e0)
    ends = [torch.cuda.Event() for _ in range(S)]

    def dev_worker(j):
        for _ in range(args.steps):
            if not args.dp_only:
                als[j].cs_run()
            als[j].run()
        ends[j].record(streams[j])

    run_threads(dev_worker)
    for j in range(S):
        cur.wait_event(ends[j])
    e1.record(cur)
    barrier()
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    for j in range(S):
        r_ = als[j].fetch()
        assert all(r_.ret(i) == len(p.qry) for i, p in enumerate(slices[j])), "bench: invalid alignment in a slice"
        del r_

    # ---- (c) end to end from host buffers through the public calls: per step every context takes its
    # slice from host memory (pack + H2D), runs all kernels, and brings candidates, scores and
    # alignments (binary CIGAR -> CIGAR/MD text) back to the host. Uploads are issued before the
    # kernels and fetches after them so that a context's host work overlaps other contexts' kernels. ----
    io = [dict() for _ in range(S)]

    def e2e_steps(j, k):
        a_ = als[j]
        for _ in range(k):
            if not args.dp_only:
                a_.cs_upload(sl_reads[j])            # host buffers -> device
            a_.upload(sl_batch[j])
            if not args.dp_only:
                m_, _ms = a_.cs_run()
            a_.run()
            if not args.dp_only:
                cstart, _sc, _lo, _rv, sw_, _mx = a_.cs_fetch()   # candidates + scores -> host
                assert m_ == cstart[-1] and sw_.size == m_
                io[j]["h2d"] = sl_reads[j].bases + 12 * sl_reads[j].n
                io[j]["d2h"] = 17 * int(m_) + 12 * sl_reads[j].n
            out = a_.fetch()
            assert len(out) == sl_batch[j].n and out.ret(0) == len(slices[j][0].qry)
            del out

    run_threads(lambda j: e2e_steps(j, 1))
    barrier()
    t0 = time.perf_counter()
    run_threads(lambda j: e2e_steps(j, args.steps))
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    st_e2e = [a_.stats() for a_ in als]
    e2e_h2d = sum(x["h2d_bytes"] for x in st_e2e) + sum(x.get("h2d", 0) for x in io)
    e2e_d2h = sum(x["d2h_bytes"] for x in st_e2e) + sum(x.get("d2h", 0) for x in io)

    t_dev = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(bases), float(cells)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = (float(x) for x in t_dev.cpu())
    tot_bases, tot_cells = (float(x) for x in tot.cpu())

    if rank == 0:
        value = tot_bases * args.steps / (dev_ms * 1e-3) / 1e9
        e2e_val = tot_bases * args.steps / (e2e_ms * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst)" if peaks else "fallback 6650 GB/s"
        # algorithmic bytes of ONE fill launch: 0.25 B per DP cell (2-bit
        # direction) + sequences read once + corridor rows (8 B/row) read once
        rows = int(batch.row_start[-1])
        seq_b = int(batch.ref_lens.sum() + batch.qry_lens.sum())
        algo_bytes = cells * 0.25 + seq_b + rows * 8
        fill_s = float(np.mean(fill_ms)) * 1e-3
        achieved = algo_bytes / fill_s / 1e9
        issue_bound_cells = 148 * 4 * 32 * float(clocks.get("sm_mhz") or 1965.0) * 1e6 / (27.0 * 2.0)
        traffic = None
        try:  # DRAM bytes of one fill launch from the committed ncu capture (same reads/step only)
            tr = json.load(open(os.path.join(ROOT, "profiles", "fill_traffic.json")))
            if tr.get("reads_per_step") == args.reads:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "aligned_gbp_per_s", "value": value, "unit": "Gbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": value / 5.56e-4,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reads_per_step_per_gpu": args.reads,
                       "read_bases_per_step": tot_bases, "dp_cells_per_step": tot_cells,
                       "parallelism": f"read-sharded x{world}, no per-step collective; {S} aligner contexts/GPU, "
                                      f"fill grid {args.fill_ctas or 'full'} CTAs/SM per launch",
                       "l2": "inputs+direction arena per step exceed L2 (direction writes alone "
                             f"{st['dir_bytes'] / 1e6:.0f} MB/step/GPU)",
                       "vs_baseline_note": "README.md:25 end-to-end 5.56e-4 Gbp/s on 10 Opteron cores (whole "
                                           "pipeline); this path is the 91 % stage"},
            "reads_per_s": args.reads * world * args.steps / (dev_ms * 1e-3),
            "gcells_per_s": tot_cells * args.steps / (dev_ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "convex_fill_kernel", "launch_ms": fill_s * 1e3,
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "gcells_per_s_kernel": cells / fill_s / 1e9,
                         "alu_pipe_bound_gcells_per_s": issue_bound_cells / 1e9,
                         "frac_of_alu_pipe_bound": cells / fill_s / issue_bound_cells,
                         "note": "ALU-pipe-bound kernel (~27 ALU-pipe SASS instr per 32-cell step, 2 cycles each "
                                 "per SMSP, DESIGN.md 4.1); HBM frac is structurally ~0.02"},
            "kernel_ms_per_step": {"fill": float(np.mean(fill_ms)), "traceback": float(np.mean(tb_ms)),
                                   "stage02_cs_vote_decode_score": float(np.mean(cs_ms))},
            "stage02": {"subreads_per_step_per_gpu": subreads.n, "candidates_per_step_per_gpu": int(n_cand),
                        "sw_cell_updates_per_step_per_gpu": int(n_cand) * 257 * 307,
                        "reference_setup_s": t_ref,
                        "note": "k-mer vote of every 256-bp sub-read + device decode + StrippedSW score of every candidate"},
            "e2e": {"value": e2e_val, "unit": "Gbp/s", "h2d_bytes_per_step": e2e_h2d,
                    "d2h_bytes_per_step": e2e_d2h, "ms_per_step": e2e_ms / args.steps,
                    "host_ms_per_slice": {k: float(np.mean([x[k] for x in st_e2e]))
                                          for k in ("host_pack_ms", "host_h2d_ms", "host_run_ms",
                                                    "host_d2h_ms", "host_text_ms")},
                    "host_threads_per_context": st_e2e[0]["host_threads"],
                    "note": f"each of the {S} contexts takes every {S}-th read of the batch per step"},
            "solo": {"gbp_per_s": bases * args.steps / (solo_ms * 1e-3) / 1e9, "ms_per_step": solo_ms / args.steps,
                     "note": "one context alone on the whole batch, same K steps (kernels not overlapped)"},
            # per context and step: cs count, sizes, 4 x (scan init + scan), vote, count widening, compaction,
            # window decode + score, fill, traceback = 16 kernels (2 with --dp-only); profiles/launches_r01_fullpath.csv
            "gpu_launches": (2 if args.dp_only else 16) * S * args.steps,
            "clocks": clocks,
        }
        # CPU baseline on this box's host cores, bounded sample of the same workload
        try:
            kind = cpu_impl_kind()
            threads = cores
            n = args.cpu_sample or max(threads, min(len(pool), 8 * threads))
            sample = pool[:n]
            st02 = None if args.dp_only else CpuStage02(genome)
            cpu_reference_run(sample, threads, kind, st02)   # warm-up: per-thread malloc arenas, caches
            t = cpu_reference_run(sample, threads, kind, st02)
            sb = sum(len(p.qry) for p in sample)
            sc = sum(p.cells for p in sample)
            line["cpu_baseline"] = {"value": sb / t / 1e9, "unit": "Gbp/s", "cores": threads, "kind": kind,
                                    "sample": f"first {n} reads of the batch ({sb} bases, {sc} DP cells), "
                                              f"{threads} threads, {t:.1f} s, stages: "
                                              + ("4 only" if st02 is None else f"0/2 ({st02.kind}) + 4"),
                                    "mcells_per_s_per_core": sc / t / threads / 1e6}
        except Exception as ex:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "Gbp/s", "cores": 0, "kind": "unavailable",
                                    "sample": repr(ex)}
        print(json.dumps(line))
    for a_ in als:
        a_.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
