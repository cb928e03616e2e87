"""Multi-GPU plumbing of the hot path: one process per GPU, reads sharded by rank, the packed
reference distributed ONCE at start-up, no per-step collective.

`broadcast_reference` ships what every rank needs to run stage 0/2 and stage 4 -- the 4-bit encoded
genome (`binRef`), the k-mer index arrays of CompactPrefixTable and, for synthetic benchmarks, the
flat genome the reads are simulated from -- as ONE packed buffer in ONE NCCL broadcast (plus a
64-byte header so that the receivers can size it). torch.distributed is the plumbing; the payload
layout is the reference's own in-memory format (ngmlr_b200.refindex).
"""
import numpy as np


def shard(n_items, rank, world):
    """Read indices of `rank`: i % world == rank; every read is independent, the
    sub-reads and intervals of a read stay on its rank."""
    return range(rank, n_items, world)


def broadcast_reference(genome, enc_ref, kidx, src=0, device=None):
    """-> (genome, EncodedReference, KmerIndex) on every rank. On `src` the arguments are the built
    reference; elsewhere they are ignored (pass None)."""
    import torch
    import torch.distributed as dist
    from . import refindex
    rank = dist.get_rank()
    hdr = torch.zeros(8 + 2 * 64, dtype=torch.int64, device=device)
    if rank == src:
        nc = len(enc_ref.ref_start)
        assert nc <= 64
        parts = [np.ascontiguousarray(genome).view(np.uint8), np.ascontiguousarray(enc_ref.enc).view(np.uint8),
                 np.ascontiguousarray(kidx.tab).view(np.uint8), np.ascontiguousarray(kidx.rci).view(np.uint8),
                 np.ascontiguousarray(kidx.pos).view(np.uint8)]
        meta = [p.size for p in parts] + [enc_ref.concat_len, nc, (kidx.k << 8) | kidx.bin_shift]
        meta += list(enc_ref.ref_start) + [0] * (64 - nc) + list(enc_ref.ref_len) + [0] * (64 - nc)
        hdr.copy_(torch.tensor(meta, dtype=torch.int64))
    dist.broadcast(hdr, src=src)
    m = [int(x) for x in hdr.cpu()]
    sizes, concat_len, nc, kb = m[:5], m[5], m[6], m[7]
    total = sum(sizes)
    buf = torch.empty(total, dtype=torch.uint8, device=device)
    if rank == src:
        at = 0
        for p in parts:
            buf[at:at + p.size].copy_(torch.from_numpy(p))
            at += p.size
    dist.broadcast(buf, src=src)          # the one collective of the run
    if rank == src:
        return genome, enc_ref, kidx
    host = buf.cpu().numpy()
    del buf
    at = 0
    arrs = []
    for sz in sizes:
        arrs.append(host[at:at + sz])
        at += sz
    enc = refindex.EncodedReference(arrs[1], concat_len, m[8:8 + nc], m[8 + 64:8 + 64 + nc])
    idx = refindex.KmerIndex(kb >> 8, kb & 0xff, arrs[2].view(np.uint32), arrs[3].view(np.int8), arrs[4].view(np.uint32))
    return arrs[0], enc, idx
