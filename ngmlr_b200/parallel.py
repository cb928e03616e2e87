"""Multi-GPU plumbing of the hot path: one process per GPU, reads sharded by rank, the packed
reference distributed ONCE at start-up, no per-step collective.

`broadcast_reference` ships the 4-bit encoded, spacer-padded genome (`binRef`, 0.5 byte per base) with its
contig table as ONE buffer in ONE NCCL broadcast (plus a 64-byte-aligned header so that the receivers
can size it). Everything else a rank needs is derived locally from it: the k-mer index is built on the
rank's own GPU (B200Aligner.build_index: CompactPrefixTable::CreateTable as scans + one radix sort, a
fraction of a second for 50 Mb), and synthetic benchmarks decode the flat genome they simulate reads from
(refindex.decode_contigs). torch.distributed is the plumbing; the payload layout is the reference's own
in-memory format (ngmlr_b200.refindex).
"""
import numpy as np


def shard(n_items, rank, world):
    """Read indices of `rank`: i % world == rank; every read is independent, the
    sub-reads and intervals of a read stay on its rank."""
    return range(rank, n_items, world)


def broadcast_reference(enc_ref, src=0, device=None):
    """-> EncodedReference on every rank. On `src` the argument is the encoded reference; elsewhere it is
    ignored (pass None)."""
    import torch
    import torch.distributed as dist
    from . import refindex
    rank = dist.get_rank()
    hdr = torch.zeros(8 + 2 * 64, dtype=torch.int64, device=device)
    if rank == src:
        nc = len(enc_ref.ref_start)
        assert nc <= 64
        payload = np.ascontiguousarray(enc_ref.enc).view(np.uint8)
        meta = [payload.size, enc_ref.concat_len, nc, 0, 0, 0, 0, 0]
        meta += list(enc_ref.ref_start) + [0] * (64 - nc) + list(enc_ref.ref_len) + [0] * (64 - nc)
        hdr.copy_(torch.tensor(meta, dtype=torch.int64))
    dist.broadcast(hdr, src=src)
    m = [int(x) for x in hdr.cpu()]
    size, concat_len, nc = m[0], m[1], m[2]
    buf = torch.empty(size, dtype=torch.uint8, device=device)
    if rank == src:
        buf.copy_(torch.from_numpy(payload))
    dist.broadcast(buf, src=src)          # the one collective of the run: the packed reference
    if rank == src:
        return enc_ref
    host = buf.cpu().numpy()
    del buf
    return refindex.EncodedReference(host, concat_len, m[8:8 + nc], m[8 + 64:8 + 64 + nc])
