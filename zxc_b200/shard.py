"""Block-range sharding of a seekable frame across ranks.

Blocks are independent by format (docs/FORMAT.md:651-662), so a frame shards into contiguous
block ranges with no data-path collective: rank r decodes blocks [b_r, b_{r+1}) from the byte range
[comp_offsets[b_r], comp_offsets[b_{r+1}]) and produces decoded bytes [b_r*bs, min(b_{r+1}*bs, total)).
Ranges are balanced by COMPRESSED bytes (the SEK prefix sums, src/lib/zxc_seekable.c:343-365),
which is what the decode time follows.

The collectives are only the two exchange steps either side of the decode: scatter of compressed
ranges from the rank that holds the frame, gather of decoded ranges.  They are written against
torch.distributed so the same code runs over NCCL/NVLink on the B200 box and over gloo in the
CPU tests (tests/test_shard_gloo.py).
"""
import numpy as np

FILE_HEADER = 16


def partition_blocks(comp_sizes, world):
    """[(b0, b1)] * world: contiguous block ranges with near-equal compressed bytes."""
    comp = np.asarray(comp_sizes, dtype=np.int64)
    n = comp.size
    if n == 0:
        return [(0, 0)] * world
    csum = np.concatenate([[0], np.cumsum(comp)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, cuts[-1]), n)
        cuts.append(b)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def rank_slice(comp_sizes, block_size, total, b0, b1):
    """Byte ranges of a block range: (src_lo, src_hi, dst_lo, dst_hi) in frame / output coordinates."""
    comp = np.asarray(comp_sizes, dtype=np.int64)
    offs = FILE_HEADER + np.concatenate([[0], np.cumsum(comp)])
    dst_lo = b0 * block_size
    dst_hi = min(b1 * block_size, total)
    return int(offs[b0]), int(offs[b1]), int(dst_lo), int(max(dst_hi, dst_lo))


def rebase_jobs(jobs, b0, b1, src_lo, dst_lo):
    """Slice a frame-wide job table (structured array src_off,dst_off,src_len,dst_cap) for one rank."""
    out = jobs[b0:b1].copy()
    out["src_off"] -= src_lo
    out["dst_off"] -= dst_lo
    return out


CHUNK = 1 << 30  # bytes per point-to-point message (keeps element counts far below 2^31)


def _pieces(t):
    return [t[o:o + CHUNK] for o in range(0, t.numel(), CHUNK)]


def scatter_ranges_p2p(frame, ranges, dist, device, src_rank=0):
    """Batched variant for NCCL: `ranges[r] = (lo, hi)` byte range of `frame` (held by src_rank) for rank r.
    All sends are posted together (one ncclGroup), so the root's NVLink egress is shared by all peers."""
    import torch
    rank = dist.get_rank()
    lo, hi = ranges[rank]
    mine = torch.empty(hi - lo, dtype=torch.uint8, device=device)
    ops = []
    if rank == src_rank:
        for r, (rlo, rhi) in enumerate(ranges):
            if r == src_rank:
                mine.copy_(frame[rlo:rhi])
            else:
                ops += [dist.P2POp(dist.isend, pc, r) for pc in _pieces(frame[rlo:rhi])]
    else:
        ops += [dist.P2POp(dist.irecv, pc, src_rank) for pc in _pieces(mine)]
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return mine


def gather_ranges_p2p(decoded, out, ranges, dist, dst_rank=0):
    """Inverse exchange into `out` (on dst_rank; its own range is expected to be in place already):
    every other rank sends `decoded` into out[lo:hi]."""
    rank = dist.get_rank()
    ops = []
    if rank == dst_rank:
        for r, (rlo, rhi) in enumerate(ranges):
            if r != dst_rank:
                ops += [dist.P2POp(dist.irecv, pc, r) for pc in _pieces(out[rlo:rhi])]
    else:
        ops += [dist.P2POp(dist.isend, pc, dst_rank) for pc in _pieces(decoded)]
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()


def scatter_ranges(frame, comp_sizes, block_size, total, dist, device="cpu", src_rank=0):
    """Rank `src_rank` holds `frame` (uint8 tensor); every rank returns (its compressed slice, b0, b1).

    One grouped send/recv per rank (NCCL: ncclSend/ncclRecv over NVLink)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = partition_blocks(comp_sizes, world)
    b0, b1 = parts[rank]
    lo, hi, _, _ = rank_slice(comp_sizes, block_size, total, b0, b1)
    mine = torch.empty(hi - lo, dtype=torch.uint8, device=device)
    if rank == src_rank:
        reqs = []
        for r, (rb0, rb1) in enumerate(parts):
            rlo, rhi, _, _ = rank_slice(comp_sizes, block_size, total, rb0, rb1)
            if r == src_rank:
                mine.copy_(frame[rlo:rhi])
            elif rhi > rlo:
                reqs.append(dist.isend(frame[rlo:rhi].contiguous(), dst=r))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(mine, src=src_rank)
    return mine, b0, b1


def gather_output(decoded, comp_sizes, block_size, total, dist, device="cpu", dst_rank=0):
    """Inverse exchange: rank `dst_rank` returns the whole decoded tensor, others None."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = partition_blocks(comp_sizes, world)
    if rank == dst_rank:
        out = torch.empty(total, dtype=torch.uint8, device=device)
        for r, (rb0, rb1) in enumerate(parts):
            _, _, dlo, dhi = rank_slice(comp_sizes, block_size, total, rb0, rb1)
            if r == dst_rank:
                out[dlo:dhi].copy_(decoded)
            elif dhi > dlo:
                buf = torch.empty(dhi - dlo, dtype=torch.uint8, device=device)
                dist.recv(buf, src=r)
                out[dlo:dhi].copy_(buf)
        return out
    if decoded.numel() > 0:
        dist.send(decoded.contiguous(), dst=dst_rank)
    return None
