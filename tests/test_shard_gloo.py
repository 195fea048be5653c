"""CPU, world_size 2 over gloo: the N > 1 path's host logic -- block-range partition balanced by
compressed bytes, scatter of compressed ranges, per-rank decode of the rebased job table, gather
of decoded ranges.  The decode stand-in here is the oracle (no GPU in this container); on the
GPU box bench.py --gpus N runs the same partition with the CUDA decode."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def test_partition_properties():
    from zxc_b200.shard import partition_blocks, rank_slice
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 7, 100, 1000):
        comp = rng.integers(8, 70000, n)
        for world in (1, 2, 3, 8):
            parts = partition_blocks(comp, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            if n >= 50 * world:
                loads = [int(comp[a:b].sum()) for a, b in parts]
                assert max(loads) - min(loads) <= 2 * 70000
    lo, hi, dlo, dhi = rank_slice([100, 200, 300], 4096, 10000, 1, 3)
    assert (lo, hi, dlo, dhi) == (116, 616, 4096, 10000)


def _worker(rank, world, port, frame_path, data_path, q):
    try:
        import torch
        import torch.distributed as dist
        import zxc_ctypes as z
        from zxc_b200.shard import gather_output, rebase_jobs, scatter_ranges, rank_slice
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        orc = z.Oracle()
        frame = np.fromfile(frame_path, dtype=np.uint8)
        data = np.fromfile(data_path, dtype=np.uint8)
        # every rank knows the table (it is tiny); only rank 0 "holds" the frame bytes
        import ctypes as C
        bs, tot = C.c_uint32(), C.c_uint64()
        nb = orc.lib.zxo_seek_parse(frame.ctypes.data, frame.size, C.byref(bs), C.byref(tot), None, 0)
        comp = np.zeros(nb, dtype=np.uint32)
        assert orc.lib.zxo_seek_parse(frame.ctypes.data, frame.size, C.byref(bs), C.byref(tot), comp.ctypes.data, nb) == nb
        fr_t = torch.from_numpy(frame) if rank == 0 else None
        mine, b0, b1 = scatter_ranges(fr_t, comp, bs.value, tot.value, dist)
        lo, hi, dlo, dhi = rank_slice(comp, bs.value, tot.value, b0, b1)
        assert mine.numel() == hi - lo
        assert np.array_equal(mine.numpy(), frame[lo:hi])
        # decode this rank's blocks one by one from its slice (offsets rebased to the slice)
        out = np.zeros(dhi - dlo, dtype=np.uint8)
        src = mine.numpy()
        p = 0
        for i in range(b0, b1):
            cap = min(bs.value, tot.value - i * bs.value)
            o = (i - b0) * bs.value
            r = orc.lib.zxo_decode_block(src.ctypes.data + p, int(comp[i]), out.ctypes.data + o, cap, None, 0, None, 0)
            assert r == cap, (i, r)
            p += int(comp[i])
        whole = gather_output(torch.from_numpy(out), comp, bs.value, tot.value, dist)
        ok = True
        if rank == 0:
            ok = np.array_equal(whole.numpy(), data)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), b0, b1))
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e), None))


def test_scatter_decode_gather_world2(tmp_path):
    import torch.multiprocessing as mp
    import zxc_corpus as zc
    import zxc_ctypes as z
    if not z.have_ref():
        pytest.skip("needs oracle/_ref to produce the frame")
    ref = z.ZxcLib(z.REF_SO)
    data = zc.silesia_shaped(3 << 20, seed=17)[: (3 << 20) - 12345]
    frame = ref.compress(data, level=3, block_size=65536, seekable=1)
    fp, dp = str(tmp_path / "f.zxc"), str(tmp_path / "d.bin")
    frame.tofile(fp)
    data.tofile(dp)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fp, dp, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    res.sort()
    assert all(r[1] is True for r in res), res
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == 48


def _p2p_worker(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist
        from zxc_b200 import shard
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        shard.CHUNK = 1000  # several messages per range
        rng = np.random.default_rng(1)
        frame = torch.from_numpy(rng.integers(0, 256, 10000, dtype=np.uint8)) if rank == 0 else None
        ranges = [(16, 4211), (4211, 10000)]
        mine = shard.scatter_ranges_p2p(frame, ranges, dist, "cpu")
        ref = np.random.default_rng(1).integers(0, 256, 10000, dtype=np.uint8)
        lo, hi = ranges[rank]
        ok = np.array_equal(mine.numpy(), ref[lo:hi])
        # the inverse: every rank contributes its (transformed) range, rank 0 ends up with all of it
        out_rng = [(0, 3000), (3000, 7777)]
        dec = torch.full((out_rng[rank][1] - out_rng[rank][0],), rank + 1, dtype=torch.uint8)
        whole = torch.zeros(7777, dtype=torch.uint8) if rank == 0 else None
        if rank == 0:
            whole[0:3000].copy_(dec)  # the root decodes in place
        shard.gather_ranges_p2p(dec, whole, out_rng, dist)
        if rank == 0:
            ok = ok and bool((whole[:3000] == 1).all()) and bool((whole[3000:] == 2).all())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok)))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


def test_p2p_scatter_gather_world2():
    """the batched send/recv exchange bench.py's pipeline leg runs over NCCL, here over gloo"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
