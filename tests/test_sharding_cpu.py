"""Multi-GPU seam logic on CPU: world_size-2 gloo all-gather of shard seam summaries, folded into
the carry-in state of the later shard, checked against the oracle's state at the cut."""
import os
import socket
import struct

import numpy as np
import pytest
import torch.multiprocessing as mp

from etl_b200 import sharding, workloads as wl


def seam_words_of(stream: bytes) -> np.ndarray:
    """The shard's stream-state transformer, restated in python from apply.rs:600-626, 1927-2006."""
    n_rec = lsn = ord_ = 0
    has_begin = closed = 0
    pos, n = 0, len(stream)
    while pos < n:
        (flen,) = struct.unpack_from(">i", stream, pos + 1)
        kind = stream[pos + 30:pos + 31] if stream[pos + 5:pos + 6] == b"w" else b"k"
        if kind == b"B":
            has_begin, closed = 1, 0
            (lsn,) = struct.unpack_from(">Q", stream, pos + 31)
            ord_ = 1
        elif kind == b"C":
            closed = 1
            ord_ += 1
        elif kind in (b"R", b"I", b"U", b"D", b"T"):
            ord_ += 1
        n_rec += 1
        pos += 1 + flen
    return np.array([n_rec, 0, 0, lsn, ord_, has_begin | (closed << 1)], dtype=np.uint64)


def _worker(rank, world, port, shards, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    words = seam_words_of(shards[rank])
    allw = sharding.all_gather_seam(words)            # the one exchange step
    carry, base = sharding.carry_for_rank(allw, rank)
    out.put((rank, carry, base))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("cut_kind", ["mid_tx", "tx_boundary"])
def test_seam_all_gather_world2_gloo(oracle_mod, cut_kind):
    w = wl.make("c2", 0.004, n_segments=1)
    stream, _ = w.generate()
    raw = stream.tobytes()
    o = oracle_mod.Oracle()
    for tid, cols in w.table_schemas().items():
        o.put_table_schema(tid, cols)
    full = o.decode(raw)
    kinds = [chr(k) for k in full.rec_kind]
    mid = full.n_records // 2
    if cut_kind == "mid_tx":
        cut_rec = next(i for i in range(mid, full.n_records) if kinds[i] in "IUD" and kinds[i - 1] in "IUD")
    else:
        cut_rec = next(i for i in range(mid, full.n_records) if kinds[i] == "B")
    cut = int(full.rec_off[cut_rec])
    shards = [raw[:cut], raw[cut:]]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shards, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (c, b)) for r, c, b in (q.get(timeout=120) for _ in range(2)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o2 = oracle_mod.Oracle()
    for tid, cols in w.table_schemas().items():
        o2.put_table_schema(tid, cols)
    head = o2.decode(shards[0])
    assert res[0] == ((0, 0, 0), 0)
    carry1, base1 = res[1]
    assert base1 == head.n_records == cut_rec
    assert carry1[0] == head.carry_out[0]
    if carry1[0]:   # inside a transaction: final_lsn and next ordinal must match the reference state
        assert carry1 == head.carry_out
    # and decoding the tail with that carry reproduces the full decode
    tail = o2.decode(shards[1], carry1)
    assert tail.first_error[0] is None
    assert np.array_equal(tail.rec_tx_ordinal, full.rec_tx_ordinal[cut_rec:])
    assert np.array_equal(tail.rec_commit_lsn, full.rec_commit_lsn[cut_rec:])


def test_cut_points_are_record_starts():
    w = wl.make("c5", 0.0005, n_segments=2)
    stream, _ = w.generate()
    raw = stream.tobytes()
    from etl_b200 import pgoutput as pg
    anchors = np.array(pg.build_anchors(raw, 2048), dtype=np.uint64)
    cuts = sharding.cut_points(anchors, len(raw), 4)
    assert cuts[0] == 0 and cuts[-1] == len(raw) and cuts == sorted(cuts)
    starts = set()
    pos = 0
    while pos < len(raw):
        starts.add(pos)
        pos += 1 + int.from_bytes(raw[pos + 1:pos + 5], "big")
    starts.add(len(raw))
    assert all(c in starts for c in cuts)
