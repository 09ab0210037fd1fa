"""Worker of tests/test_gpu_sharded.py: one process per GPU (torchrun).  Every rank builds the SAME deterministic
stream, takes its byte range, runs etl_dec_decode_sharded (NCCL inside the library) and saves its planes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_dir, name, scale = sys.argv[1], sys.argv[2], float(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import torch
    import torch.distributed as dist
    from etl_b200 import decoder, workloads as wl
    host_mode = os.environ.get("ETL_TEST_HOST_EXCHANGE") == "1"    # every rank on GPU 0, the exchanges carried by gloo
    dev_index = 0 if host_mode else rank
    torch.cuda.set_device(dev_index)
    if host_mode:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    w = wl.make(name, scale, n_segments=1)
    if int(os.environ.get("ETL_TEST_BUMP", "0")):
        w.schema_bump_ppm = int(os.environ["ETL_TEST_BUMP"])
    stream, _ = w.generate()
    cuts = np.load(os.path.join(out_dir, "cuts.npy")).tolist()
    shard = stream[cuts[rank]:cuts[rank + 1]]
    dec = decoder.Decoder(dev_index)
    for tid, cols in w.table_schemas().items():
        dec.put_table_schema(tid, cols)
    if host_mode:
        def allgather(send: bytes) -> bytes:
            t = torch.frombuffer(bytearray(send), dtype=torch.uint8)
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return b"".join(o.numpy().tobytes() for o in outs)
        dec.comm_init_host(rank, world, allgather)
    else:
        uid = [dec.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dec.comm_init(uid[0], rank, world)
    st = decoder.Stager(max(shard.nbytes, 1), 2048)
    st.append_framed(shard)
    for rep in range(2):                                # second pass: optimistic sizing + carried relation state
        if rep:
            dec.reset_relations()
        with dec.decode_sharded(st.view(), to_host=True) as bh:
            p = bh.to_host()
        print(f"[rank {rank}] rep {rep}: records {p.n_records} first_error {p.first_error} base {p.record_index_base} carry_out {p.carry_out}", flush=True)
    if os.environ.get("ETL_TEST_DIAG"):                 # the same range through the two-phase entry points, carry taken from the library's own answer
        from etl_b200 import sharding
        seam = dec.decode_begin(st.view(), to_host=True)
        words = torch.from_numpy(sharding.seam_to_words(seam).view(np.int64).copy())
        if not host_mode:
            words = words.cuda()
        allw = [torch.zeros_like(words) for _ in range(world)]
        dist.all_gather(allw, words)
        allw = np.stack([w_.cpu().numpy().view(np.uint64) for w_ in allw])
        state, base = sharding.carry_for_rank(allw, rank)
        with dec.decode_finish(state, base) as bh:
            q = bh.to_host()
        print(f"[rank {rank}] two-phase: records {q.n_records} first_error {q.first_error} carry {state} base {base}", flush=True)
        for f in ("rec_kind", "rec_flags", "rec_schema", "rec_commit_lsn", "rec_tx_ordinal", "rec_cell_base", "cell_tag", "cell_aux"):
            a_, b_ = getattr(p, f), getattr(q, f)
            n_ = min(len(a_), len(b_))
            d_ = np.nonzero(a_[:n_] != b_[:n_])[0]
            print(f"[rank {rank}] {f}: len {len(a_)} vs {len(b_)}, first diff {d_[:3]}", flush=True)
    fields = {k: getattr(p, k) for k in ("rec_off", "rec_kind", "rec_flags", "rec_rel", "rec_schema", "rec_start_lsn", "rec_commit_lsn",
                                         "rec_tx_ordinal", "rec_cell_base", "rec_tuple_bytes", "rec_heap_hint", "cell_tag", "cell_val",
                                         "cell_aux", "heap")}
    fe = p.first_error
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **fields,
             meta=np.array([p.n_records, p.n_cells, -1 if fe[0] is None else fe[0], fe[1], fe[2], fe[3], p.carry_out[0], p.carry_out[1],
                            p.carry_out[2], p.insert_bytes, p.update_bytes, p.delete_bytes, p.n_events, p.record_index_base], dtype=np.int64),
             schema_tables=np.array([s.table_id for s in p.schemas], dtype=np.int64),
             schema_offs=np.array([s.effective_off for s in p.schemas], dtype=np.int64),
             schema_ident=np.array([s.n_identity for s in p.schemas], dtype=np.int64))
    dec.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
