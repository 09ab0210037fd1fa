"""Columnar emitter (etl_dec_arrow_emit, csrc/arrow_emit.cu) against a numpy restatement of the reference's Arrow
encoders (crates/etl-destinations/src/iceberg/encoding.rs:61-330) evaluated on the ORACLE's planes."""
import ctypes as C

import numpy as np
import pytest

from etl_b200 import abi, workloads as wl

pytestmark = pytest.mark.gpu

(A_UNSUP, A_BOOL, A_I32, A_I64, A_F32, A_F64, A_UTF8, A_LBIN, A_DATE32, A_TIME64, A_TS, A_TSTZ, A_UUID) = range(13)
KIND2ARROW = {1: A_BOOL, 2: A_UTF8, 3: A_I32, 4: A_I32, 5: A_I64, 6: A_I64, 7: A_F32, 8: A_F64, 10: A_DATE32, 11: A_TIME64, 12: A_TS, 13: A_TSTZ,
              14: A_UUID, 16: A_LBIN}


def expected_columns(p, stream: bytes, schema_index: int, row_kinds: int):
    """(row record indices, per column (arrow_type, validity bool[], values / (offsets, data)))"""
    sc = p.schemas[schema_index]
    heap = p.heap.tobytes()
    rows = []
    for r in range(p.n_records):
        if int(p.rec_schema[r]) != schema_index or not int(p.rec_flags[r]) & 0x80:
            continue
        k, f = chr(int(p.rec_kind[r])), int(p.rec_flags[r])
        c0, c1 = int(p.rec_cell_base[r]), int(p.rec_cell_base[r + 1])
        if k == "I" and row_kinds & 1:
            rows.append((r, c0))
        elif k == "U" and row_kinds & 2 and not f & 4:
            rows.append((r, c1 - sc.n_cols))
        elif k == "D" and row_kinds & 4 and f & 1:
            rows.append((r, c0))
    cols = []
    for c in range(sc.n_cols):
        at = KIND2ARROW.get(int(sc.col_kind[c]), A_UNSUP)
        if at == A_UNSUP:
            cols.append((at, None, None))
            continue
        tags = np.array([int(p.cell_tag[c0 + c]) for _, c0 in rows], dtype=np.int64)
        vals = np.array([int(p.cell_val[c0 + c]) for _, c0 in rows], dtype=np.uint64)
        auxs = np.array([int(p.cell_aux[c0 + c]) for _, c0 in rows], dtype=np.int64)
        want_tags = {A_BOOL: (1,), A_I32: (3, 4), A_I64: (5, 6), A_F32: (7,), A_F64: (8,), A_UTF8: (2,), A_LBIN: (16,), A_DATE32: (10,), A_TIME64: (11,),
                     A_TS: (12,), A_TSTZ: (13,), A_UUID: (14,)}[at]
        valid = np.isin(tags, want_tags)
        sv = vals.view(np.int64)
        if at == A_BOOL:
            v = (vals & 1).astype(bool) & valid
        elif at in (A_I32, A_DATE32):
            v = np.where(valid, sv, 0).astype(np.int32)
        elif at == A_F32:
            v = np.where(valid, vals & 0xFFFFFFFF, 0).astype(np.uint32)
        elif at in (A_I64, A_F64):
            v = np.where(valid, np.where(tags == 5, vals & 0xFFFFFFFF, vals).view(np.int64) if at == A_I64 else sv, 0).astype(np.int64)
        elif at in (A_TIME64, A_TS, A_TSTZ):
            v = np.where(valid, sv * 1000000 + auxs // 1000, 0).astype(np.int64)
        elif at == A_UUID:
            v = b"".join(heap[int(x):int(x) + 16] if ok else b"\0" * 16 for x, ok in zip(vals, valid))
        else:
            src = stream if at == A_UTF8 else heap
            chunks = [bytes(src[int(x):int(x) + int(a)]) if ok else b"" for x, a, ok in zip(vals, auxs, valid)]
            offs = np.zeros(len(rows) + 1, dtype=np.int64)
            offs[1:] = np.cumsum([len(ch) for ch in chunks])
            v = (offs, b"".join(chunks))
        cols.append((at, valid, v))
    return np.array([r for r, _ in rows], dtype=np.uint64), cols


def bits(ptr, n):
    raw = np.frombuffer((C.c_uint8 * ((n + 7) // 8)).from_address(ptr), dtype=np.uint8) if n else np.zeros(0, np.uint8)
    return np.unpackbits(raw, bitorder="little")[:n].astype(bool)


@pytest.mark.parametrize("name,scale,kinds", [("c2", 0.01, 1), ("c2", 0.01, 7), ("c4", 0.002, 3), ("c3", 0.001, 3), ("c5", 0.002, 3)])
def test_arrow_columns_match_reference_encoders(oracle_mod, name, scale, kinds):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from etl_b200 import decoder
    w = wl.make(name, scale, n_segments=1)
    stream, _ = w.generate()
    raw = stream.tobytes()
    orc = oracle_mod.Oracle()
    dec = decoder.Decoder(0)
    for tid, cols in w.table_schemas().items():
        orc.put_table_schema(tid, cols)
        dec.put_table_schema(tid, cols)
    want_planes = orc.decode(raw)
    st = decoder.Stager(stream.nbytes, 2048)
    st.append_framed(stream)
    lib = abi.load()
    with dec.decode_input(st.view(), to_host=True) as bh:
        for si in range(min(len(want_planes.schemas), 6)):
            recs, cols = expected_columns(want_planes, raw, si, kinds)
            a = C.c_void_p()
            assert lib.etl_dec_arrow_emit(bh._h, si, kinds, 1, C.byref(a)) == 0
            n = lib.etl_dec_arrow_rows(a)
            assert n == len(recs)
            assert lib.etl_dec_arrow_cols(a) == len(cols)
            if n:
                got_recs = np.frombuffer((C.c_uint64 * n).from_address(lib.etl_dec_arrow_row_records(a, 1)), dtype=np.uint64)
                assert np.array_equal(got_recs, recs)
            for c, (at, valid, v) in enumerate(cols):
                col = abi.ArrowColumn()
                assert lib.etl_dec_arrow_column(a, c, 1, C.byref(col)) == 0
                assert col.arrow_type == at, (c, col.arrow_type, at)
                if at == A_UNSUP or n == 0:
                    continue
                assert np.array_equal(bits(col.validity, n), valid), f"validity of column {c}"
                if at == A_BOOL:
                    assert np.array_equal(bits(col.values, n), v)
                elif at in (A_I32, A_DATE32):
                    assert np.array_equal(np.frombuffer((C.c_int32 * n).from_address(col.values), dtype=np.int32), v)
                elif at == A_F32:
                    assert np.array_equal(np.frombuffer((C.c_uint32 * n).from_address(col.values), dtype=np.uint32), v)
                elif at in (A_I64, A_F64, A_TIME64, A_TS, A_TSTZ):
                    assert np.array_equal(np.frombuffer((C.c_int64 * n).from_address(col.values), dtype=np.int64), v)
                elif at == A_UUID:
                    assert bytes((C.c_uint8 * (16 * n)).from_address(col.values)) == v
                else:
                    offs, data = v
                    ot = C.c_int32 if at == A_UTF8 else C.c_int64
                    got_offs = np.frombuffer((ot * (n + 1)).from_address(col.offsets), dtype=np.int32 if at == A_UTF8 else np.int64)
                    assert np.array_equal(got_offs.astype(np.int64), offs), f"offsets of column {c}"
                    assert col.data_bytes == len(data)
                    assert bytes((C.c_uint8 * len(data)).from_address(col.data)) == data if data else True
            lib.etl_dec_arrow_free(a)
    st.close()
    dec.close()
