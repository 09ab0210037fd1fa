"""Canonical views over a decoded batch (the plane layout of include/etl_decode.h).

`planes_to_events` turns the record/cell planes into a list of python dicts shaped like the
reference's `Event` enum (crates/etl/src/types/event.rs:242-260) so known-answer tests read like
the reference's own assertions.  `assert_planes_equal` is the bulk bit-exact comparator used by the
parity tests (var-width payloads are compared by dereference, not by heap offset).
"""
from __future__ import annotations

import struct
from typing import Any, List, Optional

import numpy as np

CELL_NULL, CELL_BOOL, CELL_STRING, CELL_I16, CELL_I32, CELL_U32, CELL_I64, CELL_F32, CELL_F64 = range(9)
CELL_NUMERIC, CELL_DATE, CELL_TIME, CELL_TIMESTAMP, CELL_TIMESTAMPTZ, CELL_UUID, CELL_JSON, CELL_BYTES, CELL_ARRAY = range(9, 18)
CELL_MISSING = 254

RF_OLD_FULL, RF_OLD_KEY, RF_NEW_PARTIAL, RF_DDL, RF_EVENT = 1, 2, 4, 8, 0x80


def _i64(v: int) -> int:
    v = int(v)
    return v - (1 << 64) if v >= (1 << 63) else v


def numeric_from_heap(heap: bytes, off: int, nd: int, with_pushed: bool = False):
    kind, sign, weight, scale, pushed = struct.unpack_from("<BBhHH", heap, off)
    if kind == 1:
        return ("numeric", "NaN")
    if kind == 2:
        return ("numeric", "Infinity")
    if kind == 3:
        return ("numeric", "-Infinity")
    digits = list(struct.unpack_from("<%dh" % nd, heap, off + 8)) if nd else []
    if with_pushed:                                    # groups pushed before the zero strips (the digit Vec's capacity)
        return ("numeric", "-" if sign else "+", weight, scale, digits, pushed)
    return ("numeric", "-" if sign else "+", weight, scale, digits)


def decode_cell(tag: int, val: int, aux: int, stream: bytes, heap: bytes, in_array: bool = False, strict: bool = False) -> Any:
    if tag == CELL_NULL:
        return None
    if tag == CELL_MISSING:
        return ("missing",)
    if tag == CELL_BOOL:
        return bool(val)
    if tag in (CELL_I16, CELL_I32, CELL_I64):
        return _i64(val)
    if tag == CELL_U32:
        return ("u32", int(val))
    if tag == CELL_F32:
        return ("f32", int(val) & 0xFFFFFFFF)
    if tag == CELL_F64:
        return ("f64", int(val))
    if tag == CELL_STRING:
        src = heap if in_array else stream
        return bytes(src[val:val + aux]).decode("utf-8")
    if tag == CELL_JSON:
        src = heap if in_array else stream
        return ("json", bytes(src[val:val + aux]))
    if tag == CELL_NUMERIC:
        return numeric_from_heap(heap, val, aux, with_pushed=strict and not in_array)
    if tag == CELL_DATE:
        return ("date", _i64(val))
    if tag == CELL_TIME:
        return ("time", int(val), int(aux))
    if tag == CELL_TIMESTAMP:
        return ("timestamp", _i64(val), int(aux))
    if tag == CELL_TIMESTAMPTZ:
        return ("timestamptz", _i64(val), int(aux))
    if tag == CELL_UUID:
        return ("uuid", bytes(heap[val:val + 16]).hex())
    if tag == CELL_BYTES:
        return ("bytes", bytes(heap[val:val + aux]))
    if tag == CELL_ARRAY:
        ek, n = struct.unpack_from("<B3xI", heap, val)
        out = []
        for k in range(n):
            ev, ea, et = struct.unpack_from("<QIB3x", heap, val + 8 + 16 * k)
            out.append(decode_cell(et, ev, ea, stream, heap, in_array=True))
        return ("array", ek, out)
    raise ValueError(f"unknown cell tag {tag}")


def cells_of(p, rec: int, stream: bytes) -> List[Any]:
    heap = bytes(p.heap.tobytes()) if isinstance(p.heap, np.ndarray) else bytes(p.heap)
    a, b = int(p.rec_cell_base[rec]), int(p.rec_cell_base[rec + 1])
    return [decode_cell(int(p.cell_tag[i]), int(p.cell_val[i]), int(p.cell_aux[i]), stream, heap) for i in range(a, b)]


def planes_to_events(p, stream: bytes, limit: Optional[int] = None) -> List[dict]:
    """Event list up to (not including) the first error."""
    n = p.n_records if p.first_error[0] is None else p.first_error[0]
    if limit is not None:
        n = min(n, limit)
    out = []
    for r in range(n):
        kind = chr(int(p.rec_kind[r]))
        flags = int(p.rec_flags[r])
        if not flags & RF_EVENT:
            continue
        base = dict(start_lsn=int(p.rec_start_lsn[r]), commit_lsn=int(p.rec_commit_lsn[r]),
                    tx_ordinal=int(p.rec_tx_ordinal[r]))
        cells = cells_of(p, r, stream)
        if kind == "B":
            out.append(dict(kind="begin", timestamp=cells[0], xid=cells[1][1], **base))
        elif kind == "C":
            out.append(dict(kind="commit", flags=cells[0], end_lsn=cells[1] & (2**64 - 1), timestamp=cells[2], **base))
        elif kind == "R":
            out.append(dict(kind="relation", table_id=int(p.rec_rel[r]), schema=int(p.rec_schema[r]), **base))
        elif kind == "T":
            out.append(dict(kind="truncate", options=cells[0], rel_ids=[c[1] for c in cells[1:]], **base))
        else:
            sch = p.schemas[int(p.rec_schema[r])]
            ev = dict(table_id=int(p.rec_rel[r]), schema=int(p.rec_schema[r]), **base)
            n_old = sch.n_cols if flags & RF_OLD_FULL else (sch.n_identity if flags & RF_OLD_KEY else 0)
            old = None
            if flags & RF_OLD_FULL:
                old = ("full", cells[:n_old])
            elif flags & RF_OLD_KEY:
                old = ("key", cells[:n_old])
            if kind == "I":
                ev.update(kind="insert", row=cells)
            elif kind == "U":
                new = cells[n_old:]
                if flags & RF_NEW_PARTIAL:
                    present = [c for c in new if c != ("missing",)]
                    missing = [i for i, c in enumerate(new) if c == ("missing",)]
                    ev.update(kind="update", row=("partial", len(new), present, missing), old=old)
                else:
                    ev.update(kind="update", row=("full", new), old=old)
            else:
                ev.update(kind="delete", old=old)
            out.append(ev)
    return out


_REC_FIELDS = ["rec_off", "rec_kind", "rec_flags", "rec_rel", "rec_schema", "rec_start_lsn",
               "rec_commit_lsn", "rec_tx_ordinal", "rec_tuple_bytes", "rec_heap_hint"]
_VAR_TAGS = (CELL_NUMERIC, CELL_UUID, CELL_BYTES, CELL_ARRAY)


def assert_planes_equal(got, want, stream: bytes, check_heap_contents: bool = True):
    """Bit-exact comparison of two decoded batches over the valid prefix."""
    assert got.first_error == want.first_error, f"first_error {got.first_error} != {want.first_error}"
    n = want.n_records if want.first_error[0] is None else want.first_error[0]
    if want.first_error[0] is None:
        assert got.n_records == want.n_records, (got.n_records, want.n_records)
        assert got.carry_out == want.carry_out, (got.carry_out, want.carry_out)
        assert (got.insert_bytes, got.update_bytes, got.delete_bytes, got.n_events) == \
            (want.insert_bytes, want.update_bytes, want.delete_bytes, want.n_events)
    for f in _REC_FIELDS:
        a, b = getattr(got, f)[:n], getattr(want, f)[:n]
        if not np.array_equal(a, b):
            i = int(np.nonzero(a != b)[0][0])
            raise AssertionError(f"{f} differs at record {i}: got {a[i]} want {b[i]} (kind {chr(int(want.rec_kind[i]))})")
    assert np.array_equal(got.rec_cell_base[:n + 1], want.rec_cell_base[:n + 1]), "rec_cell_base differs"
    m = int(want.rec_cell_base[n])
    gt, wt = got.cell_tag[:m], want.cell_tag[:m]
    if not np.array_equal(gt, wt):
        i = int(np.nonzero(gt != wt)[0][0])
        raise AssertionError(f"cell_tag differs at cell {i}: got {gt[i]} want {wt[i]}")
    fixed = ~np.isin(wt, _VAR_TAGS)
    gv, wv = got.cell_val[:m], want.cell_val[:m]
    if not np.array_equal(gv[fixed], wv[fixed]):
        idx = np.nonzero(fixed)[0]
        i = int(idx[np.nonzero(gv[fixed] != wv[fixed])[0][0]])
        raise AssertionError(f"cell_val differs at cell {i} (tag {wt[i]}): got {gv[i]} want {wv[i]}")
    ga, wa = got.cell_aux[:m], want.cell_aux[:m]
    if not np.array_equal(ga, wa):
        i = int(np.nonzero(ga != wa)[0][0])
        raise AssertionError(f"cell_aux differs at cell {i} (tag {wt[i]}): got {ga[i]} want {wa[i]}")
    if check_heap_contents:
        gh, wh = got.heap.tobytes(), want.heap.tobytes()
        for i in np.nonzero(~fixed)[0]:
            i = int(i)
            a = decode_cell(int(gt[i]), int(gv[i]), int(ga[i]), stream, gh, strict=True)
            b = decode_cell(int(wt[i]), int(wv[i]), int(wa[i]), stream, wh, strict=True)
            if a != b:
                raise AssertionError(f"var cell {i} (tag {wt[i]}) differs: got {a} want {b}")
    # schema versions (the host installs every Relation of the batch up front; after a data error
    # only the records before it are specified, so versions are compared for clean batches only)
    if want.first_error[0] is not None:
        return
    assert len(got.schemas) == len(want.schemas), (len(got.schemas), len(want.schemas))
    for i, (a, b) in enumerate(zip(got.schemas, want.schemas)):
        assert (a.table_id, a.n_cols, a.n_identity, a.effective_off) == (b.table_id, b.n_cols, b.n_identity, b.effective_off), f"schema {i}"
        assert np.array_equal(a.col_kind, b.col_kind) and np.array_equal(a.col_flags, b.col_flags) and np.array_equal(a.col_index, b.col_index), f"schema {i} columns"
