"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by etl_b200/ (the product).  See oracle/oracle.h for the parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle_cells.c", "oracle_stream.c", "oracle_copy.c", "oracle_digest.c", "oracle.h",
                                              "oracle_internal.h", "../include/etl_decode.h")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class ColumnSchema(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type_oid", C.c_uint32), ("modifier", C.c_int32),
                ("ordinal_position", C.c_int32), ("primary_key_ordinal_position", C.c_int32),
                ("nullable", C.c_uint8), ("_pad", C.c_uint8 * 7)]


class StreamState(C.Structure):
    _fields_ = [("final_lsn", C.c_uint64), ("next_tx_ordinal", C.c_uint64), ("in_tx", C.c_uint8),
                ("_pad", C.c_uint8 * 7)]


class FirstError(C.Structure):
    _fields_ = [("record_index", C.c_uint64), ("seq", C.c_uint32), ("code", C.c_uint32),
                ("kind", C.c_uint32), ("_pad", C.c_uint32)]


class OrcSchema(C.Structure):
    _fields_ = [("table_id", C.c_uint32), ("n_cols", C.c_uint32), ("n_identity", C.c_uint32),
                ("_pad", C.c_uint32), ("snapshot_id", C.c_uint64), ("effective_off", C.c_uint64),
                ("col_kind", C.POINTER(C.c_uint8)), ("col_flags", C.POINTER(C.c_uint8)),
                ("col_index", C.POINTER(C.c_int32))]


class OrcBatch(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_cells", C.c_uint64), ("heap_bytes", C.c_uint64),
                ("rec_off", C.POINTER(C.c_uint64)), ("rec_kind", C.POINTER(C.c_uint8)),
                ("rec_flags", C.POINTER(C.c_uint8)), ("rec_rel", C.POINTER(C.c_uint32)),
                ("rec_schema", C.POINTER(C.c_int32)), ("rec_start_lsn", C.POINTER(C.c_uint64)),
                ("rec_commit_lsn", C.POINTER(C.c_uint64)), ("rec_tx_ordinal", C.POINTER(C.c_uint64)),
                ("rec_cell_base", C.POINTER(C.c_uint64)), ("rec_tuple_bytes", C.POINTER(C.c_uint32)),
                ("rec_heap_hint", C.POINTER(C.c_uint32)), ("cell_tag", C.POINTER(C.c_uint8)),
                ("cell_val", C.POINTER(C.c_uint64)), ("cell_aux", C.POINTER(C.c_uint32)),
                ("heap", C.POINTER(C.c_uint8)), ("first_error", FirstError),
                ("carry_out", StreamState), ("insert_bytes", C.c_uint64),
                ("update_bytes", C.c_uint64), ("delete_bytes", C.c_uint64), ("n_events", C.c_uint64),
                ("n_schemas", C.c_uint32), ("_pad", C.c_uint32), ("schemas", C.POINTER(OrcSchema)),
                ("cap_records", C.c_uint64), ("cap_cells", C.c_uint64), ("cap_heap", C.c_uint64),
                ("cap_schemas", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_put_table_schema.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(ColumnSchema), C.c_uint32]
        L.orc_reset_relations.argtypes = [C.c_void_p]
        L.orc_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(StreamState), C.POINTER(OrcBatch)]
        L.orc_batch_free.argtypes = [C.POINTER(OrcBatch)]
        L.orc_parse_cell.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint8),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_void_p,
                                     C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_parse_cell.restype = C.c_uint32
        L.orc_kind_for_oid.argtypes = [C.c_uint32]
        L.orc_kind_for_oid.restype = C.c_uint32
        L.orc_error_kind.argtypes = [C.c_uint32]
        L.orc_error_kind.restype = C.c_uint32
        L.orc_planes_digest.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_planes_digest.restype = None
        L.orc_batch_digest.argtypes = [C.POINTER(OrcBatch), C.POINTER(C.c_uint64)]
        L.orc_batch_digest.restype = None
        _lib = L
    return _lib


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).view(dtype).copy()


@dataclass
class SchemaInfo:
    table_id: int
    n_cols: int
    n_identity: int
    snapshot_id: int
    effective_off: int
    col_kind: np.ndarray
    col_flags: np.ndarray
    col_index: np.ndarray


@dataclass
class Planes:
    """Canonical decoded batch (same layout as etl_dec_planes in include/etl_decode.h)."""
    n_records: int
    n_cells: int
    rec_off: np.ndarray
    rec_kind: np.ndarray
    rec_flags: np.ndarray
    rec_rel: np.ndarray
    rec_schema: np.ndarray
    rec_start_lsn: np.ndarray
    rec_commit_lsn: np.ndarray
    rec_tx_ordinal: np.ndarray
    rec_cell_base: np.ndarray
    rec_tuple_bytes: np.ndarray
    rec_heap_hint: np.ndarray
    cell_tag: np.ndarray
    cell_val: np.ndarray
    cell_aux: np.ndarray
    heap: np.ndarray
    first_error: tuple  # (record_index or None, seq, code, kind)
    carry_out: tuple    # (in_tx, final_lsn, next_tx_ordinal)
    insert_bytes: int
    update_bytes: int
    delete_bytes: int
    n_events: int
    schemas: List[SchemaInfo]


def make_columns(cols: Sequence[dict]):
    arr = (ColumnSchema * max(1, len(cols)))()
    keep = []
    for i, c in enumerate(cols):
        nm = c["name"].encode()
        keep.append(nm)
        arr[i].name = nm
        arr[i].type_oid = c["type_oid"]
        arr[i].modifier = c.get("modifier", -1)
        arr[i].ordinal_position = c.get("ordinal_position", i + 1)
        pk = c.get("pk")
        arr[i].primary_key_ordinal_position = -1 if pk is None else pk
        arr[i].nullable = 1 if c.get("nullable", True) else 0
    return arr, keep


class Oracle:
    def __init__(self):
        self._l = lib()
        self._ctx = C.c_void_p(self._l.orc_create())

    def close(self):
        if self._ctx:
            self._l.orc_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def put_table_schema(self, table_id: int, cols: Sequence[dict], snapshot_id: int = 0):
        arr, _keep = make_columns(cols)
        self._l.orc_put_table_schema(self._ctx, table_id, snapshot_id, arr, len(cols))

    def reset_relations(self):
        self._l.orc_reset_relations(self._ctx)

    def decode_raw(self, buf, carry_in: Optional[tuple] = None) -> OrcBatch:
        """Decode without copying results out (for timing). Caller must free()."""
        st = StreamState()
        if carry_in:
            st.in_tx, st.final_lsn, st.next_tx_ordinal = int(carry_in[0]), carry_in[1], carry_in[2]
        b = OrcBatch()
        if isinstance(buf, np.ndarray):
            ptr, n = buf.ctypes.data, buf.nbytes
        else:
            self._keep = buf
            ptr, n = C.cast(C.c_char_p(buf), C.c_void_p).value, len(buf)
        self._l.orc_decode(self._ctx, ptr, n, C.byref(st), C.byref(b))
        return b

    def digest(self, buf, carry_in: Optional[tuple] = None):
        """Decode and return (canonical digest hex, n_records, first_error record or None) without copying planes out."""
        b = self.decode_raw(buf, carry_in)
        try:
            out = (C.c_uint64 * 4)()
            self._l.orc_batch_digest(C.byref(b), out)
            fe = b.first_error.record_index
            return "".join("%016x" % v for v in out), int(b.n_records), (None if fe == 2**64 - 1 else int(fe))
        finally:
            self.free(b)

    def free(self, b: OrcBatch):
        self._l.orc_batch_free(C.byref(b))

    def decode(self, buf, carry_in: Optional[tuple] = None) -> Planes:
        b = self.decode_raw(buf, carry_in)
        try:
            n, m = b.n_records, b.n_cells
            fe = b.first_error
            schemas = []
            for i in range(b.n_schemas):
                s = b.schemas[i]
                schemas.append(SchemaInfo(s.table_id, s.n_cols, s.n_identity, s.snapshot_id, s.effective_off,
                                          _arr(s.col_kind, s.n_cols, np.uint8), _arr(s.col_flags, s.n_cols, np.uint8),
                                          _arr(s.col_index, s.n_cols, np.int32)))
            return Planes(
                n_records=int(n), n_cells=int(m),
                rec_off=_arr(b.rec_off, n, np.uint64), rec_kind=_arr(b.rec_kind, n, np.uint8),
                rec_flags=_arr(b.rec_flags, n, np.uint8), rec_rel=_arr(b.rec_rel, n, np.uint32),
                rec_schema=_arr(b.rec_schema, n, np.int32), rec_start_lsn=_arr(b.rec_start_lsn, n, np.uint64),
                rec_commit_lsn=_arr(b.rec_commit_lsn, n, np.uint64), rec_tx_ordinal=_arr(b.rec_tx_ordinal, n, np.uint64),
                rec_cell_base=_arr(b.rec_cell_base, n + 1, np.uint64), rec_tuple_bytes=_arr(b.rec_tuple_bytes, n, np.uint32),
                rec_heap_hint=_arr(b.rec_heap_hint, n, np.uint32), cell_tag=_arr(b.cell_tag, m, np.uint8),
                cell_val=_arr(b.cell_val, m, np.uint64), cell_aux=_arr(b.cell_aux, m, np.uint32),
                heap=_arr(b.heap, b.heap_bytes, np.uint8),
                first_error=(None if fe.record_index == 2**64 - 1 else int(fe.record_index), int(fe.seq), int(fe.code), int(fe.kind)),
                carry_out=(int(b.carry_out.in_tx), int(b.carry_out.final_lsn), int(b.carry_out.next_tx_ordinal)),
                insert_bytes=int(b.insert_bytes), update_bytes=int(b.update_bytes), delete_bytes=int(b.delete_bytes),
                n_events=int(b.n_events), schemas=schemas)
        finally:
            self.free(b)


def planes_digest(planes_struct, n_valid: int) -> str:
    """Canonical digest of an etl_dec_planes struct (ctypes, host pointers) — the checker side of the parity leg."""
    out = (C.c_uint64 * 4)()
    lib().orc_planes_digest(C.byref(planes_struct), n_valid, out)
    return "".join("%016x" % v for v in out)


def copy_rows_digest(type_oids: Sequence[int], buf: np.ndarray, row_off: np.ndarray):
    """Oracle decode of COPY rows, row by row (table_row.rs:25-165): (digest of the rows before the first failing one,
    first error (row, col, code) or None)."""
    L = lib()
    oids = (C.c_uint32 * max(len(type_oids), 1))(*type_oids)
    out = (C.c_uint64 * 4)()
    er, ec, ee = C.c_uint64(), C.c_uint32(), C.c_uint32()
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    row_off = np.ascontiguousarray(row_off, dtype=np.uint64)
    L.orc_copy_rows_digest.restype = None
    L.orc_copy_rows_digest(oids, C.c_uint32(len(type_oids)), C.c_void_p(buf.ctypes.data), C.c_void_p(row_off.ctypes.data), C.c_uint64(len(row_off) - 1), out,
                           C.byref(er), C.byref(ec), C.byref(ee))
    err = None if er.value == 2**64 - 1 else (int(er.value), int(ec.value), int(ee.value))
    return "".join("%016x" % v for v in out), err


def copy_planes_digest(tags: np.ndarray, vals: np.ndarray, auxs: np.ndarray, n_valid_rows: int, n_cols: int, stream: np.ndarray, heap: np.ndarray) -> str:
    L = lib()
    out = (C.c_uint64 * 4)()
    heap = np.ascontiguousarray(heap if heap.nbytes else np.zeros(8, np.uint8))
    stream = np.ascontiguousarray(stream if stream.nbytes else np.zeros(8, np.uint8))
    L.orc_copy_planes_digest.restype = None
    L.orc_copy_planes_digest(C.c_void_p(tags.ctypes.data), C.c_void_p(vals.ctypes.data), C.c_void_p(auxs.ctypes.data), C.c_uint64(n_valid_rows),
                             C.c_uint32(n_cols), C.c_void_p(stream.ctypes.data), C.c_void_p(heap.ctypes.data), out)
    return "".join("%016x" % v for v in out)


def parse_cell(type_oid: int, text: bytes):
    """text.rs:28 for one value → (err_code, tag, val, aux, heap bytes)."""
    L = lib()
    tag = C.c_uint8()
    val = C.c_uint64()
    aux = C.c_uint32()
    hl = C.c_uint32()
    cap = 16 * len(text) + 1024
    heap = (C.c_uint8 * cap)()
    e = L.orc_parse_cell(type_oid, text, len(text), C.byref(tag), C.byref(val), C.byref(aux), heap, cap, C.byref(hl))
    return e, tag.value, val.value, aux.value, bytes(heap[:min(hl.value, cap)])


def kind_for_oid(type_oid: int) -> int:
    """ETL_K_* decode class the reference's `Type` dispatch gives this oid (oracle_cells.c)."""
    return int(lib().orc_kind_for_oid(type_oid))


E_COPY_NOT_TERMINATED, E_COPY_COLUMN_COUNT = 101, 102


def parse_copy_row(type_oids: Sequence[int], row: bytes):
    """table_row.rs:25-165 for one COPY-text row → (err_code, err_col, cells) with cells as
    [(tag, val, aux)], plus the unescaped text plane and the heap the cells point into."""
    L = lib()
    if not getattr(L, "_copy_bound", False):
        L.orc_parse_copy_row.restype = C.c_uint32
        L.orc_parse_copy_row.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint8),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
                                         C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint32)]
        L._copy_bound = True
    n = len(type_oids)
    oids = (C.c_uint32 * max(n, 1))(*type_oids)
    tags, vals, auxs = (C.c_uint8 * max(n, 1))(), (C.c_uint64 * max(n, 1))(), (C.c_uint32 * max(n, 1))()
    nv, ecol = C.c_uint32(), C.c_uint32()
    tcap, hcap = len(row) + 16, 16 * len(row) + 1024
    text, heap = (C.c_uint8 * tcap)(), (C.c_uint8 * hcap)()
    tl, hl = C.c_uint64(), C.c_uint64()
    e = L.orc_parse_copy_row(oids, n, row, len(row), tags, vals, auxs, C.byref(nv), text, tcap, C.byref(tl), heap, hcap, C.byref(hl), C.byref(ecol))
    cells = [(tags[i], vals[i], auxs[i]) for i in range(nv.value)]
    return e, (None if ecol.value == 0xFFFFFFFF else ecol.value), cells, bytes(text[:tl.value]), bytes(heap[:min(hl.value, hcap)])
