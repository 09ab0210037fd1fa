"""Event-level scenarios shared by the oracle tests (CPU) and the GPU parity tests.

Each scenario = (name, stored table schemas, stream builder) and — where the reference holds a
known answer — the expected event list.  The schemas and tuples restate the fixtures of
crates/etl/src/conversions/event.rs:1010-1066 and the tests at :1301-1637.
"""
from __future__ import annotations

from etl_b200 import pgoutput as pg
from etl_b200.pgoutput import UNCHANGED, Binary

INT8, TEXT, DATE, INT4, BOOL, TIMESTAMPTZ, NUMERIC, JSONB, UUID, BYTEA, FLOAT8 = 20, 25, 1082, 23, 16, 1184, 1700, 3802, 2950, 17, 701


def col(name, oid, pk=None, nullable=False):
    return dict(name=name, type_oid=oid, pk=pk, nullable=nullable)


# event.rs:1020-1028 composite_primary_key_schema (table 42): identity = pk columns (id, surname)
COMPOSITE_PK = [col("id", INT8, 2), col("name", TEXT), col("surname", TEXT, 1), col("city", TEXT), col("large_text", TEXT)]
# event.rs:1030-1047 alternative_identity_schema (table 43): identity mask [0,1,1,0]
ALT_IDENTITY = [col("id", INT8, 2), col("name", TEXT), col("surname", TEXT, 1), col("city", TEXT)]


def rel_cols(cols, identity_names):
    return [(1 if c["name"] in identity_names else 0, c["name"], c["type_oid"], -1) for c in cols]


class Tx:
    """Helper that wraps messages in Begin/Commit with consistent LSNs."""

    def __init__(self, w: pg.StreamWriter):
        self.w = w

    def begin(self, final_lsn=None, xid=777):
        self.final = final_lsn if final_lsn is not None else self.w.lsn + 0x10000
        self.w.emit(pg.begin(self.final, self.w.clock, xid))

    def commit(self, commit_lsn=None):
        c = self.final if commit_lsn is None else commit_lsn
        self.w.emit(pg.commit(0, c, c + 8, self.w.clock))


def stream_with(tables, relation_msgs, dml_msgs):
    """One transaction: Begin, relations, dml..., Commit."""
    w = pg.StreamWriter()
    tx = Tx(w)
    tx.begin()
    for r in relation_msgs:
        w.emit(r)
    for m in dml_msgs:
        w.emit(m)
    tx.commit()
    return w


def reference_update_delete_scenarios():
    """event.rs:1443-1637: (name, tables, writer, expected DML events [kind,row,old])."""
    out = []
    rel42 = pg.relation(42, "public", "test", "d", rel_cols(COMPOSITE_PK, {"id", "surname"}))
    rel43 = pg.relation(43, "public", "users", "i", rel_cols(ALT_IDENTITY, {"name", "surname"}))
    rel44 = pg.relation(44, "public", "users", "f", rel_cols(ALT_IDENTITY, set()))

    # :1444-1472 absent old row for non-identity change
    out.append(("update_no_old_row", {42: COMPOSITE_PK}, stream_with(None, [rel42], [
        pg.update(42, ["1", "alice", "smith", "vienna", "toast"])]),
        [dict(kind="update", old=None, row=("full", [1, "alice", "smith", "vienna", "toast"]))]))
    # :1475-1506 unrecoverable toast → partial
    out.append(("update_partial_toast", {42: COMPOSITE_PK}, stream_with(None, [rel42], [
        pg.update(42, ["1", "alice", "smith", "vienna", UNCHANGED])]),
        [dict(kind="update", old=None, row=("partial", 5, [1, "alice", "smith", "vienna"], [4]))]))
    # :1509-1539 key tuple (full width) for identity change
    out.append(("update_key_tuple", {42: COMPOSITE_PK}, stream_with(None, [rel42], [
        pg.update(42, ["1", "alice", "smithers", "rome", "toast"], key=["1", None, "smith", None, None])]),
        [dict(kind="update", old=("key", [1, "smith"]), row=("full", [1, "alice", "smithers", "rome", "toast"]))]))
    # :1542-1579 alternative identity, key tuple still sent
    out.append(("update_alt_identity_key", {43: ALT_IDENTITY}, stream_with(None, [rel43], [
        pg.update(43, ["1", "alice", "smith", "vienna"], key=[None, "alice", "smith", None])]),
        [dict(kind="update", old=("key", ["alice", "smith"]), row=("full", [1, "alice", "smith", "vienna"]))]))
    # :1582-1612 full identity old tuple
    out.append(("update_full_identity", {44: ALT_IDENTITY}, stream_with(None, [rel44], [
        pg.update(44, ["1", "alice", "smith", "vienna"], old=["1", "alice", "smith", "rome"])]),
        [dict(kind="update", old=("full", [1, "alice", "smith", "rome"]), row=("full", [1, "alice", "smith", "vienna"]))]))
    # :1615-1637 delete with key tuple
    out.append(("delete_key_tuple", {43: ALT_IDENTITY}, stream_with(None, [rel43], [
        pg.delete(43, key=[None, "alice", "smith", None])]),
        [dict(kind="delete", old=("key", ["alice", "smith"]))]))
    return out


def tuple_level_scenarios():
    """event.rs:1301-1441 restated through whole messages."""
    out = []
    two = [col("id", INT8, 1), col("d", DATE)]
    rel = pg.relation(50, "public", "t", "d", rel_cols(two, {"id"}))
    # :1301-1312 NOT NULL violation → InvalidData "Required column missing from tuple"
    out.append(("not_null_violation", {50: two}, stream_with(None, [rel], [pg.insert(50, ["1", None])]),
                ("error", 11)))
    pay = [col("id", INT8, 1), col("payload", TEXT)]
    relp = pg.relation(51, "public", "t", "d", rel_cols(pay, {"id"}))
    # :1315-1334 partial when toast cannot be recovered
    out.append(("partial_unrecoverable", {51: pay}, stream_with(None, [relp], [pg.update(51, ["1", UNCHANGED])]),
                [dict(kind="update", old=None, row=("partial", 2, [1], [1]))]))
    # :1337-1361 toast reused from Full old row
    relf = pg.relation(51, "public", "t", "f", rel_cols(pay, set()))
    out.append(("toast_from_full_old", {51: pay}, stream_with(None, [relf], [
        pg.update(51, ["1", UNCHANGED], old=["1", "toast"])]),
        [dict(kind="update", old=("full", [1, "toast"]), row=("full", [1, "toast"]))]))
    # :1364-1387 toast reused from Key row when the column is in the key (dense key tuple)
    pay2 = [col("id", INT8, 1), col("payload", TEXT, 2)]
    relk = pg.relation(52, "public", "t", "d", rel_cols(pay2, {"id", "payload"}))
    out.append(("toast_from_key_row", {52: pay2}, stream_with(None, [relk], [
        pg.update(52, ["2", UNCHANGED], key=["1", "toast"])]),
        [dict(kind="update", old=("key", [1, "toast"]), row=("full", [2, "toast"]))]))
    # :1390-1416 full-width key tuple filtered to identity columns [1,0,1,0]
    four = [col("id", INT8, 2), col("name", TEXT), col("surname", TEXT, 1), col("payload", TEXT)]
    rel4 = pg.relation(1, "public", "users", "d", rel_cols(four, {"id", "surname"}))
    out.append(("full_width_key", {1: four}, stream_with(None, [rel4], [
        pg.delete(1, key=["1", "alice", "smith", "toast"])]),
        [dict(kind="delete", old=("key", [1, "smith"]))]))
    # :1419-1441 dense key tuple
    out.append(("dense_key", {1: four}, stream_with(None, [rel4], [pg.delete(1, key=["1", "smith"])]),
                [dict(kind="delete", old=("key", [1, "smith"]))]))
    return out


def error_scenarios():
    """(name, tables, writer-or-bytes, expected (record_index, code)) — every data error the path raises."""
    out = []
    two = [col("id", INT8, 1), col("v", TEXT, None, True)]
    rel = pg.relation(60, "public", "t", "d", rel_cols(two, {"id"}))

    def one(msgs, rels=(rel,)):
        return stream_with(None, list(rels), msgs)

    out.append(("field_count_insert", {60: two}, one([pg.insert(60, ["1"])]), (2, 12)))
    out.append(("field_count_update_new", {60: two}, one([pg.update(60, ["1", "a", "b"])]), (2, 12)))
    out.append(("insert_unchanged_toast", {60: two}, one([pg.insert(60, ["1", UNCHANGED])]), (2, 13)))
    out.append(("binary_cell", {60: two}, one([pg.insert(60, ["1", Binary(b"ab")])]), (2, 10)))
    out.append(("bad_int", {60: two}, one([pg.insert(60, ["x1", "a"])]), (2, 2)))
    out.append(("bad_utf8", {60: two}, one([pg.insert(60, ["1", b"\xff\xfe"])]), (2, 1)))
    out.append(("key_shape", {60: two}, one([pg.delete(60, key=["1", "a", "b"])]), (2, 15)))
    out.append(("key_missing_value", {60: two}, one([pg.delete(60, key=[UNCHANGED])]), (2, 16)))
    noid = pg.relation(60, "public", "t", "n", rel_cols(two, set()))
    out.append(("key_no_columns", {60: two}, one([pg.delete(60, key=["1"])], rels=(noid,)), (2, 14)))
    out.append(("old_full_count", {60: two}, one([pg.delete(60, old=["1"])]), (2, 12)))
    out.append(("missing_table_state", {60: two}, one([pg.insert(61, ["1", "a"])]), (2, 19)))
    out.append(("missing_table_schema", {}, one([]), (1, 23)))
    badrel = pg.relation(60, "public", "t", "d", [(1, "id", INT8, -1), (0, "ghost", TEXT, -1)])
    out.append(("unknown_columns", {60: two}, one([], rels=(badrel,)), (1, 22)))
    # state machine
    w = pg.StreamWriter()
    w.emit(pg.insert(60, ["1", "a"]))
    out.append(("dml_outside_tx", {60: two}, w, (0, 17)))
    w = pg.StreamWriter()
    w.emit(pg.commit(0, 5, 6, 7))
    out.append(("commit_outside_tx", {60: two}, w, (0, 17)))
    w = pg.StreamWriter()
    w.emit(pg.begin(100, 1, 2))
    w.emit(pg.commit(0, 101, 108, 7))
    out.append(("commit_lsn_mismatch", {60: two}, w, (1, 18)))
    w = pg.StreamWriter()
    w.emit(pg.relation(60, "public", "t", "d", rel_cols(two, {"id"})))
    out.append(("relation_outside_tx", {60: two}, w, (0, 17)))
    w = pg.StreamWriter()
    w.emit(pg.message(1, 5, "supabase_etl_ddl", b"{}"))
    out.append(("ddl_message_outside_tx", {60: two}, w, (0, 17)))
    w = pg.StreamWriter()
    w.emit(pg.begin(100, 1, 2))
    w.emit(pg.truncate([99]))
    out.append(("truncate_unknown_table", {60: two}, w, (1, 19)))
    # malformed frames
    w = pg.StreamWriter()
    w.emit(pg.begin(100, 1, 2))
    w.emit(b"Z123")
    out.append(("unknown_tag", {60: two}, w, (1, 24)))
    w = pg.StreamWriter()
    w.emit(pg.begin(100, 1, 2))
    w.emit(pg.insert(60, ["1", "a"])[:-1])            # text length runs past the frame
    out.append(("truncated_tuple", {60: two}, w, (1, 24)))
    w = pg.StreamWriter()
    w.emit(pg.begin(100, 1, 2))
    w.emit(b"I" + (60).to_bytes(4, "big") + b"X" + pg.encode_tuple(["1"]))
    out.append(("bad_tuple_tag", {60: two}, w, (1, 24)))
    return out
