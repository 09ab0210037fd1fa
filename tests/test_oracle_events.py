"""Oracle vs the reference's event-level known answers (event.rs:1301-1637) and the apply-loop
state machine (apply.rs:600-626, 1927-2006).  Pure CPU."""
import pytest

import scenarios as sc
from canon import planes_to_events
from etl_b200 import pgoutput as pg
from oracle import pyoracle as po


def run_oracle(tables, stream: bytes, carry=None):
    o = po.Oracle()
    for tid, cols in tables.items():
        o.put_table_schema(tid, cols)
    return o.decode(stream, carry)


def dml(events):
    return [{k: e[k] for k in ("kind", "row", "old") if k in e} for e in events if e["kind"] in ("insert", "update", "delete")]


@pytest.mark.parametrize("name,tables,w,expected", sc.reference_update_delete_scenarios() + sc.tuple_level_scenarios(),
                         ids=lambda x: x if isinstance(x, str) else None)
def test_reference_event_vectors(name, tables, w, expected):
    stream = w.bytes()
    p = run_oracle(tables, stream)
    if isinstance(expected, tuple) and expected[0] == "error":
        assert p.first_error[0] is not None and p.first_error[2] == expected[1]
        return
    assert p.first_error[0] is None, p.first_error
    assert dml(planes_to_events(p, stream)) == expected


@pytest.mark.parametrize("name,tables,w,expected", sc.error_scenarios(), ids=lambda x: x if isinstance(x, str) else None)
def test_error_scenarios(name, tables, w, expected):
    stream = w.bytes()
    p = run_oracle(tables, stream)
    assert (p.first_error[0], p.first_error[2]) == expected, p.first_error
    assert p.n_records == expected[0]


def test_state_machine_ordinals_and_lsns():
    """apply.rs:1927-2006: Begin resets the ordinal to 0; R/I/U/D/T/C each consume one; M/O/Y/keepalive none."""
    two = [sc.col("id", sc.INT8, 1), sc.col("v", sc.TEXT, None, True)]
    w = pg.StreamWriter()
    w.emit_keepalive()
    w.emit(pg.begin(0xAAAA, 11, 5))
    w.emit(pg.origin(1, "o"))
    w.emit(pg.relation(60, "public", "t", "d", sc.rel_cols(two, {"id"})))
    w.emit(pg.type_msg(99, "public", "ty"))
    w.emit(pg.insert(60, ["1", "a"]))
    w.emit(pg.message(1, 9, "other_prefix", b"zz"))
    w.emit(pg.message(1, 9, "supabase_etl_ddl", b"{}"))
    w.emit(pg.update(60, ["1", None]))
    w.emit(pg.truncate([60], 3))
    w.emit(pg.delete(60, key=["1", None]))
    w.emit(pg.commit(1, 0xAAAA, 0xAAB0, 12))
    w.emit(pg.begin(0xBBBB, 13, 6))
    w.emit(pg.insert(60, ["2", None]))
    stream = w.bytes()
    p = run_oracle({60: two}, stream)
    assert p.first_error[0] is None
    ev = planes_to_events(p, stream)
    assert [(e["kind"], e["tx_ordinal"], e["commit_lsn"]) for e in ev] == [
        ("begin", 0, 0xAAAA), ("relation", 1, 0xAAAA), ("insert", 2, 0xAAAA), ("update", 3, 0xAAAA),
        ("truncate", 4, 0xAAAA), ("delete", 5, 0xAAAA), ("commit", 6, 0xAAAA), ("begin", 0, 0xBBBB),
        ("insert", 1, 0xBBBB)]
    assert ev[0]["timestamp"] == 11 and ev[0]["xid"] == 5
    assert ev[6]["flags"] == 1 and ev[6]["end_lsn"] == 0xAAB0 and ev[6]["timestamp"] == 12
    assert ev[4]["options"] == 3 and ev[4]["rel_ids"] == [60]
    assert p.carry_out == (1, 0xBBBB, 2)
    assert p.n_records == 14 and p.n_events == 9
    assert [chr(k) for k in p.rec_kind] == list("kBORYIMMUTDCBI")
    # metrics (event.rs:260-270): insert "1"+"a" and "2"; update "1"; delete key "1"
    assert (p.insert_bytes, p.update_bytes, p.delete_bytes) == (3, 1, 1)


def test_carry_in_state_and_relation_cache_across_batches():
    two = [sc.col("id", sc.INT8, 1), sc.col("v", sc.TEXT, None, True)]
    w1 = pg.StreamWriter()
    w1.emit(pg.begin(0x500, 1, 2))
    w1.emit(pg.relation(60, "public", "t", "d", sc.rel_cols(two, {"id"})))
    w1.emit(pg.insert(60, ["1", "a"]))
    w2 = pg.StreamWriter()
    w2.emit(pg.insert(60, ["2", "b"]))
    w2.emit(pg.commit(0, 0x500, 0x508, 3))
    o = po.Oracle()
    o.put_table_schema(60, two)
    p1 = o.decode(w1.bytes())
    assert p1.carry_out == (1, 0x500, 3) and len(p1.schemas) == 1
    p2 = o.decode(w2.bytes(), p1.carry_out)
    assert p2.first_error[0] is None
    ev = planes_to_events(p2, w2.bytes())
    assert [(e["kind"], e["tx_ordinal"], e["commit_lsn"]) for e in ev] == [("insert", 3, 0x500), ("commit", 4, 0x500)]
    assert len(p2.schemas) == 1 and p2.schemas[0].effective_off == 0 and int(p2.rec_schema[0]) == 0
    assert p2.carry_out[0] == 0


def test_relation_masks_schema_rs_406_438():
    """Column filtering: a stored column absent from the Relation message is not replicated;
    identity = flag bit0, or every relation column under REPLICA IDENTITY FULL (event.rs:351-366)."""
    cols = [sc.col("id", sc.INT8, 1), sc.col("secret", sc.TEXT), sc.col("v", sc.INT4, None, True)]
    w = pg.StreamWriter()
    w.emit(pg.begin(0x10, 1, 2))
    w.emit(pg.relation(70, "public", "t", "d", [(1, "id", sc.INT8, -1), (0, "v", sc.INT4, -1)]))
    w.emit(pg.insert(70, ["5", "6"]))
    w.emit(pg.relation(70, "public", "t", "f", [(0, "id", sc.INT8, -1), (0, "v", sc.INT4, -1)]))
    w.emit(pg.update(70, ["5", "7"], old=["5", "6"]))
    w.emit(pg.commit(0, 0x10, 0x18, 3))
    stream = w.bytes()
    p = run_oracle({70: cols}, stream)
    assert p.first_error[0] is None
    s0, s1 = p.schemas
    assert (s0.n_cols, s0.n_identity, list(s0.col_index), list(s0.col_flags)) == (2, 1, [0, 2], [2, 1])
    assert (s1.n_cols, s1.n_identity, list(s1.col_flags)) == (2, 2, [2, 3])
    assert s1.effective_off == w.relation_offsets[1]
    ev = planes_to_events(p, stream)
    assert ev[2]["row"] == [5, 6] and ev[2]["schema"] == 0
    assert ev[4]["row"] == ("full", [5, 7]) and ev[4]["old"] == ("full", [5, 6]) and ev[4]["schema"] == 1
