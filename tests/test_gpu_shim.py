"""The shim stand-in (etl_shim_materialise, csrc/shim_materialise.cpp): planes → owned Vec<Event>-shaped rows with
Event::size_hint per event.  Checked against a Python restatement of the reference's formulas
(types/event.rs:288-312, types/table_row.rs:250-345, conversions/event.rs:601-671) evaluated on the ORACLE's planes,
and against Python's json module for the Json trees."""
import ctypes as C
import json

import numpy as np
import pytest

from etl_b200 import abi, workloads as wl

pytestmark = pytest.mark.gpu

LAYOUT = dict(cell=32, table_row=32, partial=72, begin=40, commit=48, insert=96, update=176, delete=104, truncate=56,
              rts=64, relation=40, json_value=32, usize=8)      # = kDefaultLayout (the Rust sizes are parameters of the shim)


class Num(str):
    """a JSON number kept as text (serde_json arbitrary_precision)"""


def vec_cap(n):
    c = 4
    if n == 0:
        return 0
    while c < n:
        c <<= 1
    return c


def json_bytes(v):
    if isinstance(v, Num):                               # Value::Number: no heap in the estimate (table_row.rs:333)
        return 0
    if isinstance(v, str):
        return len(v.encode())
    if isinstance(v, list):
        return vec_cap(len(v)) * LAYOUT["json_value"] + sum(json_bytes(x) for x in v)
    if isinstance(v, dict):
        return sum(len(k.encode()) + json_bytes(x) for k, x in v.items())
    return 0


def cell_bytes(tag, val, aux, stream, heap, cloned):
    if tag in (2, 16):
        return aux
    if tag == 9:
        kind = heap[val]
        pushed = int.from_bytes(heap[val + 6:val + 8], "little")
        return 0 if kind else 2 * (aux if cloned else vec_cap(pushed))
    if tag == 15:
        return json_bytes(json.loads(bytes(stream[val:val + aux]), parse_int=Num, parse_float=Num))
    return 0


def partial_cap(n_cols, first_missing, n_present):
    cap, ln = max(4, n_cols - first_missing), 0
    def need(w):
        nonlocal cap
        if w > cap:
            cap = max(cap * 2, w, 4)
    need(first_missing)
    ln = first_missing
    while ln < n_present:
        need(ln + 1)
        ln += 1
    return cap


def expected_hints(p, stream):
    heap = p.heap.tobytes()
    out = []
    for r in range(p.n_records):
        fl = int(p.rec_flags[r])
        if not fl & 0x80:
            continue
        kind = chr(int(p.rec_kind[r]))
        a, b = int(p.rec_cell_base[r]), int(p.rec_cell_base[r + 1])
        if kind == "B":
            out.append(LAYOUT["begin"]); continue
        if kind == "C":
            out.append(LAYOUT["commit"]); continue
        if kind == "R":
            out.append(LAYOUT["relation"]); continue
        if kind == "T":
            out.append(LAYOUT["truncate"] + (b - a - 1) * LAYOUT["rts"]); continue
        sc = p.schemas[int(p.rec_schema[r])]
        n_old = sc.n_cols if fl & 1 else (sc.n_identity if fl & 2 else 0)
        cells = [(int(p.cell_tag[i]), int(p.cell_val[i]), int(p.cell_aux[i])) for i in range(a, b)]
        old, new = cells[:n_old], cells[n_old:]
        h = 0
        if n_old:
            h += LAYOUT["table_row"] + n_old * LAYOUT["cell"] + sum(cell_bytes(*c, stream, heap, False) for c in old)
        if kind in "IU":
            present = [c for c in new if c[0] != 254]
            body = sum(cell_bytes(*c, stream, heap, kind == "U" and c[0] == 9 and c in old) for c in present)
            if len(present) != len(new):
                first = next(i for i, c in enumerate(new) if c[0] == 254)
                h += LAYOUT["partial"] + LAYOUT["table_row"] + partial_cap(sc.n_cols, first, len(present)) * LAYOUT["cell"] + body \
                    + vec_cap(len(new) - len(present)) * LAYOUT["usize"]
            else:
                h += LAYOUT["table_row"] + sc.n_cols * LAYOUT["cell"] + body
        h += {"I": LAYOUT["insert"], "U": LAYOUT["update"], "D": LAYOUT["delete"]}[kind]
        out.append(h)
    return out


@pytest.mark.parametrize("name,scale", [("c2", 0.01), ("c3", 0.002), ("c5", 0.002), ("c4", 0.001)])
def test_materialised_events_size_hints(oracle_mod, name, scale):
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
    want = expected_hints(orc.decode(raw), raw)
    st = decoder.Stager(stream.nbytes, 2048)
    st.append_framed(stream)
    lib = abi.load()
    with dec.decode_input(st.view(), to_host=True) as bh:
        lst = C.c_void_p()
        assert lib.etl_shim_materialise(bh._h, st.view().host_buf, None, C.byref(lst)) == 0
        n = lib.etl_shim_event_count(lst)
        assert n == len(want) == bh.summary().n_events
        got = [lib.etl_shim_size_hint(lst, i) for i in range(n)]
        assert got == want
        assert lib.etl_shim_total_size_hint(lst) == sum(want)
        lib.etl_shim_event_list_free(lst)
    st.close()
    dec.close()


def test_materialised_json_trees_match_python_json(oracle_mod):
    """serde_json semantics of the tree builder: keys sorted, duplicate keys keep the last value, escapes decoded,
    numbers verbatim (arbitrary_precision, text.rs:625-636)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import scenarios as sc
    from etl_b200 import decoder, pgoutput as pg
    docs = ['{"b":1,"a":[1,2,{"z":null}],"b":2}', '  [1e309, -0.0, 12345678901234567890123, "\\u00e9\\ud83e\\udd14\\n"] ', '"x"', "true",
            '{"k":{"k":{"k":[[],{}]}},"":""}', '{"esc":"a\\\\b\\"c\\/d"}']
    cols = [sc.col("id", sc.INT8, 1), sc.col("j", 3802, None, True)]
    rel = pg.relation(77, "public", "j", "d", sc.rel_cols(cols, {"id"}))
    w = sc.stream_with({77: cols}, [rel], [pg.insert(77, [str(i), d]) for i, d in enumerate(docs)])
    raw = w.bytes()
    dec = decoder.Decoder(0)
    dec.put_table_schema(77, cols)
    st = decoder.Stager(len(raw), 2048)
    st.append_framed(raw)
    lib = abi.load()
    with dec.decode_input(st.view(), to_host=True) as bh:
        assert bh.summary().first_error.record_index == 2**64 - 1
        lst = C.c_void_p()
        assert lib.etl_shim_materialise(bh._h, st.view().host_buf, None, C.byref(lst)) == 0
        ev = 2                                               # Begin, Relation, then the inserts
        for d in docs:
            buf = C.create_string_buffer(4096)
            n = lib.etl_shim_json_text(lst, ev, 1, buf, 4096)
            assert n >= 0
            got = json.loads(buf.value.decode(), parse_int=Num, parse_float=Num)
            want = json.loads(d, parse_int=Num, parse_float=Num)
            assert got == want and list(got) == sorted(got) if isinstance(got, dict) else got == want
            ev += 1
        lib.etl_shim_event_list_free(lst)
    st.close()
    dec.close()
