"""Committed fixtures (tests/golden/workload_digests.json, made by tools/make_golden.py): the generator, the
oracle and the GPU path must all reproduce the same bytes / events for small instances of C1..C5."""
import hashlib
import json
import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "workload_digests.json")))


def _stream(name):
    from etl_b200 import workloads as wl
    g = GOLD[name]
    w = wl.make(name, g["scale"], n_segments=g["n_segments"])
    stream, _ = w.generate()
    return w, stream.tobytes()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_golden(oracle_mod, name):
    from make_golden import digest_events
    w, stream = _stream(name)
    g = GOLD[name]
    assert len(stream) == g["stream_bytes"] and hashlib.sha256(stream).hexdigest() == g["stream_sha256"], "generator drifted"
    o = oracle_mod.Oracle()
    for tid, cols in w.table_schemas().items():
        o.put_table_schema(tid, cols)
    p = o.decode(stream)
    assert p.first_error[0] is None and p.n_records == g["n_records"]
    assert digest_events(p, stream) == g["events_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_reproduces_golden(name):
    from etl_b200 import decoder
    from make_golden import digest_events
    w, stream = _stream(name)
    g = GOLD[name]
    dec = decoder.Decoder(0)
    for tid, cols in w.table_schemas().items():
        dec.put_table_schema(tid, cols)
    p = dec.decode(stream)
    dec.close()
    assert p.first_error[0] is None and p.n_records == g["n_records"]
    assert digest_events(p, stream) == g["events_sha256"]
