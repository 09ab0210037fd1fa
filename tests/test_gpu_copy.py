"""COPY-text rows on the device (etl_dec_copy_decode: k_copy_rows + k_heavy) against the oracle's restatement of
parse_table_row_from_postgres_copy_bytes, on the reference's own test vectors (table_row.rs:206-533) and on bulk
synthetic tables (canonical digests, strings by content)."""
import struct

import numpy as np
import pytest

from canon import decode_cell

pytestmark = pytest.mark.gpu

INT4, TEXT, BOOL, FLOAT8, NUMERIC, JSONB, TSTZ, UUID, BYTEA, INT8, DATE = 23, 25, 16, 701, 1700, 3802, 1184, 2950, 17, 20, 1082
IN_HEAP = 1 << 63
E_NOT_TERMINATED, E_COLUMN_COUNT = 25, 26            # etl_error_code; the oracle's row function says 101 / 102
ORC2ETL = {101: E_NOT_TERMINATED, 102: E_COLUMN_COUNT}


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from etl_b200 import decoder
    return decoder


def cols_of(oids):
    return [dict(name=f"c{i}", type_oid=o, pk=1 if i == 0 else None, nullable=True) for i, o in enumerate(oids)]


def gpu_rows(gpu, oids, rows):
    dec = gpu.Decoder(0)
    dec.put_table_schema(7, cols_of(oids))
    b = dec.copy_decode(7, rows)
    dec.close()
    return b


def values_of(b, r):
    heap, stream = b.heap.tobytes(), b.stream.tobytes()
    out = []
    for c in range(b.n_cols):
        i = r * b.n_cols + c
        t, v, a = int(b.cell_tag[i]), int(b.cell_val[i]), int(b.cell_aux[i])
        if t in (2, 15) and v & IN_HEAP:
            out.append(decode_cell(t, v & ~IN_HEAP, a, heap, heap))
        else:
            out.append(decode_cell(t, v, a, stream, heap))
    return out


REF_ROWS = [   # (type oids, row bytes) — crates/etl/src/conversions/table_row.rs:206-533
    ([INT4, TEXT, BOOL], b"123\tJohn Doe\tt\n"), ([INT4, TEXT, BOOL], b"456\t\\N\tf\n"), ([INT4, TEXT, BOOL], b"0\t\tf\n"), ([INT4], b"42\n"),
    ([INT4, FLOAT8, TEXT, BOOL], b"123\t3.15\tHello World\tt\n"), ([INT4, TEXT, BOOL], b"123\t John Doe \tt\n"),
    ([INT4], b"42"), ([INT4, TEXT, BOOL], b"123\tJohn\n"), ([INT4, TEXT, BOOL], b"123\tJohn\tt\textra\n"), ([TEXT], b"Hello\xff\xfe\n"),
    ([INT4], b"not_a_number\n"), ([INT4, TEXT, BOOL], b"\t\t\n"), ([TEXT], b"Text\\\\\n"), ([TEXT], b"\\N\n"), ([TEXT], b"\\\\N\n"), ([TEXT], b"\\\\A\n"),
    ([TEXT, TEXT], b"value\\twith\\ttabs\tnormal\\tvalue\n"), ([TEXT] * 3, b"\\tstart\tmiddle\\nvalue\tend\\r\n"),
    ([TEXT], "Hello\\t🌍\\nWorld\\r测试\n".encode()), ([TEXT], b"\\b\n"), ([TEXT], b"\\f\n"), ([TEXT], b"\\v\n"), ([TEXT], b"\\x\n"), ([TEXT], b'\\"\n'),
    ([TEXT], b"\n"), ([INT4] * 50, "\t".join(str(i) for i in range(50)).encode() + b"\n"), ([INT4], b"42\nxyz"), ([INT4], b"42\n43\n"),
    ([TEXT], b"tail\\"), ([TEXT], b"ok\n\\"), ([INT4, TEXT], b"1\t\xc3\n"), ([INT4, TEXT], b"x\t\xc3\n"), ([NUMERIC, JSONB], b"12.50\t{\"a\": [1, 2]}\n"),
    ([NUMERIC, JSONB], b"NaN\t{\"a\": \\\\\"x\\\\\"}\n"), ([BYTEA, UUID], b"\\\\x00ff10\t550e8400-e29b-41d4-a716-446655440000\n"),
    ([TSTZ, DATE, INT8], b"2024-03-01 12:34:56.123456+00\t2024-02-29\t-9223372036854775808\n"), ([TSTZ], b"2024-03-01 12:34:56+05:30\n"),
]


@pytest.mark.parametrize("idx", range(len(REF_ROWS)))
def test_reference_rows(gpu, oracle_mod, idx):
    oids, row = REF_ROWS[idx]
    e, ecol, cells, text, heap = oracle_mod.parse_copy_row(oids, row)
    b = gpu_rows(gpu, oids, [row])
    if e:
        step = 0 if e == 1 and not _field_error(oracle_mod, oids, row) else 1 + (ecol if ecol is not None else 0)
        assert b.first_error[0] == 0 and b.first_error[2] == ORC2ETL.get(e, e), (b.first_error, e, ecol)
        if e != 1:
            assert b.first_error[1] == step
    else:
        assert b.first_error[0] is None, b.first_error
        want = [decode_cell(t, v, a, text, heap) for t, v, a in cells]
        assert values_of(b, 0) == want


def _field_error(oracle_mod, oids, row):
    return False


def synth_rows(n, seed, bad_at=None):
    """rows of (int4, text, bool, numeric, jsonb, timestamptz, uuid, bytea, float8, int8) with NULLs, escapes and non-ASCII text"""
    import uuid as _uuid
    rng = np.random.default_rng(seed)
    oids = [INT4, TEXT, BOOL, NUMERIC, JSONB, TSTZ, UUID, BYTEA, FLOAT8, INT8]
    words = ["alpha", "beta", "gamma delta", "tab\\there", "line\\nbreak", "back\\\\slash", "caf\u00e9", "\u6d4b\u8bd5", "", "x" * 300, "\\N-not-null", "q\\\\N"]
    out = []
    for i in range(n):
        f = [str(int(rng.integers(-2**31, 2**31))), words[int(rng.integers(0, len(words)))], "t" if rng.integers(0, 2) else "f",
             ("%d.%02d" % (int(rng.integers(0, 10**9)), int(rng.integers(0, 100)))) if rng.integers(0, 50) else "NaN",
             '{"k": %d, "s": "v%d", "a": [true, null, 1.5e3]}' % (int(rng.integers(0, 1000)), i % 97),
             "2024-%02d-%02d %02d:%02d:%02d.%06d+00" % (int(rng.integers(1, 13)), int(rng.integers(1, 29)), int(rng.integers(0, 24)), int(rng.integers(0, 60)),
                                                         int(rng.integers(0, 60)), int(rng.integers(0, 10**6))),
             str(_uuid.UUID(int=int(rng.integers(0, 2**63)) << 64 | int(rng.integers(0, 2**63)))), "\\\\x" + bytes(rng.integers(0, 256, size=int(rng.integers(0, 12)), dtype=np.uint8)).hex(),
             repr(float(rng.standard_normal() * 10.0 ** int(rng.integers(-5, 6)))), str(int(rng.integers(-2**62, 2**62)))]
        for c in range(1, len(f)):
            if rng.integers(0, 20) == 0:
                f[c] = "\\N"
        if bad_at is not None and i == bad_at:
            f[0] = "12x"
        out.append(("\t".join(f) + "\n").encode())
    return oids, out


@pytest.mark.parametrize("n,bad_at", [(1, None), (33, None), (5000, None), (5000, 3777), (200000, None)])
def test_bulk_rows_match_oracle(gpu, oracle_mod, n, bad_at):
    oids, rows = synth_rows(n, 1234 + n, bad_at)
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in rows])
    buf = np.frombuffer(b"".join(rows), dtype=np.uint8)
    want, err = oracle_mod.copy_rows_digest(oids, buf, offs)
    dec = gpu.Decoder(0)
    dec.put_table_schema(7, cols_of(oids))
    b = dec.copy_decode(7, buf, offs)
    dec.close()
    if err is None:
        assert b.first_error[0] is None, b.first_error
        n_ok = n
    else:
        assert (b.first_error[0], b.first_error[1], b.first_error[2]) == (err[0], 1 + err[1], ORC2ETL.get(err[2], err[2]))
        n_ok = err[0]
    got = oracle_mod.copy_planes_digest(b.cell_tag, b.cell_val, b.cell_aux, n_ok, len(oids), b.stream, b.heap)
    assert got == want
