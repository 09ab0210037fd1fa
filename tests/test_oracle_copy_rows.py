"""Oracle restatement of the COPY-text row parser (SURVEY §8f N1 groundwork) against the reference's own
tests: crates/etl/src/conversions/table_row.rs:206-533.  CPU only; no device path binds to it yet."""
import os
import struct
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
from canon import decode_cell  # noqa: E402

INT4, TEXT, BOOL, FLOAT8 = 23, 25, 16, 701
F64_3_15 = struct.unpack("<Q", struct.pack("<d", 3.15))[0]
E_UTF8, E_PARSE_INT, E_BOOL = 1, 2, 8


def rows(oracle_mod, oids, data):
    e, ecol, cells, text, heap = oracle_mod.parse_copy_row(oids, data)
    vals = [decode_cell(t, v, a, text, heap) for t, v, a in cells]
    return e, ecol, vals


def S(x):
    return x


def test_simple_null_empty_single_and_mixed(oracle_mod):
    three = [INT4, TEXT, BOOL]
    assert rows(oracle_mod, three, b"123\tJohn Doe\tt\n") == (0, None, [123, S("John Doe"), True])     # :206-218
    assert rows(oracle_mod, three, b"456\t\\N\tf\n") == (0, None, [456, None, False])              # :220-232
    assert rows(oracle_mod, three, b"0\t\tf\n") == (0, None, [0, S(""), False])                         # :234-246
    assert rows(oracle_mod, [INT4], b"42\n") == (0, None, [42])                                                   # :248-258
    e, _, vals = rows(oracle_mod, [INT4, FLOAT8, TEXT, BOOL], b"123\t3.15\tHello World\tt\n")                               # :260-279
    assert e == 0 and vals == [123, ("f64", F64_3_15), S("Hello World"), True]
    assert rows(oracle_mod, three, b"123\t John Doe \tt\n") == (0, None, [123, S(" John Doe "), True])  # :370-382


def test_row_level_errors(oracle_mod):
    three = [INT4, TEXT, BOOL]
    assert rows(oracle_mod, [INT4], b"42")[0] == oracle_mod.E_COPY_NOT_TERMINATED                                          # :281-292
    e, ecol, vals = rows(oracle_mod, three, b"123\tJohn\n")                                                                # :294-304 "row contains 2 columns"
    assert e == oracle_mod.E_COPY_COLUMN_COUNT and ecol == 2 and len(vals) == 2
    e, ecol, _ = rows(oracle_mod, three, b"123\tJohn\tt\textra\n")                                                         # :306-316 "at least 4 columns"
    assert e == oracle_mod.E_COPY_COLUMN_COUNT and ecol == 3
    assert rows(oracle_mod, [TEXT], b"Hello\xff\xfe\n")[0] == E_UTF8                                                       # :318-326
    assert rows(oracle_mod, [INT4], b"not_a_number\n")[:2] == (E_PARSE_INT, 0)                                             # :328-336
    assert rows(oracle_mod, three, b"\t\t\n")[:2] == (E_PARSE_INT, 0)                                                      # :410-418 empty int field


def test_escapes_and_null_marker(oracle_mod):
    assert rows(oracle_mod, [TEXT], b"Text\\\\\n") == (0, None, [S("Text\\")])                                             # :338-348
    assert rows(oracle_mod, [TEXT], b"\\N\n") == (0, None, [None])                                                    # :350-368
    assert rows(oracle_mod, [TEXT], b"\\\\N\n") == (0, None, [None])
    assert rows(oracle_mod, [TEXT], b"\\\\A\n") == (0, None, [S("\\A")])
    assert rows(oracle_mod, [TEXT, TEXT], b"value\\twith\\ttabs\tnormal\\tvalue\n") == (0, None, [S("value\twith\ttabs"), S("normal\tvalue")])   # :420-434
    assert rows(oracle_mod, [TEXT] * 3, b"\\tstart\tmiddle\\nvalue\tend\\r\n") == (0, None, [S("\tstart"), S("middle\nvalue"), S("end\r")])       # :436-452
    assert rows(oracle_mod, [TEXT], "Hello\\t🌍\\nWorld\\r测试\n".encode()) == (0, None, [S("Hello\t🌍\nWorld\r测试")])                              # :454-468
    for raw, want in [(b"\\b\n", "\x08"), (b"\\f\n", "\x0c"), (b"\\n\n", "\n"), (b"\\r\n", "\r"), (b"\\t\n", "\t"), (b"\\v\n", "\x0b"),
                      (b"\\\\\n", "\\"), (b"\\x\n", "x"), (b"\\1\n", "1"), (b"\\!\n", "!"), (b"\\@\n", "@"), (b'\\"\n', '"')]:                     # :470-510
        assert rows(oracle_mod, [TEXT], raw) == (0, None, [S(want)]), raw
    assert rows(oracle_mod, [TEXT], b"\n") == (0, None, [S("")])                                                            # :512-533


def test_large_row_and_trailing_bytes(oracle_mod):
    data = "\t".join(str(i) for i in range(50)).encode() + b"\n"                                                            # :384-408
    assert rows(oracle_mod, [INT4] * 50, data) == (0, None, [i for i in range(50)])
    # whatever follows the last LF without a terminator of its own is dropped when the input ends (:88-96)
    assert rows(oracle_mod, [INT4], b"42\nxyz") == (0, None, [42])
    # ... but a complete second line is more fields of the same row
    assert rows(oracle_mod, [INT4], b"42\n43\n")[0] == oracle_mod.E_COPY_COLUMN_COUNT
