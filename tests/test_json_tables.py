"""The table-driven JSON acceptor used on the device (tools/gen_json_tables.py) against the oracle's
serde_json restatement: same verdict on hand-written edge cases and on a seeded fuzz corpus."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import gen_json_tables as jt  # noqa: E402

EDGE = [b"", b" ", b"0", b"-0", b"-", b"01", b"1.", b"1.5", b"1e5", b"1E+5", b"1e", b"1e+", b"-1.25e-3 ", b"true", b"tru", b"truee",
        b"false", b"null", b"nul", b'""', b'"a"', b'"a', b'"\\n"', b'"\\x"', b'"\\u0041"', b'"\\u004"', b'"\\ud800"', b'"\\ud800\\udc00"',
        b'"\\ud800\\n"', b'"\\udc00"', b'"\\ud800\\ud800"', b'"tab\there"', b'"nl\n"', b"[]", b"[ ]", b"[1]", b"[1,]", b"[,1]", b"[1 2]",
        b"{}", b"{ }", b'{"a":1}', b'{"a":1,}', b'{"a" 1}', b'{"a":}', b"{1:2}", b'{"a":1}}', b'{"a":[1,{"b":null}]}', b"[1]]", b"[}",
        b'{"a":1]', b"1 2", b"1,", b" 1 ", b"\t[\n1\r,2 ] ", b'"\xc3\xa9"', b"\xc3\xa9", b"[" * 127 + b"]" * 127, b"[" * 128 + b"]" * 128,
        b"[" * 126 + b"{}" + b"]" * 126, b"[" * 127 + b"{}" + b"]" * 127, b'{"a":' * 60 + b"1" + b"}" * 60,
        b"[" * 70 + b'{"k":[' * 20 + b"]}" * 20 + b"]" * 70, b"-e", b"+1", b".5", b"0.0e-0", b"0e", b"00", b"-01", b"1.e5", b"t", b"nulll",
        b'{"a":true,"b":false,"c":null,"d":[1,2.5,-3e2,"x"]}', b'["a",]', b'{"a":1 "b":2}', b'{"a":1,,"b":2}', b'[1,,2]', b'"\\u00e9\\\\"',
        b'"\\/"', b'"\\b\\f\\r\\t"', b'"\\uD83D\\uDE00"', b'"\\uD83D\\u0041"', b'"\\uD83Dx"', b'{"\\u0041":1}', b'{"a\\"b":1}']


def test_tables_match_oracle(oracle_mod):
    def oracle_ok(b):
        return oracle_mod.parse_cell(114, b)[0] == 0
    for b in EDGE:
        assert jt.simulate(b) == oracle_ok(b), b
    rng = random.Random(7)
    toks = [b"{", b"}", b"[", b"]", b",", b":", b'"', b"\\", b" ", b"\n", b"0", b"1", b"9", b"-", b"+", b".", b"e", b"E", b"true", b"false",
            b"null", b'"k"', b'"v\\n"', b'"\\u12aB"', b'"\\ud83d\\ude00"', b"12.5e-3", b"a", b"u", b"\x01", b"\xc3\xa9", b'"x y"', b"t", b"n", b"f"]
    good = [b for b in EDGE if oracle_ok(b) and b]
    n_ok = 0
    for i in range(30000):
        if i % 3 == 0:
            b = b"".join(rng.choice(toks) for _ in range(rng.randint(1, 12)))
        elif i % 3 == 1:   # mutate a valid document
            g = bytearray(rng.choice(good))
            for _ in range(rng.randint(1, 2)):
                op = rng.randint(0, 2)
                p = rng.randrange(len(g)) if g else 0
                if op == 0 and g:
                    del g[p]
                elif op == 1:
                    g[p:p] = rng.choice(toks)
                elif g:
                    g[p] = rng.choice(b'{}[],:"\\ 0123456789-+.eEtrufalsn\x01\t')
            b = bytes(g)
        else:              # generated valid documents
            def gen(d):
                k = rng.randint(0, 7 if d < 4 else 4)
                if k == 0:
                    return str(rng.randint(-10**6, 10**6)).encode()
                if k == 1:
                    return rng.choice([b"true", b"false", b"null"])
                if k == 2:
                    return ("%g" % (rng.random() * 10 ** rng.randint(-8, 8))).encode()
                if k in (3, 4):
                    return b'"' + rng.choice([b"abc", b"x\\ny", b"\\u00e9", b"\\ud83d\\ude00", b"", b"sp ace"]) + b'"'
                if k in (5, 6):
                    return b"[" + rng.choice([b",", b" , "]).join(gen(d + 1) for _ in range(rng.randint(0, 3))) + b"]"
                return b"{" + b",".join(b'"k%d": ' % j + gen(d + 1) for j in range(rng.randint(0, 3))) + b"}"
            b = gen(0)
        try:
            b.decode("utf-8")          # cells are UTF-8 validated before the JSON grammar (event.rs:972)
        except UnicodeDecodeError:
            continue
        want = oracle_ok(b)
        n_ok += want
        assert jt.simulate(b) == want, b
    assert 4000 < n_ok < 25000, n_ok   # the corpus exercises both verdicts


def test_generated_header_is_current():
    """etl_b200/csrc/json_tables.cuh is what the generator emits."""
    import tempfile
    path = os.path.join(os.path.dirname(__file__), "..", "etl_b200", "csrc", "json_tables.cuh")
    with tempfile.NamedTemporaryFile("r", suffix=".cuh") as t:
        jt.emit(t.name)
        assert open(t.name).read() == open(path).read()
