"""The DEVICE cell parsers, compiled for the host with one-lane stand-ins for the warp intrinsics
(tests/emul/host_parsers.cpp — test infrastructure, not a product path), fuzzed against the oracle at a volume the
GPU tests cannot afford: 100 000 spellings per decode class (ETL_HOST_FUZZ_N; 400 000 each was run once: 6.4 M, all equal), each through the exact path and through the fast
path k_rows / k_heavy take.  Same verdict, same error code, same typed value."""
import ctypes as C
import os
import random
import struct
import subprocess
import sys
import uuid

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
from canon import decode_cell  # noqa: E402
from test_gpu_parity import FUZZ_KINDS  # noqa: E402  (seed spellings per decode class)

N_PER_KIND = int(os.environ.get("ETL_HOST_FUZZ_N", "100000"))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emul", "host_parsers.cpp")
    so = os.path.join(HERE, "emul", "libhost_parsers.so")
    deps = [src] + [os.path.join(ROOT, "etl_b200", "csrc", f) for f in ("cell_parsers.cuh", "float_parse.cuh", "float_tables.cuh", "json_tables.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "etl_b200", "csrc"), "-o", so, src])
    L = C.CDLL(so)
    L.emu_parse_cell.restype = C.c_uint32
    L.emu_parse_cell.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.c_uint32, C.POINTER(C.c_uint32)]
    cap = 1 << 16
    heap = (C.c_uint8 * cap)()
    tag, val, aux, hl = C.c_uint8(), C.c_uint64(), C.c_uint32(), C.c_uint32()

    def parse(kind, text, fast):
        e = L.emu_parse_cell(kind, text, len(text), fast, C.byref(tag), C.byref(val), C.byref(aux), heap, cap, C.byref(hl))
        if e:
            return e, None
        return 0, decode_cell(tag.value, val.value, aux.value, text, bytes(heap))
    return parse


def _generators(oid, rng):
    """extra value generators per class, on top of the mutations of the seed spellings"""
    def digits(n):
        return "".join(rng.choice("0123456789") for _ in range(n))
    if oid in (21, 23, 20, 26):
        return [lambda: rng.choice(["", "-", "+"]) + digits(rng.randint(1, 21)), lambda: str(rng.randint(-2**63 - 5, 2**64 + 5))]
    if oid == 1700:
        return [lambda: rng.choice(["", "-", "+"]) + digits(rng.randint(0, 40)) + rng.choice(["", ".", "." + digits(rng.randint(1, 30))]) +
                rng.choice(["", "", "e" + str(rng.randint(-50, 50)), "E+" + digits(2)])]
    if oid in (701, 700):
        def f():
            m = digits(rng.randint(1, 25))
            return rng.choice(["", "-"]) + m[:1] + "." + m[1:] + "e" + str(rng.randint(-340, 320))
        return [f, lambda: repr(rng.uniform(-1, 1) * 10.0 ** rng.randint(-300, 300)), lambda: "%.9g" % (rng.uniform(-1, 1) * 10.0 ** rng.randint(-44, 38)),
                lambda: digits(rng.randint(1, 30)) + "." + digits(rng.randint(0, 30))]
    if oid in (1114, 1184, 1082, 1083):
        def ts():
            d = "%04d-%02d-%02d" % (rng.randint(0, 10000), rng.randint(0, 13), rng.randint(0, 32))
            t = "%02d:%02d:%02d" % (rng.randint(0, 24), rng.randint(0, 60), rng.randint(0, 61)) + rng.choice(["", "", "." + digits(rng.randint(1, 10))])
            z = rng.choice(["+00", "-07", "+05:30", "+0530", "Z", "+14", "-12:00", "+23:59", " +00", "+1", "+24"])
            return {1082: d, 1083: t, 1114: d + rng.choice([" ", "T", "  "]) + t, 1184: d + " " + t + z}[oid]
        return [ts]
    if oid == 2950:
        def u():
            s = str(uuid.UUID(int=rng.getrandbits(128)))
            k = rng.randint(0, 5)
            return [s, s.upper(), s.replace("-", ""), "{" + s + "}", "urn:uuid:" + s, s[:rng.randint(0, 36)]][k]
        return [u]
    if oid == 17:
        return [lambda: "\\x" + "".join(rng.choice("0123456789abcdefABCDEF") for _ in range(rng.randint(0, 40)))]
    if oid == 3802:
        def gen(d=0):
            k = rng.randint(0, 7 if d < 4 else 4)
            if k == 0:
                return str(rng.randint(-10**9, 10**9))
            if k == 1:
                return rng.choice(["true", "false", "null"])
            if k == 2:
                return "%g" % (rng.random() * 10 ** rng.randint(-8, 8))
            if k in (3, 4):
                return '"' + rng.choice(["abc", "x\\ny", "\\u00e9", "\\ud83d\\ude00", "", "sp ace", "é✓"]) + '"'
            if k in (5, 6):
                return "[" + rng.choice([",", " , "]).join(gen(d + 1) for _ in range(rng.randint(0, 3))) + "]"
            return "{" + ",".join('"k%d": ' % j + gen(d + 1) for j in range(rng.randint(0, 3))) + "}"
        return [gen]
    return []


@pytest.mark.parametrize("oid", [k[0] for k in FUZZ_KINDS])
def test_device_parsers_match_oracle_on_host(emu, oracle_mod, oid):
    _, seeds, alphabet = next(k for k in FUZZ_KINDS if k[0] == oid)
    kind = oracle_mod.kind_for_oid(oid)
    rng = random.Random(oid * 7919)
    gens = _generators(oid, rng)
    n_ok = 0
    for it in range(N_PER_KIND):
        if gens and it % 2:
            s = list(rng.choice(gens)())
            n_mut = rng.choice([0, 0, 1])
        else:
            s = list(rng.choice(seeds))
            n_mut = rng.randint(0, 2)
        for _ in range(n_mut):
            op, p, ch = rng.randint(0, 2), rng.randint(0, len(s)), rng.choice(alphabet)
            if op == 0:
                s.insert(p, ch)
            elif s and op == 1:
                del s[min(p, len(s) - 1)]
            elif s:
                s[min(p, len(s) - 1)] = ch
        text = "".join(s).encode()
        if it % 97 == 0 and text:                       # now and then: invalid UTF-8
            text = text[:len(text) // 2] + b"\xff" + text[len(text) // 2:]
        e, tag, val, aux, heap = oracle_mod.parse_cell(oid, text)
        want = (e, None) if e else (0, decode_cell(tag, val, aux, text, heap))
        n_ok += e == 0
        for fast in (0, 1):
            got = emu(kind, text, fast)
            assert got == want, (oid, text, "fast" if fast else "exact", got, want)
    assert n_ok > N_PER_KIND // 20, (oid, n_ok)


from test_gpu_parity import ARRAY_COLS  # noqa: E402


@pytest.mark.parametrize("oid", [k[0] for k in ARRAY_COLS])
def test_device_array_parser_matches_oracle_on_host(emu, oracle_mod, oid):
    """text.rs:184-249 + element dispatch: the device splitter (array_parse.cuh) against the oracle on mutated
    array literals — quoting, escapes, NULL spellings, empty elements, element parse errors."""
    _, valid, invalid = next(k for k in ARRAY_COLS if k[0] == oid)
    kind = oracle_mod.kind_for_oid(oid)
    if not kind & 0x20:
        pytest.skip("not an array decode class in the oracle")
    seeds = list(valid) + list(invalid) + ["{}", "{NULL}", '{"a,b",c}', '{"\\"q\\"",\\\\}', "{ 1 , 2 }"]
    alphabet = '{}",\\ NULnul0123456789.-+:eabtfx'
    rng = random.Random(oid * 104729)
    n_ok = 0
    n = max(2000, N_PER_KIND // 4)
    for _ in range(n):
        s = list(rng.choice(seeds))
        for _ in range(rng.randint(0, 3)):
            op, p, ch = rng.randint(0, 2), rng.randint(0, len(s)), rng.choice(alphabet)
            if op == 0:
                s.insert(p, ch)
            elif s and op == 1:
                del s[min(p, len(s) - 1)]
            elif s:
                s[min(p, len(s) - 1)] = ch
        text = "".join(s).encode()
        e, tag, val, aux, heap = oracle_mod.parse_cell(oid, text)
        want = (e, None) if e else (0, decode_cell(tag, val, aux, text, heap))
        n_ok += e == 0
        got = emu(kind, text, 1)
        assert got == want, (oid, text, got, want)
    assert n_ok > n // 20, (oid, n_ok)
