"""CPU-side checks of the C ABI: the library loads, exports every symbol include/etl_decode.h
declares, the stager indexes frames correctly, and the decode path refuses to run without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from etl_b200 import abi, pgoutput as pg, workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = abi.load()
    header = open(os.path.join(ROOT, "include", "etl_decode.h")).read()
    declared = set(re.findall(r"\b(etl_(?:dec|stage|shim)_[a-z_0-9]+)\s*\(", header))
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.etl_dec_abi_version() == 2


def test_struct_layouts_match_header_sizes(tmp_path):
    """sizeof / field offsets the C compiler gives for include/etl_decode.h vs the ctypes mirror (guards against
    silent ABI drift in the bindings)."""
    import subprocess
    pairs = [("etl_column_schema", abi.ColumnSchema), ("etl_stream_state", abi.StreamState), ("etl_first_error", abi.FirstError),
             ("etl_dec_input", abi.DecInput), ("etl_dec_seam", abi.Seam), ("etl_dec_planes", abi.Planes),
             ("etl_dec_summary", abi.Summary), ("etl_dec_schema_info", abi.SchemaInfo)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "etl_decode.h"', 'int main(void) {']
    for cname, ct in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi_sizes.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in pairs:
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


@pytest.mark.parametrize("stride", [256, 2048, 32768])
def test_stager_anchor_index_matches_definition(stride):
    from etl_b200.decoder import Stager
    w = wl.make("c5", 0.0005, n_segments=2)
    stream, stats = w.generate()
    st = Stager(stream.nbytes, stride)
    st.append_framed(stream)
    v = st.view()
    anchors = np.ctypeslib.as_array(C.cast(v.anchors, abi.u64p), shape=(int(v.n_anchors),)).copy()
    rels = np.ctypeslib.as_array(C.cast(v.relation_offsets, abi.u64p), shape=(int(v.n_relations),)).copy()
    raw = stream.tobytes()
    assert anchors.tolist() == pg.build_anchors(raw, stride)
    assert rels.tolist() == pg.scan_relation_offsets(raw)
    assert v.len == len(raw) and v.anchor_stride == stride
    assert np.array_equal(st.host_array(), stream)
    st.close()


def test_stager_append_bodies_equals_framed():
    from etl_b200.decoder import Stager
    w = pg.StreamWriter()
    w.emit(pg.begin(5, 6, 7))
    w.emit(pg.relation(9, "public", "t", "d", [(1, "id", 20, -1)]))
    w.emit_keepalive()
    w.emit(pg.commit(0, 5, 13, 8))
    raw = w.bytes()
    st = Stager(1 << 16, 256)
    pos = 0
    while pos < len(raw):
        n = int.from_bytes(raw[pos + 1:pos + 5], "big")
        st.append(raw[pos + 5:pos + 1 + n])
        pos += 1 + n
    assert st.host_array().tobytes() == raw
    v = st.view()
    assert v.n_relations == 1 and v.n_anchors == -(-len(raw) // 256)
    # the frame-length hint: the longest frame, 'd' + length field + body (both ways of staging agree)
    longest, pos = 0, 0
    while pos < len(raw):
        n = int.from_bytes(raw[pos + 1:pos + 5], "big")
        longest = max(longest, 1 + n)
        pos += 1 + n
    assert v.max_frame_len == longest
    st2 = Stager(1 << 16, 256)
    st2.append_framed(raw)
    assert st2.view().max_frame_len == longest
    st2.reset()
    assert st2.view().max_frame_len == 0
    st2.close()
    st.close()


def test_decoder_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from etl_b200.decoder import Decoder, DecodeError
    with pytest.raises(DecodeError):
        Decoder(0)
