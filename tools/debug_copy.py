"""Per-cell comparison of the device COPY decode with the oracle's row parser (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util
import numpy as np
spec = importlib.util.spec_from_file_location("tgc", os.path.join(ROOT, "tests", "test_gpu_copy.py"))
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
from canon import decode_cell
from etl_b200 import decoder
from oracle import pyoracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
oids, rows = m.synth_rows(n, 1234 + n, None)
dec = decoder.Decoder(0)
dec.put_table_schema(7, m.cols_of(oids))
b = dec.copy_decode(7, rows)
print("first_error", b.first_error, "rows", b.n_rows, "cols", b.n_cols)
bad = 0
for r, row in enumerate(rows):
    e, ecol, cells, text, heap = pyoracle.parse_copy_row(oids, row)
    want = [decode_cell(t, v, a, text, heap) for t, v, a in cells]
    got = m.values_of(b, r)
    if got != want:
        bad += 1
        if bad <= 5:
            for c, (g, w_) in enumerate(zip(got, want)):
                if g != w_:
                    print(f"row {r} col {c}: got {g!r} want {w_!r} raw tag/val/aux {int(b.cell_tag[r*b.n_cols+c])} {int(b.cell_val[r*b.n_cols+c]):#x} {int(b.cell_aux[r*b.n_cols+c])}")
print("rows differing:", bad, "of", n)
