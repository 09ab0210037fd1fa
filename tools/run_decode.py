#!/usr/bin/env python
"""Decode one synthetic workload a few times with the stream resident in HBM (for ncu / timing).

  python tools/run_decode.py <workload> <scale> [iters] [stride]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from etl_b200 import abi, decoder, workloads as wl  # noqa: E402

name, scale = sys.argv[1], float(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
stride = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
w = wl.make(name, scale)
stream, stats = w.generate()
st = decoder.Stager(stream.nbytes, stride)
st.append_framed(stream)
dec = decoder.Decoder(0, stream=torch.cuda.current_stream().cuda_stream)
for tid, cols in w.table_schemas().items():
    dec.put_table_schema(tid, cols)
v = st.view()
d_stream = torch.empty(stream.nbytes + 64, dtype=torch.uint8, device="cuda")
d_stream[:stream.nbytes].copy_(torch.from_numpy(st.host_array()))
anch = np.ctypeslib.as_array(abi.C.cast(v.anchors, abi.u64p), shape=(int(v.n_anchors),))
d_anch = torch.from_numpy(np.concatenate([anch, np.array([stream.nbytes], dtype=np.uint64)]).view(np.int64)).cuda()
torch.cuda.synchronize()
for i in range(iters):
    inp = st.view()
    inp.dev_buf, inp.dev_anchors = d_stream.data_ptr(), d_anch.data_ptr()
    with dec.decode_input(inp, to_host=False) as bh:
        s = bh.summary()
    print(f"{name}@{scale:g} iter {i}: {stream.nbytes} B, {stats['frames']} frames, index {s.index_ms:.3f} ms, emit {s.emit_ms:.3f} ms (frames {s.frames_ms:.3f} walk {s.walk_ms:.3f} cells {s.cells_ms:.3f} dead {s.spans_ms:.3f}), "
          f"{stream.nbytes / (s.index_ms + s.emit_ms) / 1e6:.1f} GB/s (index+emit), err={s.first_error.record_index != 2**64 - 1}")
