#!/usr/bin/env python
"""Where does the time of one 8 MiB decode go?  Wall time per stage of the Python call next to the GPU-side event times.

  python tools/batch_probe.py [workload] [scale] [calls]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from etl_b200 import decoder, workloads as wl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 500
timing = (sys.argv[4] != "notiming") if len(sys.argv) > 4 else True
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
w = wl.make(name, scale, n_segments=1)
stream, stats = w.generate()
target = 8 << 20
cuts, pos, n, nxt = [0], 0, int(stream.nbytes), target
mv = memoryview(stream)
while pos + 5 <= n:
    if pos >= nxt:
        cuts.append(pos)
        nxt = pos + target
    pos += 1 + int.from_bytes(mv[pos + 1:pos + 5], "big")
cuts.append(n)
parts = [stream[a:b] for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
staged = [bench.Staged([p], 2048, dev, torch) for p in parts]
dec = decoder.Decoder(0, stream=torch.cuda.current_stream().cuda_stream)
for tid, cols in w.table_schemas().items():
    dec.put_table_schema(tid, cols)
T = {k: [] for k in ("view", "decode", "summary", "free", "total", "gpu_index", "gpu_emit", "gpu_kernel")}
done = 0
while done < calls + 2 * len(staged):
    carry = None
    for st in staged:
        t0 = time.perf_counter()
        inp = st.view(True, carry)
        t1 = time.perf_counter()
        bh = dec.decode_input(inp, to_host=False, timing=timing)
        t2 = time.perf_counter()
        s = bh.summary()
        carry = (int(s.carry_out.in_tx), int(s.carry_out.final_lsn), int(s.carry_out.next_tx_ordinal))
        t3 = time.perf_counter()
        bh.free()
        t4 = time.perf_counter()
        done += 1
        if done > 2 * len(staged):
            for k, v in (("view", t1 - t0), ("decode", t2 - t1), ("summary", t3 - t2), ("free", t4 - t3), ("total", t4 - t0),
                         ("gpu_index", s.index_ms * 1e-3), ("gpu_emit", s.emit_ms * 1e-3), ("gpu_kernel", s.kernel_ms * 1e-3)):
                T[k].append(v)
print(f"{name} x{scale}: {len(parts)} batches of ~8 MiB, {len(T['total'])} calls, launches per decode {s.gpu_launches / max(done, 1):.1f} (cumulative counter)")
for k, v in T.items():
    a = np.sort(np.array(v)) * 1e6
    print(f"  {k:10s} p50 {a[len(a) // 2]:8.1f} us   mean {a.mean():8.1f} us   p99 {a[int(len(a) * 0.99)]:8.1f} us")
