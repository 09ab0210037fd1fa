#!/usr/bin/env python
"""bench.py — WAL GB/s and events/s of the batched pgoutput decode path on N B200s of one node.

  python bench.py --gpus 1 --steps 5 --warmup 3                      (driver launches N>1 via torchrun)
  python bench.py --impl reference ...                                (the CPU path on the host cores)

A "step" is one pass of the decode hot path over the staged synthetic stream:
  value  — stream + anchor index already resident in HBM, results left in HBM (CUDA-event timed)
  e2e    — the same call through the C ABI with HOST (pinned) buffers: H2D of the stream and its
           anchor index, the kernels, and the D2H of every result plane are inside the timed region
Workload (default c5): BASELINE.json configs[4], "10 GiB synthetic pgoutput buffer, mixed ops +
TOASTed text"; every GPU decodes a full-size (10 GiB) byte-range shard of one longer stream (weak scaling:
rank r owns segments r*64 .. r*64+63), the shard seams are stitched by one all-gather of 48-byte summaries.
Inputs are far larger than L2 (126 MB), so no flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the BASELINE.json size of the workload")
    ap.add_argument("--stride", type=int, default=2048, help="anchor stride of the staged stream")
    ap.add_argument("--cpu-sample-gib", type=float, default=10.0, help="bounded sample for the CPU baseline")
    ap.add_argument("--gen-threads", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while a timed region runs."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self._nv:
            self._stop.clear()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        if self._t:
            self._stop.set()
            self._t.join()

    def report(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def generate_segments(w, segs, threads):
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(threads, len(segs)))) as ex:
        parts = list(ex.map(w.generate_segment, segs))
    return [p[0] for p in parts], [p[1] for p in parts]


def cpu_decode_segments(w, seg_arrays, threads):
    """Decode independent segments with the CPU oracle on `threads` host threads. Returns seconds."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle

    def work(idx):
        o = pyoracle.Oracle()
        for tid, cols in w.table_schemas().items():
            o.put_table_schema(tid, cols)
        n = 0
        for i in idx:
            b = o.decode_raw(seg_arrays[i])
            assert b.first_error.record_index == 2**64 - 1
            n += b.n_records
            o.free(b)
        return n

    threads = max(1, min(threads, len(seg_arrays)))
    chunks = [list(range(t, len(seg_arrays), threads)) for t in range(threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        recs = sum(ex.map(work, chunks))
    return time.perf_counter() - t0, recs


def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path (oracle port — the Rust reference cannot be
    built in this image) on all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from etl_b200 import workloads as wl
    w = wl.make(args.workload, args.scale)
    cores = os.cpu_count() or 1
    seg_bytes = w.bytes_per_segment or (w.segment_capacity() // 2)
    n_sample = max(1, min(w.n_segments, max(cores, int(args.cpu_sample_gib * 4 * GIB / max(seg_bytes, 1)))))
    n_sample = min(n_sample, w.n_segments)
    arrays, stats = generate_segments(w, list(range(n_sample)), args.gen_threads or cores)
    total_bytes = sum(a.nbytes for a in arrays)
    frames = sum(s["frames"] for s in stats)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_decode_segments(w, arrays, cores)
    secs = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_decode_segments(w, arrays, cores)
        secs += dt
    ms = secs / args.steps * 1e3
    val = total_bytes / (secs / args.steps) / 1e9
    sample = f"{n_sample} of {w.n_segments} segments ({total_bytes / GIB:.2f} GiB) of workload {w.name}, {min(cores, n_sample)} threads"
    line = {"impl": "reference", "metric": "wal_decode_throughput", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "events_per_s": frames / (secs / args.steps),
            "config": {"workload": f"{w.name}: {w.description}", "scale": args.scale, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": min(cores, n_sample), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from etl_b200 import abi, decoder, sharding, workloads as wl

    assert torch.cuda.is_available(), "bench.py needs CUDA devices (the decode path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("ETL_NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    # ---- this rank's byte range of the workload
    w = wl.make(args.workload, args.scale)
    S = w.n_segments
    # weak scaling: every GPU decodes a full-size shard (segments rank*S .. rank*S+S-1 of one longer stream)
    my_segs = list(range(rank * S, (rank + 1) * S))
    cores = os.cpu_count() or 8
    t0 = time.perf_counter()
    arrays, stats = generate_segments(w, my_segs, args.gen_threads or max(1, cores // world))
    nbytes = sum(a.nbytes for a in arrays)
    frames = sum(s["frames"] for s in stats)
    stager = decoder.Stager(nbytes, args.stride)
    for a in arrays:
        stager.append_framed(a)
    gen_s = time.perf_counter() - t0
    n_my_segs = len(arrays)
    if not (rank == 0 and world == 1 and not args.no_cpu_baseline):
        arrays = None                                 # the pinned staged copy is all the timed legs need
    host_view = stager.view()
    host_arr = stager.host_array()

    dec = decoder.Decoder(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    for tid, cols in w.table_schemas().items():
        dec.put_table_schema(tid, cols)

    # ---- resident copies for the `value` leg
    d_stream = torch.empty(max(nbytes, 1) + 64, dtype=torch.uint8, device=dev)
    d_stream[:nbytes].copy_(torch.from_numpy(host_arr))
    anchors_np = np.ctypeslib.as_array(abi.C.cast(host_view.anchors, abi.u64p), shape=(int(host_view.n_anchors),))
    d_anchors = torch.from_numpy(np.concatenate([anchors_np, np.array([nbytes], dtype=np.uint64)]).view(np.int64)).to(dev)
    torch.cuda.synchronize()

    def make_input(resident: bool):
        inp = stager.view()
        if resident:
            inp.dev_buf = d_stream.data_ptr()
            inp.dev_anchors = d_anchors.data_ptr()
        return inp

    last = {}

    def step(resident: bool):
        inp = make_input(resident)
        seam = dec.decode_begin(inp, to_host=not resident)
        carry, base = (0, 0, 0), 0
        if world > 1:  # the one exchange step: all-gather of the shard seam summaries
            allw = sharding.all_gather_seam(sharding.seam_to_words(seam), dev)
            carry, base = sharding.carry_for_rank(allw, rank)
        bh = dec.decode_finish(carry, base)
        s = bh.summary()
        if s.first_error.record_index != 2**64 - 1:
            raise RuntimeError(f"decode reported a data error at record {s.first_error.record_index} code {s.first_error.code}")
        last.update(launches=s.gpu_launches, emit_ms=s.emit_ms, index_ms=s.index_ms, kernel_ms=s.kernel_ms,
                    h2d=s.h2d_bytes, d2h=s.d2h_bytes, n_records=seam.n_records, n_cells=seam.n_cells, span_bytes=s.span_bytes)
        bh.free()
        return s

    def timed(resident: bool, steps: int):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        emit, index, launches = [], [], 0
        e0.record()
        for _ in range(steps):
            s = step(resident)
            emit.append((s.frames_ms, s.walk_ms, s.spans_ms, s.kernel_ms, s.cells_ms))
            index.append(s.index_ms)
            launches += s.gpu_launches
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), emit, index, launches

    for _ in range(max(args.warmup, 3)):
        step(True)
    sampler = ClockSampler(local_rank)
    with sampler:
        total_ms, emit_ms, index_ms, launches = timed(True, args.steps)
    ms_per_step = total_ms / args.steps

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step(False)
        e2e_ms, _, _, _ = timed(False, args.steps)
        e2e = {"ms_per_step": e2e_ms / args.steps, "h2d": last["h2d"], "d2h": last["d2h"]}

    # ---- totals over ranks
    tot = torch.tensor([nbytes, frames, last["n_records"], last["n_cells"], e2e["h2d"] if e2e else 0, e2e["d2h"] if e2e else 0],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_bytes, total_frames = float(tot[0].item()), float(tot[1].item())

    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        # the scalar oracle port, one thread (the reference's streaming decode is a single task, apply.rs:839-967)
        budget = int(args.cpu_sample_gib * GIB)
        pick, acc = [], 0
        for i, a in enumerate(arrays):
            if pick and acc + a.nbytes > budget:
                break
            pick.append(i)
            acc += a.nbytes
        secs, recs = cpu_decode_segments(w, [arrays[i] for i in pick], 1)
        cpu_baseline = {"value": acc / secs / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                        "sample": f"first {len(pick)} of {len(arrays)} segments of this run's stream ({acc / GIB:.2f} GiB, {recs} msgs) in {secs:.1f} s",
                        "events_per_s": recs / secs}

    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic bytes: every byte of the staged stream once + the anchor index
        algo_bytes = nbytes + 8 * (int(host_view.n_anchors) + 1)
        km = np.mean(np.array(emit_ms, dtype=np.float64), axis=0)           # frames, walk, dead-segment pass, critical path (ms, rank 0)
        kern = {"k_act*+k_index+k_scan+k_tile_prefix": float(np.mean(index_ms)), "k_frames": float(km[0]),
                "k_bin_scan+k_perm+k_walk": float(km[1]), "k_cells+k_copy": float(km[4]),
                "k_utf8_dead (side stream, overlapped)": float(km[2])}
        span_bytes = int(last["span_bytes"])
        # bytes each kernel is responsible for: k_utf8_dead streams the segments without a frame start (the inside of
        # TOAST-sized values), k_walk everything else in the DML tuples, k_frames / k_index the frame heads (counted
        # with k_walk's share here)
        kbytes = {"k_utf8_dead": span_bytes, "k_walk": algo_bytes - span_bytes, "k_cells": algo_bytes - span_bytes}
        ktime = {"k_utf8_dead": float(km[2]), "k_walk": float(km[1]), "k_cells": float(km[4])}
        dominant = max(ktime, key=lambda k: ktime[k])
        pipeline_ms = float(km[3])                                          # index + records passes incl. the join with the side stream
        emit_avg = ktime[dominant]
        achieved = kbytes[dominant] / (emit_avg * 1e-3) / 1e9
        traffic, alone_us = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get(f"{w.name}/{dominant}")
            alone_us = tj.get(f"{w.name}/{dominant}/serialised_us") if args.scale == 1.0 else None
        except Exception:
            pass
        line = {
            "metric": "wal_decode_throughput", "value": total_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "events_per_s": total_frames / (ms_per_step * 1e-3),
            "config": {"workload": f"{w.name}: {w.description}", "scale": args.scale, "bytes_total": int(total_bytes),
                       "msgs_total": int(total_frames), "records": int(tot[2].item()), "cells": int(tot[3].item()),
                       "parallelism": f"{n_gpus} byte-range shards of {S} segments each (one per GPU), one seam all-gather" if n_gpus > 1 else "single GPU",
                       "anchor_stride": args.stride, "l2_policy": "inputs (>=1.25 GiB per GPU) larger than the 126 MB L2",
                       "generate_s": round(gen_s, 2)},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kbytes[dominant], "avg_launch_ms": emit_avg,
                         "alone_under_ncu": ({"ms": alone_us / 1e3, "achieved": kbytes[dominant] / (alone_us * 1e-6) / 1e9,
                                              "frac": kbytes[dominant] / (alone_us * 1e-6) / 1e9 / peak,
                                              "source": "profiles/traffic.json (ncu launch list of this command: kernels serialised, not sharing HBM)"}
                                             if alone_us else None),
                         "kernels_ms": kern,
                         "pipeline": {"algorithmic_bytes": algo_bytes, "ms": pipeline_ms,
                                      "achieved": algo_bytes / (pipeline_ms * 1e-3) / 1e9,
                                      "frac": algo_bytes / (pipeline_ms * 1e-3) / 1e9 / peak},
                         "k_utf8_dead": {"algorithmic_bytes": span_bytes,
                                         "achieved": span_bytes / max(ktime["k_utf8_dead"], 1e-6) / 1e6,
                                         "frac": span_bytes / max(ktime["k_utf8_dead"], 1e-6) / 1e6 / peak,
                                         "note": "timed while the index/records passes run on the main stream"}},
            "gpu_launches": launches,
            "clocks": sampler.report(),
        }
        if e2e:
            line["e2e"] = {"value": total_bytes / (e2e["ms_per_step"] * 1e-3) / 1e9, "unit": "GB/s",
                           "h2d_bytes_per_step": int(tot[4].item()), "d2h_bytes_per_step": int(tot[5].item()),
                           "ms_per_step": e2e["ms_per_step"]}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
