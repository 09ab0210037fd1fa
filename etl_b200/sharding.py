"""Multi-GPU byte-range sharding: seam summaries and their fold.

The staged stream is cut at record starts into one contiguous byte range per rank.  Each rank runs
the index + scan passes locally (`etl_dec_decode_begin`) which yields a fixed-size seam summary —
the stream-state transformer of its shard (apply.rs:600-626, 1927-2006: last Begin's final_lsn,
ordinal count, open/closed) plus record / cell / heap counts.  ONE all-gather of those summaries
is the only exchange step; every rank then folds the summaries of the ranks before it into its
carry-in state and record-index base and runs the emit pass (`etl_dec_decode_finish`).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

SEAM_WORDS = 6  # n_records, n_cells, heap_bytes, lsn, ord, flags(has_begin | closed << 1)


def seam_to_words(seam) -> np.ndarray:
    return np.array([seam.n_records, seam.n_cells, seam.heap_bytes, seam.lsn, seam.ord,
                     int(seam.has_begin) | (int(seam.closed) << 1)], dtype=np.uint64)


def fold_state(state: Tuple[int, int, int], words: Sequence[int]) -> Tuple[int, int, int]:
    """Apply one shard's transformer to (in_tx, final_lsn, next_tx_ordinal)."""
    in_tx, lsn, ord_ = state
    has_begin, closed = int(words[5]) & 1, (int(words[5]) >> 1) & 1
    if has_begin:
        return (0 if closed else 1, int(words[3]), int(words[4]))
    return (0 if closed else in_tx, lsn, ord_ + int(words[4]))


def carry_for_rank(all_words: np.ndarray, rank: int, carry_in: Tuple[int, int, int] = (0, 0, 0)) -> Tuple[Tuple[int, int, int], int]:
    """Carry-in stream state and global record-index base for `rank` from the gathered summaries."""
    state, base = carry_in, 0
    for r in range(rank):
        state = fold_state(state, all_words[r])
        base += int(all_words[r][0])
    return state, base


def all_gather_seam(words: np.ndarray, device=None) -> np.ndarray:
    """The one exchange step: all-gather of the fixed-size seam summaries (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.from_numpy(words.view(np.int64).copy())
    if device is not None:
        t = t.to(device, non_blocking=True)
    out = torch.empty(world * SEAM_WORDS, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().view(np.uint64).reshape(world, SEAM_WORDS)


def cut_points(anchors: np.ndarray, length: int, world: int) -> List[int]:
    """Byte offsets (record starts taken from the anchor index) cutting [0, length) into `world` ranges."""
    cuts = [0]
    stride_pos = np.linspace(0, len(anchors), world + 1)[1:-1]
    for p in stride_pos:
        cuts.append(int(anchors[min(int(p), len(anchors) - 1)]))
    cuts.append(length)
    return cuts
