"""pgoutput (proto_version 1) wire-format writer.

Builds the byte stream the decoder consumes: CopyData-framed XLogData / keepalive bodies carrying
Begin/Commit/Relation/Insert/Update/Delete/Truncate/Message/Origin/Type messages.  Grammar:
SURVEY.md Appendix B.  The tuple encoder mirrors the fixture encoder the reference's own tests use
(crates/etl/src/conversions/event.rs:1068-1146); the keepalive layout is pinned by
crates/etl/docs/replication_trace.txt:481-485.

This module is host-side tooling (fixtures, synthetic workloads); it is not on the decode path.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence, Tuple, Union

# A tuple cell on the wire: None → 'n', UNCHANGED → 'u', bytes/str → 't', Binary(b) → 'b'.
UNCHANGED = object()


@dataclass(frozen=True)
class Binary:
    data: bytes


CellIn = Union[None, object, bytes, str, Binary]


def encode_tuple(cells: Sequence[CellIn]) -> bytes:
    """TupleData: i16 ncols, then per column 'n' | 'u' | 't' i32 len bytes | 'b' i32 len bytes."""
    out = bytearray(struct.pack(">h", len(cells)))
    for c in cells:
        if c is None:
            out += b"n"
        elif c is UNCHANGED:
            out += b"u"
        elif isinstance(c, Binary):
            out += b"b" + struct.pack(">i", len(c.data)) + c.data
        else:
            data = c.encode("utf-8") if isinstance(c, str) else bytes(c)
            out += b"t" + struct.pack(">i", len(data)) + data
    return bytes(out)


def begin(final_lsn: int, timestamp: int, xid: int) -> bytes:
    return b"B" + struct.pack(">QqI", final_lsn, timestamp, xid)


def commit(flags: int, commit_lsn: int, end_lsn: int, timestamp: int) -> bytes:
    return b"C" + struct.pack(">bQQq", flags, commit_lsn, end_lsn, timestamp)


def origin(lsn: int, name: str) -> bytes:
    return b"O" + struct.pack(">Q", lsn) + name.encode() + b"\0"


def type_msg(oid: int, namespace: str, name: str) -> bytes:
    return b"Y" + struct.pack(">I", oid) + namespace.encode() + b"\0" + name.encode() + b"\0"


def relation(rel_id: int, namespace: str, name: str, replident: str,
             columns: Sequence[Tuple[int, str, int, int]]) -> bytes:
    """columns: (flags, name, type_oid, typmod); flags bit0 = part of the replica identity."""
    out = bytearray(b"R" + struct.pack(">I", rel_id))
    out += namespace.encode() + b"\0" + name.encode() + b"\0" + replident.encode()
    out += struct.pack(">h", len(columns))
    for flags, cname, oid, typmod in columns:
        out += struct.pack(">b", flags) + cname.encode() + b"\0" + struct.pack(">Ii", oid, typmod)
    return bytes(out)


def insert(rel_id: int, new: Sequence[CellIn]) -> bytes:
    return b"I" + struct.pack(">I", rel_id) + b"N" + encode_tuple(new)


def update(rel_id: int, new: Sequence[CellIn], old: Optional[Sequence[CellIn]] = None,
           key: Optional[Sequence[CellIn]] = None) -> bytes:
    out = bytearray(b"U" + struct.pack(">I", rel_id))
    if old is not None and key is not None:
        raise ValueError("update body cannot contain both old and key tuples")
    if old is not None:
        out += b"O" + encode_tuple(old)
    if key is not None:
        out += b"K" + encode_tuple(key)
    out += b"N" + encode_tuple(new)
    return bytes(out)


def delete(rel_id: int, old: Optional[Sequence[CellIn]] = None,
           key: Optional[Sequence[CellIn]] = None) -> bytes:
    if (old is None) == (key is None):
        raise ValueError("delete body requires exactly one of old / key tuple")
    tag, tup = (b"O", old) if old is not None else (b"K", key)
    return b"D" + struct.pack(">I", rel_id) + tag + encode_tuple(tup)


def truncate(rel_ids: Sequence[int], options: int = 0) -> bytes:
    return b"T" + struct.pack(">ib", len(rel_ids), options) + b"".join(struct.pack(">I", r) for r in rel_ids)


def message(flags: int, lsn: int, prefix: str, content: bytes) -> bytes:
    return b"M" + struct.pack(">bQ", flags, lsn) + prefix.encode() + b"\0" + struct.pack(">i", len(content)) + content


def xlogdata(wal_start: int, wal_end: int, send_time: int, payload: bytes) -> bytes:
    """CopyData body of an XLogData message: 'w' + 24-byte header + pgoutput message."""
    return b"w" + struct.pack(">QQq", wal_start, wal_end, send_time) + payload


def keepalive(wal_end: int, send_time: int, reply: int) -> bytes:
    return b"k" + struct.pack(">QqB", wal_end, send_time, reply)


def frame(body: bytes) -> bytes:
    """CopyData framing as it appears on the TCP wire and in the staged buffer: 'd' int32(len+4)."""
    return b"d" + struct.pack(">i", len(body) + 4) + body


@dataclass
class StreamWriter:
    """Accumulates framed messages, assigning monotonically increasing WAL positions."""
    lsn: int = 0x1000000
    clock: int = 800_000_000_000_000  # µs since 2000-01-01
    chunks: List[bytes] = field(default_factory=list)
    relation_offsets: List[int] = field(default_factory=list)
    size: int = 0

    def emit(self, payload: bytes, wal_start: Optional[int] = None) -> int:
        if wal_start is None:
            wal_start = self.lsn
            self.lsn += 8 + len(payload)
        fr = frame(xlogdata(wal_start, self.lsn, self.clock, payload))
        if payload[:1] == b"R":
            self.relation_offsets.append(self.size)
        self.chunks.append(fr)
        off = self.size
        self.size += len(fr)
        self.clock += 7
        return off

    def emit_keepalive(self, reply: int = 0) -> int:
        fr = frame(keepalive(self.lsn, self.clock, reply))
        self.chunks.append(fr)
        off = self.size
        self.size += len(fr)
        return off

    def emit_raw_frame(self, fr: bytes) -> int:
        self.chunks.append(fr)
        off = self.size
        self.size += len(fr)
        return off

    def bytes(self) -> bytes:
        return b"".join(self.chunks)


def build_anchors(stream: bytes, stride: int) -> List[int]:
    """anchors[k] = offset of the first frame starting at or after k*stride (len(stream) if none).
    Same contract as etl_stage_* (include/etl_decode.h)."""
    n = len(stream)
    n_anchors = max(1, -(-n // stride))
    anchors = [n] * n_anchors
    pos, k = 0, 0
    while pos < n and k < n_anchors:
        while k < n_anchors and k * stride <= pos:
            anchors[k] = pos
            k += 1
        (flen,) = struct.unpack_from(">i", stream, pos + 1)
        pos += 1 + flen
    return anchors


def scan_relation_offsets(stream: bytes) -> List[int]:
    out, pos, n = [], 0, len(stream)
    while pos < n:
        (flen,) = struct.unpack_from(">i", stream, pos + 1)
        if stream[pos + 5:pos + 6] == b"w" and stream[pos + 30:pos + 31] == b"R":
            out.append(pos)
        pos += 1 + flen
    return out
