// wal_kernels.cuh — sm_100a kernels of the batched pgoutput decode path.
//
//   k_index   (pass A)  one thread per anchor segment walks the 'd'+len32 frame chain from global
//                       memory (only frame heads are touched), classifies frames and reduces a
//                       per-tile summary {records, cells, heap bytes, stream-state transformer}.
//   k_scan    (pass B)  single-block exclusive scan of the per-group summaries (the state
//                       transformer is associative, so commit_lsn / tx_ordinal become a scan).
//   k_emit    (pass C)  one CTA per 32 KiB tile: coalesced 16-byte loads stage the tile in shared
//                       memory, segment walkers rebuild the frame list there, then thread-per-frame
//                       parsing of TupleData cells writes the record / cell / heap planes; large
//                       text cells are validated block-cooperatively.
//
// Reference semantics: apply.rs:1687-2248 (state machine), event.rs:376-979 (tuples → rows),
// text.rs:28-173 (cells).  HBM-bound integer/byte work — no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cell_parsers.cuh"

namespace etl {

constexpr int kTileBytes = 32768;      // nominal tile = kTileBytes of stream (frames that START inside it)
constexpr int kTileCap = 36864;        // shared-memory window; bytes past it are read from global
constexpr int kEmitThreads = 256;
constexpr int kMaxTileFrames = 2048;   // > kTileBytes/23 + segments
constexpr int kIndexThreads = 256;
constexpr int kBigCell = 512;          // text cells at least this long are validated cooperatively
constexpr int kBigQueue = 128;

// ---- stream-state transformer (apply.rs:600-626, 1927-2006) + counters; associative under fold()
struct Summ {
  uint64_t lsn;      // final_lsn of the last Begin (valid if HAS_B)
  uint64_t ord;      // HAS_B: next ordinal after the span; else number of ordinal consumers in the span
  uint64_t n_cells;
  uint64_t heap;
  uint32_t n_rec;
  uint32_t flags;    // 1 HAS_B, 2 CLOSED (a Commit follows the last Begin / any Commit if no Begin)
};
constexpr uint32_t S_HAS_B = 1, S_CLOSED = 2;

__host__ __device__ __forceinline__ Summ summ_identity() { return Summ{0, 0, 0, 0, 0, 0}; }
__host__ __device__ __forceinline__ Summ fold(const Summ& a, const Summ& b) {
  Summ r;
  r.n_rec = a.n_rec + b.n_rec;
  r.n_cells = a.n_cells + b.n_cells;
  r.heap = a.heap + b.heap;
  if (b.flags & S_HAS_B) { r.flags = b.flags; r.lsn = b.lsn; r.ord = b.ord; }
  else {
    r.flags = (a.flags & S_HAS_B) | ((b.flags & S_CLOSED) ? S_CLOSED : (a.flags & S_CLOSED));
    r.lsn = a.lsn;
    r.ord = a.ord + b.ord;
  }
  return r;
}

struct DevSchema {
  uint32_t table_id;
  uint32_t n_cols;
  uint32_t n_ident;
  uint32_t col_base;      // into col_kind / col_flags
  uint64_t effective_off; // stream offset from which this version applies
  uint32_t batch_index;   // index reported in rec_schema
  uint32_t has_heap;      // any numeric / bytea / uuid / array column
};

struct DecodeParams {
  const uint8_t* buf;
  uint64_t len;
  const uint64_t* anchors;   // n_anchors + 1 entries (last = len)
  uint32_t n_anchors;
  uint32_t anchor_stride;
  uint32_t segs_per_tile;
  uint32_t n_tiles;
  uint32_t tiles_per_group;  // tiles folded per k_index CTA
  uint32_t n_groups;
  const DevSchema* schemas;  // sorted by (table_id, effective_off)
  uint32_t n_schemas;
  const uint8_t* col_kind;
  const uint8_t* col_flags;  // bit0 nullable, bit1 identity
  // pass A outputs
  uint32_t* seg_frames;      // frames starting in each segment
  Summ* tile_summ;           // per tile
  Summ* group_summ;          // per group of tiles
  Summ* group_prefix;        // exclusive prefix per group (pass B)
  Summ* total;               // [0] = fold of everything (shard seam summary)
  Summ* tile_prefix;         // exclusive prefix per tile (pass B2), carry not included
  unsigned int* tile_counter;  // dynamic tile scheduler of the emit pass
  // carry-in (known when pass C runs)
  Summ carry;
  uint64_t record_index_base;  // global index of this shard's first record (multi-GPU)
  // outputs
  uint64_t* rec_off; uint8_t* rec_kind; uint8_t* rec_flags; uint32_t* rec_rel; int32_t* rec_schema;
  uint64_t* rec_start_lsn; uint64_t* rec_commit_lsn; uint64_t* rec_tx_ordinal; uint64_t* rec_cell_base;
  uint8_t* cell_tag; uint64_t* cell_val; uint32_t* cell_aux;
  uint8_t* heap;
  unsigned long long* heap_top;     // bump pointer: rounds of the emit pass reserve their heap bytes here
  uint64_t heap_cap;                // 0 when no schema of the batch has a heap-kind column
  unsigned long long* first_error;  // atomicMin key: rec_index << 24 | seq << 6 | code
  unsigned long long* metrics;      // [0] insert bytes [1] update bytes [2] delete bytes [3] events
  // Relation frames rejected on the host (missing stored schema / unknown columns / malformed)
  const uint64_t* rel_error_off; const uint32_t* rel_error_code; const uint32_t* rel_error_seq; uint32_t n_rel_errors;
};

// ---- big-endian readers on byte-addressed (unaligned) generic pointers
__device__ __forceinline__ uint32_t be32(const uint8_t* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}
__device__ __forceinline__ uint64_t be64(const uint8_t* p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
__device__ __forceinline__ uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | (uint32_t)p[1]; }

// error sequencing inside one record (same numbering as the oracle)
constexpr uint32_t SEQ_MALFORMED = 0, SEQ_STATE = 1, SEQ_TABLE = 2, SEQ_OLD_SHAPE = 0x10000, SEQ_NEW_SHAPE = 0x20000;
__device__ __forceinline__ uint32_t seq_old_cell(uint32_t i) { return 0x10001u + i; }
__device__ __forceinline__ uint32_t seq_new_cell(uint32_t i) { return 0x20001u + i; }
__device__ __forceinline__ void report_error(const DecodeParams& P, uint64_t rec_index, uint32_t seq, uint32_t code) {
  unsigned long long key = ((unsigned long long)rec_index << 24) | ((unsigned long long)(seq & 0x3FFFFu) << 6) | (code & 63u);
  atomicMin(P.first_error, key);
}

// schema version for (table_id, frame offset): last entry with that table and effective_off <= off
__device__ __forceinline__ const DevSchema* find_schema(const DecodeParams& P, uint32_t table_id, uint64_t off) {
  int lo = 0, hi = (int)P.n_schemas;  // upper_bound on (table_id, off)
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    const DevSchema& s = P.schemas[mid];
    bool le = (s.table_id < table_id) || (s.table_id == table_id && s.effective_off <= off);
    if (le) lo = mid + 1; else hi = mid;
  }
  if (lo == 0) return nullptr;
  const DevSchema* s = &P.schemas[lo - 1];
  return s->table_id == table_id ? s : nullptr;
}

// ---- frame head: everything the index pass and the emit pass agree on
struct FrameHead {
  uint32_t flen;       // CopyData length field (counts itself)
  uint32_t kind;       // pgoutput tag, 'k' for keepalive, 0 when malformed at the frame level
  uint32_t rel;        // relation id (R/I/U/D), relation count (T)
  uint32_t old_tag;    // 'O' | 'K' | 0 for U/D
  bool malformed;
};
// p points at the frame's 'd'; avail = bytes from p to the end of the stream
__device__ __forceinline__ FrameHead read_head(const uint8_t* p, uint64_t avail) {
  FrameHead h;
  h.kind = 0; h.rel = 0; h.old_tag = 0; h.malformed = true; h.flen = 4;
  if (avail < 5 || p[0] != 'd') { h.flen = (uint32_t)(avail > 0 ? avail - 1 : 0); return h; }
  uint32_t flen = be32(p + 1);
  if (flen < 4 || 1ull + flen > avail) { h.flen = (uint32_t)(avail - 1); return h; }  // chain ends here
  h.flen = flen;
  uint32_t blen = flen - 4;
  if (blen < 1) return h;
  uint32_t t = p[5];
  if (t == 'k') { if (blen < 18) return h; h.kind = 'k'; h.malformed = false; return h; }
  if (t != 'w' || blen < 26) return h;
  uint32_t tag = p[30];
  uint32_t mlen = blen - 26;  // message bytes after the tag
  h.kind = tag;
  switch (tag) {
    case 'B': if (mlen < 20) return h; break;
    case 'C': if (mlen < 25) return h; break;
    case 'R': case 'I': case 'U': case 'D':
      if (mlen < 5) return h;
      h.rel = be32(p + 31);
      if (tag == 'U' || tag == 'D') { uint32_t tt = p[35]; if (tt == 'O' || tt == 'K') h.old_tag = tt; }
      break;
    case 'T': {
      if (mlen < 5) return h;
      int32_t n = (int32_t)be32(p + 31);
      if (n > 0 && (uint64_t)n * 4 > (uint64_t)mlen - 5) return h;
      h.rel = n > 0 ? (uint32_t)n : 0u;
      break;
    }
    case 'O': case 'Y': case 'M': break;
    default: return h;  // unknown tag
  }
  h.malformed = false;
  return h;
}

// output cells of a frame (depends only on kind / old tag / schema — never on the tuple contents)
__device__ __forceinline__ uint32_t frame_out_cells(const FrameHead& h, const DevSchema* s) {
  switch (h.kind) {
    case 'B': return 2;
    case 'C': return 3;
    case 'T': return 1 + h.rel;
    case 'I': return s ? s->n_cols : 0;
    case 'U': return s ? (s->n_cols + (h.old_tag == 'O' ? s->n_cols : (h.old_tag == 'K' ? s->n_ident : 0))) : 0;
    case 'D': return s ? (h.old_tag == 'O' ? s->n_cols : (h.old_tag == 'K' ? s->n_ident : 0)) : 0;
    default: return 0;
  }
}
__device__ __forceinline__ Summ frame_state_elem(const FrameHead& h, const uint8_t* p) {
  Summ e = summ_identity();
  e.n_rec = 1;
  if (h.malformed) return e;
  switch (h.kind) {
    case 'B': e.flags = S_HAS_B; e.lsn = be64(p + 31); e.ord = 1; break;
    case 'C': e.flags = S_CLOSED; e.ord = 1; break;
    case 'R': case 'I': case 'U': case 'D': case 'T': e.ord = 1; break;
    default: break;
  }
  return e;
}

// heap bytes a text cell of `kind` and length n may need (upper bound; identical in passes A and C)
__device__ __forceinline__ uint32_t cell_heap_bound(uint32_t kind, uint32_t n) {
  switch (kind) {
    case ETL_K_NUMERIC: return numeric_heap_bound(n);
    case ETL_K_BYTES: return bytea_heap_bound(n);
    case ETL_K_UUID: return 16;
    default: return 0;
  }
}

// Walk one TupleData, calling f(col_index, tag, value_ptr, len) per cell. Returns bytes consumed
// or 0 if the tuple runs past `end` / has an unknown cell tag (malformed frame).
template <typename F>
__device__ __forceinline__ uint32_t walk_tuple(const uint8_t* p, const uint8_t* end, int32_t* ncols_out, F&& f) {
  if (p + 2 > end) return 0;
  int32_t n = (int32_t)(int16_t)be16(p);
  if (n < 0) n = 0;
  *ncols_out = n;
  const uint8_t* q = p + 2;
  for (int32_t i = 0; i < n; i++) {
    if (q + 1 > end) return 0;
    uint32_t tag = *q++;
    if (tag == 'n' || tag == 'u') { f(i, tag, q, 0u); continue; }
    if (tag != 't' && tag != 'b') return 0;
    if (q + 4 > end) return 0;
    int32_t l = (int32_t)be32(q);
    q += 4;
    if (l < 0 || (uint64_t)l > (uint64_t)(end - q)) return 0;
    f(i, tag, q, (uint32_t)l);
    q += l;
  }
  return (uint32_t)(q - p);
}

// heap bytes reserved for a DML frame = sum of cell_heap_bound over its text cells, mapped to
// columns positionally (dense key tuples: k-th cell → k-th identity column).  An upper bound on
// what the emit pass allocates; both passes call this same function so prefixes agree exactly.
__device__ __noinline__ uint32_t frame_heap_bytes(const DecodeParams& P, const FrameHead& h, const DevSchema* s,
                                                  const uint8_t* p) {
  if (!s || !s->has_heap) return 0;
  const uint8_t* end = p + 1 + h.flen;
  const uint8_t* kinds = P.col_kind + s->col_base;
  const uint8_t* flags = P.col_flags + s->col_base;
  const uint32_t n_cols = s->n_cols, n_ident = s->n_ident;
  uint32_t total = 0;
  const uint8_t* q = p + 36;  // first tuple (after 'N' / 'O' / 'K' marker at p[35])
  const int n_tuples = (h.kind == 'U' && h.old_tag) ? 2 : 1;
  if (h.kind == 'U' && !h.old_tag && p[35] != 'N') return 0;
  for (int t = 0; t < n_tuples; t++) {
    if (q + 2 > end) return total;
    const bool dense_key = (t == 0) && (h.kind != 'I') && h.old_tag == 'K' && (uint32_t)(int32_t)(int16_t)be16(q) == n_ident;
    uint32_t cmap = 0;
    int32_t nc;
    uint32_t used = walk_tuple(q, end, &nc, [&](int32_t i, uint32_t tag, const uint8_t*, uint32_t len) {
      uint32_t col = (uint32_t)i;
      if (dense_key) { while (cmap < n_cols && !(flags[cmap] & 2)) cmap++; col = cmap++; }
      if (tag == 't' && col < n_cols) total += cell_heap_bound(kinds[col], len);
    });
    if (!used) return total;
    q += used;
    if (t == 0 && n_tuples == 2) { if (q >= end || *q != 'N') return total; q++; }
  }
  return total;
}

// structure of a DML message body (what LogicalReplicationMessage::parse would reject) + Σ text lengths
__device__ __noinline__ bool dml_structure_ok(const FrameHead& h, const uint8_t* p, unsigned long long* tbytes) {
  const uint8_t* end = p + 1 + h.flen;
  const uint8_t* t = p + 35;
  unsigned long long tb = 0;
  int32_t nc;
  auto count = [&](int32_t, uint32_t tag, const uint8_t*, uint32_t l) { if (tag == 't' || tag == 'b') tb += l; };
  if (t >= end) return false;
  uint32_t tt = *t++;
  if (h.kind == 'I') { if (tt != 'N') return false; }
  else if (h.kind == 'U') {
    if (tt == 'O' || tt == 'K') {
      uint32_t used = walk_tuple(t, end, &nc, count);
      if (!used) return false;
      t += used;
      if (t >= end || *t != 'N') return false;
      t++;
    } else if (tt != 'N') return false;
  } else if (tt != 'O' && tt != 'K') return false;
  if (!walk_tuple(t, end, &nc, count)) return false;
  *tbytes = tb;
  return true;
}
__device__ __forceinline__ const uint8_t* cstr_end(const uint8_t* q, const uint8_t* end) {
  while (q < end && *q) q++;
  return q < end ? q + 1 : nullptr;
}

// ================================================================================================
// pass A: index.  grid = n_groups, block = tiles_per_group * segs_per_tile threads.
__global__ void __launch_bounds__(kIndexThreads) k_index(DecodeParams P) {
  const uint32_t spt = P.segs_per_tile;
  const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
  Summ acc = summ_identity();
  uint32_t nframes = 0;
  if (seg < P.n_anchors) {
    uint64_t pos = P.anchors[seg];
    const uint64_t stop = P.anchors[seg + 1];
    while (pos < stop) {
      const uint8_t* p = P.buf + pos;
      FrameHead h = read_head(p, P.len - pos);
      Summ e = frame_state_elem(h, p);
      if (!h.malformed) {
        const DevSchema* s = nullptr;
        if (h.kind == 'I' || h.kind == 'U' || h.kind == 'D') s = find_schema(P, h.rel, pos);
        e.n_cells = frame_out_cells(h, s);
      }
      acc = fold(acc, e);
      nframes++;
      pos += 1ull + h.flen;
    }
    P.seg_frames[seg] = nframes;
  }
  // fold across the segments of each tile, then across the tiles of the group (ordered shuffles)
  // generic ordered fold over the block through shared memory (blockDim <= 256)
  __shared__ Summ sh[kIndexThreads];
  sh[threadIdx.x] = acc;
  __syncthreads();
  // per-tile fold by the first lane of each tile
  const uint32_t tile_in_block = threadIdx.x / spt;
  if (threadIdx.x % spt == 0) {
    Summ t = summ_identity();
    for (uint32_t k = 0; k < spt && threadIdx.x + k < blockDim.x; k++) t = fold(t, sh[threadIdx.x + k]);
    uint32_t tile = blockIdx.x * P.tiles_per_group + tile_in_block;
    if (tile < P.n_tiles) P.tile_summ[tile] = t;
    sh[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Summ g = summ_identity();
    for (uint32_t k = 0; k < P.tiles_per_group; k++) g = fold(g, sh[k * spt]);
    P.group_summ[blockIdx.x] = g;
  }
}

// pass B2: per-tile exclusive prefix = group prefix ⊕ earlier tiles of the group (thread per tile)
__global__ void __launch_bounds__(256) k_tile_prefix(DecodeParams P) {
  const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= P.n_tiles) return;
  const uint32_t g = tile / P.tiles_per_group;
  Summ pre = P.group_prefix[g];
  for (uint32_t t = g * P.tiles_per_group; t < tile; t++) pre = fold(pre, P.tile_summ[t]);
  P.tile_prefix[tile] = pre;
}

// ================================================================================================
// pass B: exclusive scan of group summaries (single CTA; n_groups is len / (tiles_per_group*32 KiB))
__global__ void __launch_bounds__(512) k_scan(DecodeParams P) {
  __shared__ Summ sh[512];
  const uint32_t n = P.n_groups;
  const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = threadIdx.x * per, hi = min(lo + per, n);
  Summ acc = summ_identity();
  for (uint32_t i = lo; i < hi; i++) acc = fold(acc, P.group_summ[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials with the non-commutative fold
  for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
    Summ v = sh[threadIdx.x];
    if (threadIdx.x >= d) v = fold(sh[threadIdx.x - d], v);
    __syncthreads();
    sh[threadIdx.x] = v;
    __syncthreads();
  }
  Summ run = threadIdx.x ? sh[threadIdx.x - 1] : summ_identity();
  for (uint32_t i = lo; i < hi; i++) {
    P.group_prefix[i] = run;
    run = fold(run, P.group_summ[i]);
  }
  if (threadIdx.x == blockDim.x - 1) P.total[0] = sh[blockDim.x - 1];
}

// ================================================================================================
// pass C: emit.  Persistent CTAs pull 32 KiB tiles from a global counter.  Per tile:
//   1. stage the tile in shared memory (coalesced 16-byte loads)
//   2. frame list (one walker thread per anchor segment, in shared memory)
//   3. per 256-frame chunk: classify heads → four u32 warp-shuffle scans (cells, ordinal consumers,
//      last Begin, last Commit) give every frame its cell base and stream state → record plane
//   4. DML frames: register-resident thread-per-frame WALKERS follow the TupleData chain and emit a
//      16-byte descriptor per text cell into a shared batch (NULL / unchanged-TOAST cells are
//      resolved by the walker); the batch is counting-sorted by decode class so warps run one
//      parser at a time; thread-per-CELL parsing; long text is validated by a warp (≥ 96 B) or by
//      the whole CTA (≥ 2 KiB) with 16-byte loads.
struct CellDesc {
  uint32_t toff;      // offset of the value bytes relative to the tile start
  uint32_t len;       // value length
  uint32_t dest_rel;  // output cell index relative to the tile's first cell
  uint32_t meta;      // frame slot (8) | kind (8) | is_new (1) | wire index (15); 0xFFFFFFFF = empty
};
struct WideText {
  const uint8_t* ptr;
  uint32_t len;
  uint32_t seq;
  uint32_t rec_local;
};

constexpr int kDescCap = 1024;
constexpr int kDescPerThread = kDescCap / kEmitThreads;
constexpr int kWideCap = 128;
constexpr int kBigCap = 16;
constexpr int kWideLen = 96;     // text cells at least this long are validated by a warp
constexpr int kBigLen = 2048;    // ... and these by the whole CTA
constexpr uint32_t kCellUnresolved = 253;  // internal: unchanged-TOAST cell awaiting its old value

struct EmitShared {
  alignas(16) uint8_t tile[kTileCap + 32];
  uint16_t foff[kMaxTileFrames];            // frame starts relative to the tile start (< 32 KiB)
  CellDesc desc[kDescCap];
  uint16_t perm[kDescCap];                  // descriptor order after the counting sort by kind
  const uint8_t* fi_base[kEmitThreads];     // per chunk slot: frame bytes (window or global)
  uint64_t fi_off[kEmitThreads];            // absolute stream offset of the frame
  uint64_t b_lsn[kEmitThreads];             // Begin frames of the chunk: final_lsn
  uint32_t b_cons[kEmitThreads];            // ... and inclusive ordinal-consumer count
  uint32_t fi_rec[kEmitThreads];            // shard-local record index
  WideText wide[kWideCap];
  WideText big[kBigCap];
  uint32_t seg_base[132];
  uint32_t ws_cells[kEmitThreads / 32], ws_cons[kEmitThreads / 32];
  int32_t ws_b[kEmitThreads / 32], ws_c[kEmitThreads / 32];
  uint32_t hist[32], kstart[32];
  uint32_t carry_cells, carry_cons, carry_b_cons;
  int32_t carry_b, carry_c;
  uint64_t carry_b_lsn;
  uint32_t n_desc, n_wide, wide_next, n_big, n_sorted, round_heap_need, round_heap_used, tile_idx;
  unsigned long long round_heap_base;
  unsigned long long metrics[4];
};

// unaligned little-endian 8-byte load built from aligned 32-bit words (works on the shared window
// and on global memory; may touch up to 3 bytes before and 11 after p — buffers are padded)
__device__ __forceinline__ uint64_t ld64u(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  const uint32_t sh = (uint32_t)(a & 3u) * 8u;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
  return ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
}
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ uint64_t bswap64(uint64_t v) { return ((uint64_t)bswap32((uint32_t)v) << 32) | bswap32((uint32_t)(v >> 32)); }

// UTF-8 check of a short/medium cell with word loads: all-ASCII words pass immediately
__device__ __forceinline__ bool utf8_valid_fast(const uint8_t* s, uint32_t n) {
  uint32_t hi = 0;
  uint32_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t x = ld64u(s + i); hi |= (uint32_t)(x >> 32) | (uint32_t)x; }
  if (i < n) {
    uint64_t x = ld64u(s + i);
    x &= (1ull << (8 * (n - i))) - 1ull;  // 1..7 valid bytes
    hi |= (uint32_t)(x >> 32) | (uint32_t)x;
  }
  if (!(hi & 0x80808080u)) return true;
  return utf8_valid(s, n);
}

// text.rs:28-173 dispatch for one text cell (bytes already UTF-8 validated). `soff` = absolute
// stream offset of the value bytes.
__device__ __forceinline__ uint32_t parse_text_cell(uint32_t kind, const uint8_t* s, uint32_t n, uint64_t soff,
                                                    HeapCursor& hc, CellOut& o) {
  o.aux = 0;
  int64_t iv;
  uint32_t e;
  switch (kind) {
    case ETL_K_STRING: o.tag = ETL_CELL_STRING; o.val = soff; o.aux = n; return 0;
    case ETL_K_I32: e = parse_int(s, n, true, 2147483647ull, 2147483648ull, &iv); o.tag = ETL_CELL_I32; o.val = (uint64_t)iv; return e;
    case ETL_K_I64: e = parse_int(s, n, true, 9223372036854775807ull, 9223372036854775808ull, &iv); o.tag = ETL_CELL_I64; o.val = (uint64_t)iv; return e;
    case ETL_K_I16: e = parse_int(s, n, true, 32767ull, 32768ull, &iv); o.tag = ETL_CELL_I16; o.val = (uint64_t)iv; return e;
    case ETL_K_U32: e = parse_int(s, n, false, 4294967295ull, 0ull, &iv); o.tag = ETL_CELL_U32; o.val = (uint64_t)iv; return e;
    case ETL_K_BOOL:
      if (n == 1 && (s[0] == 't' || s[0] == 'f')) { o.tag = ETL_CELL_BOOL; o.val = s[0] == 't'; return 0; }
      return ETL_E_BOOL;
    case ETL_K_NUMERIC: return parse_numeric(s, n, hc, o);
    case ETL_K_TIMESTAMPTZ:
      if (parse_timestamptz_fmt(s, n, true, o)) return 0;   // text.rs:111 %#z
      if (parse_timestamptz_fmt(s, n, false, o)) return 0;  // text.rs:113 %:z
      return ETL_E_DATETIME;
    case ETL_K_JSON:
      if (!json_valid(s, n)) return ETL_E_JSON;
      o.tag = ETL_CELL_JSON; o.val = soff; o.aux = n; return 0;
    case ETL_K_DATE: {
      Cur c{s, n}; int64_t days;
      if (!parse_date_part(c, &days) || c.n != 0) return ETL_E_DATETIME;
      o.tag = ETL_CELL_DATE; o.val = (uint64_t)days; return 0;
    }
    case ETL_K_TIME: {
      Cur c{s, n}; int64_t secs; uint32_t ns;
      if (!parse_time_part(c, &secs, &ns) || c.n != 0) return ETL_E_DATETIME;
      o.tag = ETL_CELL_TIME; o.val = (uint64_t)secs; o.aux = ns; return 0;
    }
    case ETL_K_TIMESTAMP: {
      Cur c{s, n}; int64_t days, secs; uint32_t ns;
      if (!parse_ts_prefix(c, &days, &secs, &ns) || c.n != 0) return ETL_E_DATETIME;
      o.tag = ETL_CELL_TIMESTAMP; o.val = (uint64_t)(days * 86400 + secs); o.aux = ns; return 0;
    }
    case ETL_K_UUID: return parse_uuid(s, n, hc, o);
    case ETL_K_BYTES: return parse_bytea(s, n, hc, o);
    default: return ETL_E_MALFORMED_FRAME;  // unsupported decode class: rejected on the host before launch
  }
}

// ---- walker: one thread follows the TupleData chain(s) of one DML frame (event.rs:376-919)
enum : uint32_t { W_OLD_HDR = 0, W_OLD_CELLS = 1, W_NEW_HDR = 2, W_NEW_CELLS = 3, W_DONE = 4 };
struct Walker {
  const uint8_t* base;   // frame start
  const uint8_t* p;      // next byte to read
  const uint8_t* end;    // frame end
  const uint8_t* kinds;
  const uint8_t* flags;
  uint64_t rec_index;    // global record index (error keys)
  uint32_t toff0;        // offset of the frame start relative to the tile start
  uint32_t n_cols, n_ident;
  uint32_t cell0;        // first output cell of the frame, relative to the tile's cell base
  int32_t remaining;     // wire cells left in the current tuple
  uint32_t wire_i, cmap, k_out, key_i, n_old;
  uint32_t kind, old_tag, stage;
  bool dense, partial, has_unresolved;
  bool emit;             // false after the first data error: keep checking structure only, because a
                         // malformed frame (parser error) outranks every conversion error of the record
};
__device__ __forceinline__ void put_cell(const DecodeParams& P, uint64_t idx, uint32_t tag, uint64_t val, uint32_t aux) {
  P.cell_tag[idx] = (uint8_t)tag; P.cell_val[idx] = val; P.cell_aux[idx] = aux;
}
#define W_DATA_ERROR(seq_, code_) do { report_error(P, w.rec_index, (seq_), (code_)); w.emit = false; } while (0)
#define W_MALFORMED() do { report_error(P, w.rec_index, SEQ_MALFORMED, ETL_E_MALFORMED_FRAME); w.stage = W_DONE; } while (0)

// advance the walker by at most `quota` text cells, writing descriptors to d[0..quota); returns count
__device__ __forceinline__ uint32_t walker_run(const DecodeParams& P, Walker& w, uint32_t slot, uint64_t tile_cell0,
                                               CellDesc* d, uint32_t quota, unsigned long long& tbytes) {
  uint32_t n = 0;
  while (w.stage != W_DONE && n < quota) {
    if (w.stage == W_OLD_HDR || w.stage == W_NEW_HDR) {
      const bool is_new = w.stage == W_NEW_HDR;
      if (is_new) {
        if (w.p >= w.end || *w.p != 'N') { W_MALFORMED(); break; }
        w.p++;
      }
      if (w.p + 2 > w.end) { W_MALFORMED(); break; }
      int32_t nc = (int32_t)(int16_t)be16(w.p);
      if (nc < 0) nc = 0;
      w.p += 2;
      w.remaining = nc; w.wire_i = 0; w.cmap = 0; w.k_out = 0;
      if (!is_new) {
        if (w.old_tag == 'K') {                     // normalize_key_tuple_to_row event.rs:879-919
          w.n_old = w.n_ident;
          w.dense = (uint32_t)nc == w.n_ident;
          if (w.emit) {
            if (w.n_ident == 0) W_DATA_ERROR(SEQ_OLD_SHAPE, ETL_E_KEY_NO_COLUMNS);
            else if (!w.dense && (uint32_t)nc != w.n_cols) W_DATA_ERROR(SEQ_OLD_SHAPE, ETL_E_KEY_SHAPE);
          }
        } else {                                    // convert_tuple_to_row event.rs:550-583
          w.n_old = w.n_cols;
          if (w.emit && (uint32_t)nc != w.n_cols) W_DATA_ERROR(SEQ_OLD_SHAPE, ETL_E_FIELD_COUNT);
        }
        w.stage = W_OLD_CELLS;
      } else {
        if (w.emit && (uint32_t)nc != w.n_cols) W_DATA_ERROR(SEQ_NEW_SHAPE, ETL_E_FIELD_COUNT);
        w.stage = W_NEW_CELLS;
      }
      continue;
    }
    if (w.remaining == 0) {
      w.stage = (w.stage == W_OLD_CELLS && w.kind != 'D') ? W_NEW_HDR : W_DONE;
      continue;
    }
    if (w.p >= w.end) { W_MALFORMED(); break; }
    const uint64_t x = ld64u(w.p);
    const uint32_t tag = (uint32_t)(x & 0xFFu);
    const uint32_t len = bswap32((uint32_t)(x >> 8));
    const uint32_t vrel = (uint32_t)(w.p - w.base) + 5u;   // value offset inside the frame
    const bool has_body = tag == 't' || tag == 'b';
    if (!has_body && tag != 'n' && tag != 'u') { W_MALFORMED(); break; }
    if (has_body) {
      if (w.p + 5 > w.end || (int32_t)len < 0 || (uint64_t)len > (uint64_t)(w.end - w.p - 5)) { W_MALFORMED(); break; }
      tbytes += len;
      w.p += 5 + (uint64_t)len;
    } else w.p += 1;
    const uint32_t i = w.wire_i++;
    w.remaining--;
    if (!w.emit) continue;                          // structure-only after a data error
    const bool is_new = w.stage == W_NEW_CELLS;
    uint32_t col = i, dest;
    if (!is_new && w.old_tag == 'K') {
      if (w.dense) { while (w.cmap < w.n_cols && !(w.flags[w.cmap] & 2)) w.cmap++; col = w.cmap++; }
      else if (!(w.flags[i] & 2)) continue;         // full-width key: non-identity entries are not decoded
      dest = w.cell0 + w.k_out++;
    } else dest = w.cell0 + (is_new ? w.n_old : 0u) + i;
    const uint32_t seq = is_new ? seq_new_cell(i) : seq_old_cell(i);
    const uint32_t cflags = w.flags[col];
    const bool resolver_key = is_new && w.kind == 'U' && w.old_tag == 'K' && (cflags & 2);
    if (tag == 't') {
      if (resolver_key) w.key_i++;
      CellDesc& cd = d[n++];
      cd.toff = w.toff0 + vrel;
      cd.len = len;
      cd.dest_rel = dest;
      cd.meta = (slot << 24) | ((uint32_t)w.kinds[col] << 16) | ((is_new ? 1u : 0u) << 15) | (i & 0x7FFFu);
      continue;
    }
    if (tag == 'n') {                               // convert_tuple_data_to_cell event.rs:941-957
      if (resolver_key) w.key_i++;
      if (cflags & 1) put_cell(P, tile_cell0 + dest, ETL_CELL_NULL, 0, 0);
      else W_DATA_ERROR(seq, ETL_E_NOT_NULL);
      continue;
    }
    if (tag == 'u') {                               // event.rs:958-970 + OldRowResolver :722-762
      if (is_new && w.kind == 'U') {
        uint32_t src = 0xFFFFFFFFu;
        if (w.old_tag == 'O') src = w.cell0 + i;
        else if (resolver_key) src = w.cell0 + w.key_i++;
        if (src != 0xFFFFFFFFu) { put_cell(P, tile_cell0 + dest, kCellUnresolved, tile_cell0 + src, 0); w.has_unresolved = true; }
        else { put_cell(P, tile_cell0 + dest, ETL_CELL_MISSING, 0, 0); w.partial = true; }
      } else W_DATA_ERROR(seq, (!is_new && w.old_tag == 'K') ? ETL_E_KEY_MISSING_VALUE : ETL_E_FULL_ROW_MISSING);
      continue;
    }
    if (resolver_key) w.key_i++;
    W_DATA_ERROR(seq, ETL_E_BINARY_FORMAT);         // 'b'
  }
  return n;
}

// UTF-8 validation of one long text cell by `nthreads` cooperating threads (a warp or the CTA):
// 16-byte aligned chunks, 4 independent loads in flight per thread; an all-ASCII chunk costs one
// load + one test; a chunk with high bits is checked with the position-local rule over [lo, hi+3)
// so the following chunk never has to look back.
__device__ __forceinline__ bool utf8_wide_bad(const uint8_t* ptr, uint32_t len, uint32_t t, uint32_t nthreads) {
  bool bad = false;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(ptr);
  const uint32_t headn = min(len, (uint32_t)((16u - (uint32_t)(a0 & 15u)) & 15u));
  if (t == 0 && headn) bad |= !utf8_chunk_valid(ptr, len, 0, min(headn + 3u, len));
  const uint32_t nchunks = (len - headn) / 16u;
  const uint4* body = reinterpret_cast<const uint4*>(ptr + headn);
  uint32_t c = t;
  for (; c + 3 * nthreads < nchunks; c += 4 * nthreads) {
    const uint4 x0 = body[c], x1 = body[c + nthreads], x2 = body[c + 2 * nthreads], x3 = body[c + 3 * nthreads];
    const uint32_t h0 = (x0.x | x0.y | x0.z | x0.w), h1 = (x1.x | x1.y | x1.z | x1.w), h2 = (x2.x | x2.y | x2.z | x2.w), h3 = (x3.x | x3.y | x3.z | x3.w);
    if ((h0 | h1 | h2 | h3) & 0x80808080u) {
      if (h0 & 0x80808080u) { const uint32_t lo = headn + c * 16u; bad |= !utf8_chunk_valid(ptr, len, lo, min(lo + 19u, len)); }
      if (h1 & 0x80808080u) { const uint32_t lo = headn + (c + nthreads) * 16u; bad |= !utf8_chunk_valid(ptr, len, lo, min(lo + 19u, len)); }
      if (h2 & 0x80808080u) { const uint32_t lo = headn + (c + 2 * nthreads) * 16u; bad |= !utf8_chunk_valid(ptr, len, lo, min(lo + 19u, len)); }
      if (h3 & 0x80808080u) { const uint32_t lo = headn + (c + 3 * nthreads) * 16u; bad |= !utf8_chunk_valid(ptr, len, lo, min(lo + 19u, len)); }
    }
  }
  for (; c < nchunks; c += nthreads) {
    const uint4 x = body[c];
    if ((x.x | x.y | x.z | x.w) & 0x80808080u) {
      const uint32_t lo = headn + c * 16u;
      bad |= !utf8_chunk_valid(ptr, len, lo, min(lo + 19u, len));
    }
  }
  const uint32_t tail0 = headn + nchunks * 16u;
  if (t == nthreads - 1 && tail0 < len) bad |= !utf8_chunk_valid(ptr, len, tail0, len);
  return bad;
}

__device__ __forceinline__ uint32_t warp_incl_sum(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
  return v;
}
__device__ __forceinline__ int32_t warp_incl_max(int32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int32_t u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, u); }
  return v;
}

__global__ void __launch_bounds__(kEmitThreads, 3) k_emit(DecodeParams P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  EmitShared& sh = *reinterpret_cast<EmitShared*>(smem_raw);
  const uint32_t spt = P.segs_per_tile;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int kWarps = kEmitThreads / 32;
  for (;;) {
    __syncthreads();  // previous tile fully consumed
    if (threadIdx.x == 0) sh.tile_idx = atomicAdd(P.tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = sh.tile_idx;
    if (tile >= P.n_tiles) break;
    const uint32_t seg0 = tile * spt;
    const uint32_t seg1 = min(seg0 + spt, P.n_anchors);
    const uint64_t t_begin = P.anchors[seg0];
    const uint64_t t_end = P.anchors[seg1];
    if (t_end <= t_begin) continue;
    // frames per segment → exclusive bases (warp 0)
    if (wid == 0) {
      uint32_t tot = 0;
      for (uint32_t s0 = 0; s0 < seg1 - seg0; s0 += 32) {
        const uint32_t si = s0 + lane;
        const uint32_t cnt = si < seg1 - seg0 ? P.seg_frames[seg0 + si] : 0u;
        const uint32_t inc = warp_incl_sum(cnt, lane);
        if (si < seg1 - seg0) sh.seg_base[si] = tot + inc - cnt;
        tot += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) {
        sh.seg_base[seg1 - seg0] = tot;
        sh.carry_cells = 0; sh.carry_cons = 0; sh.carry_b = -1; sh.carry_c = -1; sh.carry_b_lsn = 0; sh.carry_b_cons = 0;
        sh.metrics[0] = sh.metrics[1] = sh.metrics[2] = sh.metrics[3] = 0;
      }
    }
    const Summ tile_pre = fold(P.carry, P.tile_prefix[tile]);
    const uint64_t tile_cell0 = tile_pre.n_cells;
    const uint64_t tile_rec0 = tile_pre.n_rec;
    const bool in_tx0 = (tile_pre.flags & S_HAS_B) && !(tile_pre.flags & S_CLOSED);
    // ---- 1. stage the tile
    const uint64_t win0 = t_begin & ~15ull;
    const uint32_t lead = (uint32_t)(t_begin - win0);
    uint64_t wb = (t_end - win0 + 15ull) & ~15ull;
    const uint64_t to_end = ((uint64_t)(P.len - win0) + 15ull) & ~15ull;
    if (to_end < wb) wb = to_end;
    if ((uint64_t)kTileCap < wb) wb = (uint64_t)kTileCap;
    const uint32_t win_bytes = (uint32_t)wb;
    {
      const uint4* src = reinterpret_cast<const uint4*>(P.buf + win0);
      uint4* dst = reinterpret_cast<uint4*>(sh.tile);
      for (uint32_t i = threadIdx.x; i < win_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const uint32_t n_frames = min(sh.seg_base[seg1 - seg0], (uint32_t)kMaxTileFrames);
    // ---- 2. frame list
    if (threadIdx.x < seg1 - seg0) {
      uint64_t pos = P.anchors[seg0 + threadIdx.x];
      const uint64_t stop = P.anchors[seg0 + threadIdx.x + 1];
      uint32_t k = sh.seg_base[threadIdx.x];
      while (pos < stop && k < (uint32_t)kMaxTileFrames) {
        const uint32_t rel = (uint32_t)(pos - win0);
        uint32_t flen;
        if (rel + 8 <= win_bytes) {
          const uint64_t x = ld64u(sh.tile + rel);
          const uint64_t avail = P.len - pos;
          flen = bswap32((uint32_t)(x >> 8));
          if ((x & 0xFFu) != 'd' || avail < 5 || flen < 4 || 1ull + flen > avail) flen = (uint32_t)(avail > 0 ? avail - 1 : 0);
        } else flen = read_head(P.buf + pos, P.len - pos).flen;
        sh.foff[k++] = (uint16_t)(pos - t_begin);
        pos += 1ull + flen;
      }
    }
    __syncthreads();
    // ---- 3/4. chunks of 256 frames
    for (uint32_t c0 = 0; c0 < n_frames; c0 += blockDim.x) {
      const uint32_t f = c0 + threadIdx.x;
      const bool active = f < n_frames;
      FrameHead h;
      h.malformed = true; h.kind = 0; h.flen = 0; h.rel = 0; h.old_tag = 0;
      const uint8_t* fp = nullptr;
      uint64_t foff_abs = 0;
      uint32_t toff0 = 0;
      const DevSchema* s = nullptr;
      uint32_t my_cells = 0, my_cons = 0;
      uint64_t my_lsn = 0;
      bool isB = false, isC = false;
      if (active) {
        toff0 = sh.foff[f];
        foff_abs = t_begin + toff0;
        const uint32_t rel = lead + toff0;
        const uint8_t* gp = P.buf + foff_abs;
        h = read_head((rel + 40 <= win_bytes) ? sh.tile + rel : gp, P.len - foff_abs);
        fp = (rel + 1ull + h.flen + 16 <= win_bytes) ? sh.tile + rel : gp;
        if (!h.malformed) {
          if (h.kind == 'I' || h.kind == 'U' || h.kind == 'D') s = find_schema(P, h.rel, foff_abs);
          my_cells = frame_out_cells(h, s);
          isB = h.kind == 'B'; isC = h.kind == 'C';
          my_cons = (isB || isC || h.kind == 'R' || h.kind == 'I' || h.kind == 'U' || h.kind == 'D' || h.kind == 'T') ? 1u : 0u;
          if (isB) my_lsn = be64(fp + 31);
        }
      }
      if (threadIdx.x == 0) { sh.n_desc = 0; sh.n_wide = 0; sh.wide_next = 0; sh.n_big = 0; }
      // ---- four warp scans + cross-warp combine
      const uint32_t inc_cells = warp_incl_sum(my_cells, lane);
      const uint32_t inc_cons = warp_incl_sum(my_cons, lane);
      const int32_t inc_b = warp_incl_max(isB ? (int32_t)f : -1, lane);
      const int32_t inc_c = warp_incl_max(isC ? (int32_t)f : -1, lane);
      if (lane == 31) { sh.ws_cells[wid] = inc_cells; sh.ws_cons[wid] = inc_cons; sh.ws_b[wid] = inc_b; sh.ws_c[wid] = inc_c; }
      __syncthreads();
      uint32_t pre_cells = sh.carry_cells, pre_cons = sh.carry_cons;
      int32_t pre_b = sh.carry_b, pre_c = sh.carry_c;
      uint32_t tot_cells = pre_cells, tot_cons = pre_cons;
      int32_t tot_b = pre_b, tot_c = pre_c;
#pragma unroll
      for (int k = 0; k < kWarps; k++) {
        if (k < wid) { pre_cells += sh.ws_cells[k]; pre_cons += sh.ws_cons[k]; pre_b = max(pre_b, sh.ws_b[k]); pre_c = max(pre_c, sh.ws_c[k]); }
        tot_cells += sh.ws_cells[k]; tot_cons += sh.ws_cons[k]; tot_b = max(tot_b, sh.ws_b[k]); tot_c = max(tot_c, sh.ws_c[k]);
      }
      const uint32_t cells_excl = pre_cells + inc_cells - my_cells;   // tile-relative first output cell
      const uint32_t cons_incl = pre_cons + inc_cons;
      const uint32_t cons_excl = cons_incl - my_cons;
      int32_t eb = __shfl_up_sync(0xffffffffu, inc_b, 1), ec = __shfl_up_sync(0xffffffffu, inc_c, 1);
      if (lane == 0) { eb = -1; ec = -1; }
      const int32_t last_b = max(pre_b, eb), last_c = max(pre_c, ec);  // last Begin / Commit strictly before f
      if (isB) { sh.b_lsn[threadIdx.x] = my_lsn; sh.b_cons[threadIdx.x] = cons_incl; }
      const uint64_t old_carry_b_lsn = sh.carry_b_lsn;
      const uint32_t old_carry_b_cons = sh.carry_b_cons;
      __syncthreads();
      // stream state seen by this frame (apply.rs:600-626, 1927-2006)
      bool in_tx; uint64_t st_lsn; uint64_t st_ord;
      if (last_b >= 0) {
        const bool in_chunk = last_b >= (int32_t)c0;
        st_lsn = in_chunk ? sh.b_lsn[last_b - (int32_t)c0] : old_carry_b_lsn;
        const uint32_t bc = in_chunk ? sh.b_cons[last_b - (int32_t)c0] : old_carry_b_cons;
        st_ord = (uint64_t)(cons_excl - bc) + 1ull;
        in_tx = last_b > last_c;
      } else {
        st_lsn = tile_pre.lsn; st_ord = tile_pre.ord + cons_excl; in_tx = (last_c >= 0) ? false : in_tx0;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        sh.carry_cells = tot_cells; sh.carry_cons = tot_cons;
        if (tot_b >= (int32_t)c0) { sh.carry_b_lsn = sh.b_lsn[tot_b - (int32_t)c0]; sh.carry_b_cons = sh.b_cons[tot_b - (int32_t)c0]; }
        sh.carry_b = tot_b; sh.carry_c = tot_c;
      }
      Walker w;
      w.stage = W_DONE; w.has_unresolved = false; w.partial = false; w.n_cols = 0; w.n_old = 0; w.cell0 = 0;
      unsigned long long tb = 0;
      const uint64_t my_cell0 = tile_cell0 + cells_excl;
      if (active) {
        const uint64_t ridx = tile_rec0 + f;
        const uint64_t gidx = P.record_index_base + ridx;
        uint64_t commit_lsn = 0, ordinal = 0, start_lsn = 0;
        uint32_t rflags = 0;
        int32_t rschema = -1;
        uint32_t rrel = h.rel;
        bool ok = true;
        bool wellformed = !h.malformed;
        if (wellformed && h.kind == 'O') wellformed = (h.flen >= 4 + 26 + 8) && cstr_end(fp + 39, fp + 1 + h.flen) != nullptr;
        if (wellformed && h.kind == 'Y') {
          const uint8_t* fe = fp + 1 + h.flen;
          const uint8_t* q1 = (h.flen >= 4 + 26 + 4) ? cstr_end(fp + 35, fe) : nullptr;
          wellformed = q1 != nullptr && cstr_end(q1, fe) != nullptr;
        }
        if (wellformed && (h.kind == 'I' || h.kind == 'U' || h.kind == 'D')) {
          const uint32_t tt = fp[35];  // tuple marker (mlen >= 5 is guaranteed by read_head)
          if (h.kind == 'I') wellformed = tt == 'N';
          else if (h.kind == 'U') wellformed = tt == 'N' || tt == 'O' || tt == 'K';
          else wellformed = tt == 'O' || tt == 'K';
        }
        if (!wellformed) { report_error(P, gidx, SEQ_MALFORMED, ETL_E_MALFORMED_FRAME); ok = false; }
        else if (h.kind == 'k') { start_lsn = be64(fp + 6); rrel = fp[22]; }
        else {
          start_lsn = bswap64(ld64u(fp + 6));          // wal_start apply.rs:1700
          const uint8_t* m = fp + 31;                  // message body after the tag
          switch (h.kind) {
            case 'B':                                  // apply.rs:1927-1943
              commit_lsn = my_lsn; ordinal = 0; rflags = ETL_RF_EVENT;
              put_cell(P, my_cell0, ETL_CELL_I64, be64(m + 8), 0);
              put_cell(P, my_cell0 + 1, ETL_CELL_U32, be32(m + 16), 0);
              break;
            case 'C': {                                // apply.rs:1946-2006
              if (!in_tx) { report_error(P, gidx, SEQ_STATE, ETL_E_TX_STATE); ok = false; break; }
              const uint64_t cl = be64(m + 1);
              if (cl != st_lsn) { report_error(P, gidx, SEQ_STATE, ETL_E_COMMIT_LSN); ok = false; break; }
              commit_lsn = cl; ordinal = st_ord; rflags = ETL_RF_EVENT;
              put_cell(P, my_cell0, ETL_CELL_I32, (uint64_t)(int64_t)(int8_t)m[0], 0);
              put_cell(P, my_cell0 + 1, ETL_CELL_I64, be64(m + 9), 0);
              put_cell(P, my_cell0 + 2, ETL_CELL_I64, be64(m + 17), 0);
              break;
            }
            case 'R':                                  // apply.rs:2012-2089 (masks are built on the host)
              for (uint32_t k = 0; k < P.n_rel_errors; k++)
                if (P.rel_error_off[k] == foff_abs) { report_error(P, gidx, P.rel_error_seq[k], P.rel_error_code[k]); ok = false; }
              if (!in_tx) { report_error(P, gidx, SEQ_STATE, ETL_E_TX_STATE); ok = false; break; }
              commit_lsn = st_lsn; ordinal = st_ord; rflags = ETL_RF_EVENT;
              { const DevSchema* rs = find_schema(P, h.rel, foff_abs); if (rs && rs->effective_off == foff_abs) rschema = (int32_t)rs->batch_index; }
              break;
            case 'I': case 'U': case 'D': {            // apply.rs:2092-2203
              // the tuple structure is validated by the walker (a malformed frame outranks state errors)
              if (!in_tx) report_error(P, gidx, SEQ_STATE, ETL_E_TX_STATE);
              commit_lsn = st_lsn; ordinal = st_ord;
              if (!s) {                                // no schema to walk with: structure check only
                unsigned long long ignored;
                if (!dml_structure_ok(h, fp, &ignored)) report_error(P, gidx, SEQ_MALFORMED, ETL_E_MALFORMED_FRAME);
                report_error(P, gidx, SEQ_TABLE, ETL_E_MISSING_TABLE_STATE); ok = false; break;
              }
              rschema = (int32_t)s->batch_index; rflags = ETL_RF_EVENT;
              if (h.old_tag == 'O') rflags |= ETL_RF_OLD_FULL; else if (h.old_tag == 'K') rflags |= ETL_RF_OLD_KEY;
              break;
            }
            case 'T': {                                // apply.rs:2206-2248
              if (!in_tx) { report_error(P, gidx, SEQ_STATE, ETL_E_TX_STATE); ok = false; break; }
              commit_lsn = st_lsn; ordinal = st_ord;
              put_cell(P, my_cell0, ETL_CELL_I32, (uint64_t)(int64_t)(int8_t)m[4], 0);
              for (uint32_t i = 0; i < h.rel; i++) {
                const uint32_t rid = be32(m + 5 + 4 * i);
                const DevSchema* ts = find_schema(P, rid, foff_abs);
                if (!ts) { report_error(P, gidx, SEQ_TABLE, ETL_E_MISSING_TABLE_STATE); ok = false; break; }
                put_cell(P, my_cell0 + 1 + i, ETL_CELL_U32, rid, ts->batch_index);
              }
              if (h.rel > 0) rflags = ETL_RF_EVENT;
              break;
            }
            case 'M': {                                // apply.rs:1808-1924
              const uint8_t* end = fp + 1 + h.flen;
              const uint8_t* q = m + 9;
              const char* ddl = "supabase_etl_ddl";
              bool is_ddl = true; uint32_t k = 0; bool term = false;
              if (q > end) { report_error(P, gidx, SEQ_MALFORMED, ETL_E_MALFORMED_FRAME); ok = false; break; }
              for (; q + k < end; k++) { uint32_t ch = q[k]; if (!ch) { term = true; break; } if (k >= 16 || ch != (uint32_t)(uint8_t)ddl[k]) is_ddl = false; }
              if (!term || !utf8_valid(q, k)) { report_error(P, gidx, SEQ_MALFORMED, ETL_E_MALFORMED_FRAME); ok = false; break; }
              is_ddl = is_ddl && k == 16;
              const uint8_t* cq = q + k + 1;
              if (cq + 4 > end) { report_error(P, gidx, SEQ_MALFORMED, ETL_E_MALFORMED_FRAME); ok = false; break; }
              const int32_t cl = (int32_t)be32(cq);
              if (cl < 0 || (uint64_t)cl > (uint64_t)(end - cq - 4)) { report_error(P, gidx, SEQ_MALFORMED, ETL_E_MALFORMED_FRAME); ok = false; break; }
              if (is_ddl) { rflags |= ETL_RF_DDL_MESSAGE; if (!in_tx) { report_error(P, gidx, SEQ_STATE, ETL_E_TX_STATE); ok = false; } }
              break;
            }
            default: break;                            // Origin / Type: structure only
          }
        }
        P.rec_off[ridx] = foff_abs; P.rec_kind[ridx] = (uint8_t)h.kind; P.rec_flags[ridx] = (uint8_t)rflags;
        P.rec_rel[ridx] = rrel; P.rec_schema[ridx] = rschema; P.rec_start_lsn[ridx] = start_lsn;
        P.rec_commit_lsn[ridx] = commit_lsn; P.rec_tx_ordinal[ridx] = ordinal; P.rec_cell_base[ridx] = my_cell0;
        if (ok && (rflags & ETL_RF_EVENT)) atomicAdd(&sh.metrics[3], 1ull);
        sh.fi_base[threadIdx.x] = fp; sh.fi_off[threadIdx.x] = foff_abs; sh.fi_rec[threadIdx.x] = (uint32_t)ridx;
        if (ok && s && (h.kind == 'I' || h.kind == 'U' || h.kind == 'D')) {
          w.base = fp; w.end = fp + 1 + h.flen;
          w.kinds = P.col_kind + s->col_base; w.flags = P.col_flags + s->col_base;
          w.rec_index = gidx; w.toff0 = toff0; w.n_cols = s->n_cols; w.n_ident = s->n_ident;
          w.cell0 = cells_excl; w.emit = true;
          w.remaining = 0; w.wire_i = 0; w.cmap = 0; w.k_out = 0; w.key_i = 0; w.n_old = 0;
          w.kind = h.kind; w.old_tag = h.old_tag; w.dense = false;
          if (h.kind != 'I' && h.old_tag) { w.stage = W_OLD_HDR; w.p = fp + 36; }  // old image first
          else { w.stage = W_NEW_HDR; w.p = fp + 35; }                             // 'N' marker, then the new tuple
        }
      }
      // ---- walker rounds
      const uint32_t n_walkers = __syncthreads_count(w.stage != W_DONE);
      if (n_walkers) {
        uint32_t quota = (uint32_t)kDescCap / n_walkers;
        if (quota > 255) quota = 255;
        const uint32_t win_rel = win_bytes - lead;    // tile-relative end of the shared window
        for (;;) {
          if (threadIdx.x < 32) sh.hist[threadIdx.x] = 0;
          if (threadIdx.x == 0) { sh.round_heap_need = 0; sh.round_heap_used = 0; }
          if (w.stage != W_DONE) {
            const uint32_t slot0 = atomicAdd(&sh.n_desc, quota);
            const uint32_t got = walker_run(P, w, threadIdx.x, tile_cell0, sh.desc + slot0, quota, tb);
            for (uint32_t k = got; k < quota; k++) sh.desc[slot0 + k].meta = 0xFFFFFFFFu;
          }
          __syncthreads();
          const uint32_t nd = sh.n_desc;
          // ---- counting sort of the batch by decode class (+ heap need of the round)
          uint32_t rank[kDescPerThread];
          uint32_t need = 0;
#pragma unroll
          for (int k = 0; k < kDescPerThread; k++) {
            const uint32_t di = threadIdx.x + k * kEmitThreads;
            rank[k] = 0xFFFFFFFFu;
            if (di < nd) {
              const uint32_t meta = sh.desc[di].meta;
              if (meta != 0xFFFFFFFFu) {
                const uint32_t kind = (meta >> 16) & 31u;
                rank[k] = atomicAdd(&sh.hist[kind], 1u);
                need += cell_heap_bound(kind, sh.desc[di].len);
              }
            }
          }
          if (P.heap_cap && need) atomicAdd(&sh.round_heap_need, need);
          __syncthreads();
          if (threadIdx.x < 32) {
            const uint32_t cnt = sh.hist[threadIdx.x];
            const uint32_t inc = warp_incl_sum(cnt, lane);
            sh.kstart[threadIdx.x] = inc - cnt;
            if (threadIdx.x == 31) sh.n_sorted = inc;
            if (threadIdx.x == 0 && sh.round_heap_need) sh.round_heap_base = atomicAdd(P.heap_top, (unsigned long long)sh.round_heap_need);
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < kDescPerThread; k++) {
            const uint32_t di = threadIdx.x + k * kEmitThreads;
            if (rank[k] != 0xFFFFFFFFu) sh.perm[sh.kstart[(sh.desc[di].meta >> 16) & 31u] + rank[k]] = (uint16_t)di;
          }
          __syncthreads();
          const uint32_t ns = sh.n_sorted;
          const unsigned long long round_heap = sh.round_heap_base;
          // ---- thread-per-cell parsing, kind-sorted
          for (uint32_t j = threadIdx.x; j < ns; j += blockDim.x) {
            const CellDesc cd = sh.desc[sh.perm[j]];
            const uint32_t slot = cd.meta >> 24, kind = (cd.meta >> 16) & 0xFFu, wi = cd.meta & 0x7FFFu;
            const uint32_t seq = ((cd.meta >> 15) & 1u) ? seq_new_cell(wi) : seq_old_cell(wi);
            const uint32_t len = cd.len;
            const uint8_t* v = ((uint64_t)cd.toff + len + 16 <= win_rel) ? sh.tile + lead + cd.toff : P.buf + t_begin + cd.toff;
            const uint64_t soff = t_begin + cd.toff;
            CellOut o;
            uint32_t code = 0;
            if (kind == ETL_K_STRING && len >= (uint32_t)kWideLen) {
              bool queued = false;
              if (len >= (uint32_t)kBigLen) {
                const uint32_t bs = atomicAdd(&sh.n_big, 1u);
                if (bs < (uint32_t)kBigCap) { sh.big[bs] = WideText{v, len, seq, sh.fi_rec[slot]}; queued = true; }
              }
              if (!queued) {
                const uint32_t ws = atomicAdd(&sh.n_wide, 1u);
                if (ws < (uint32_t)kWideCap) { sh.wide[ws] = WideText{v, len, seq, sh.fi_rec[slot]}; queued = true; }
              }
              if (!queued && !utf8_valid(v, len)) code = ETL_E_UTF8;
              o.tag = ETL_CELL_STRING; o.val = soff; o.aux = len;
            } else if (!utf8_valid_fast(v, len)) code = ETL_E_UTF8;         // event.rs:972
            else {
              const uint32_t hb = cell_heap_bound(kind, len);
              HeapCursor hc{P.heap, 0};
              if (hb) hc.pos = round_heap + atomicAdd(&sh.round_heap_used, hb);
              code = parse_text_cell(kind, v, len, soff, hc, o);
            }
            if (code) report_error(P, P.record_index_base + sh.fi_rec[slot], seq, code);
            else put_cell(P, tile_cell0 + cd.dest_rel, o.tag, o.val, o.aux);
          }
          __syncthreads();
          // ---- long text: the whole CTA per very large cell, then one warp per cell
          const uint32_t nbig = min(sh.n_big, (uint32_t)kBigCap);
          for (uint32_t b = 0; b < nbig; b++) {
            const WideText wt = sh.big[b];
            const bool bad = utf8_wide_bad(wt.ptr, wt.len, threadIdx.x, blockDim.x);
            if (__syncthreads_or(bad) && threadIdx.x == 0) report_error(P, P.record_index_base + wt.rec_local, wt.seq, ETL_E_UTF8);
          }
          const uint32_t nw = min(sh.n_wide, (uint32_t)kWideCap);
          for (;;) {
            uint32_t wi = 0;
            if (lane == 0) wi = atomicAdd(&sh.wide_next, 1u);
            wi = __shfl_sync(0xffffffffu, wi, 0);
            if (wi >= nw) break;
            const WideText wt = sh.wide[wi];
            const bool bad = utf8_wide_bad(wt.ptr, wt.len, (uint32_t)lane, 32u);
            if (__any_sync(0xffffffffu, bad) && lane == 0) report_error(P, P.record_index_base + wt.rec_local, wt.seq, ETL_E_UTF8);
          }
          __syncthreads();
          if (threadIdx.x == 0) { sh.n_desc = 0; sh.n_wide = 0; sh.wide_next = 0; sh.n_big = 0; }
          if (!__syncthreads_or(w.stage != W_DONE)) break;
        }
      }
      // ---- per-frame epilogue: unchanged-TOAST values copied from the old image; Partial flag; metrics
      if (active && s && (h.kind == 'I' || h.kind == 'U' || h.kind == 'D')) {
        if (w.has_unresolved) {
          const uint64_t c0n = tile_cell0 + w.cell0 + w.n_old;
          for (uint32_t i = 0; i < w.n_cols; i++)
            if (P.cell_tag[c0n + i] == kCellUnresolved) {
              const uint64_t src = P.cell_val[c0n + i];
              put_cell(P, c0n + i, P.cell_tag[src], P.cell_val[src], P.cell_aux[src]);
            }
        }
        if (w.partial) P.rec_flags[tile_rec0 + f] |= ETL_RF_NEW_PARTIAL;
        if (tb) atomicAdd(&sh.metrics[h.kind == 'I' ? 0 : (h.kind == 'U' ? 1 : 2)], tb);
      }
      __syncthreads();
    }
    if (threadIdx.x < 4 && sh.metrics[threadIdx.x]) atomicAdd(&P.metrics[threadIdx.x], sh.metrics[threadIdx.x]);
  }
}
#undef W_DATA_ERROR
#undef W_MALFORMED

}  // namespace etl
