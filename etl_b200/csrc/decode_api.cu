// decode_api.cu — extern "C" ABI (include/etl_decode.h) of the B200 pgoutput decode engine.
//
// Host side only does what the reference does per batch / per Relation message (rare, control
// path): staging, schema catalogue, Relation → ReplicationMask / IdentityMask
// (apply.rs:2012-2089, event.rs:325-369, etl-postgres/src/types/schema.rs:288-323,406-438,527-535),
// buffer management and kernel orchestration.  Every per-row / per-cell operation of the hot path
// runs in the sm_100a kernels of wal_kernels.cuh; there is no CPU decode fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "etl_decode.h"
#include "wal_kernels.cuh"

using namespace etl;

namespace {

struct StoredCol {
  std::string name;
  uint32_t type_oid;
  int32_t modifier, ordinal, pk;
  uint8_t nullable;
};
struct StoredTable {
  uint64_t snapshot_id = 0;
  std::vector<StoredCol> cols;
};
struct RelVersion {  // ReplicatedTableSchema (schema.rs:651-900)
  uint32_t table_id = 0;
  uint64_t snapshot_id = 0;
  uint64_t effective_off = 0;
  uint32_t n_ident = 0;
  std::vector<uint8_t> kind, flags;
  std::vector<int32_t> index;
};

// text.rs:28-173 + utils.rs:7-16: type oid → decode class
uint32_t kind_for_oid(uint32_t oid) {
  switch (oid) {
    case 16: return ETL_K_BOOL;
    case 1000: return ETL_K_ARRAY | ETL_K_BOOL;
    case 18: case 1042: case 1043: case 19: case 25: case 790: return ETL_K_STRING;
    case 1002: case 1014: case 1015: case 1003: case 1009: case 791: return ETL_K_ARRAY | ETL_K_STRING;
    case 21: return ETL_K_I16;
    case 1005: return ETL_K_ARRAY | ETL_K_I16;
    case 23: return ETL_K_I32;
    case 1007: return ETL_K_ARRAY | ETL_K_I32;
    case 20: return ETL_K_I64;
    case 1016: return ETL_K_ARRAY | ETL_K_I64;
    case 700: return ETL_K_F32;
    case 1021: return ETL_K_ARRAY | ETL_K_F32;
    case 701: return ETL_K_F64;
    case 1022: return ETL_K_ARRAY | ETL_K_F64;
    case 1700: return ETL_K_NUMERIC;
    case 1231: return ETL_K_ARRAY | ETL_K_NUMERIC;
    case 17: return ETL_K_BYTES;
    case 1001: return ETL_K_ARRAY | ETL_K_BYTES;
    case 1082: return ETL_K_DATE;
    case 1182: return ETL_K_ARRAY | ETL_K_DATE;
    case 1083: return ETL_K_TIME;
    case 1183: return ETL_K_ARRAY | ETL_K_TIME;
    case 1114: return ETL_K_TIMESTAMP;
    case 1115: return ETL_K_ARRAY | ETL_K_TIMESTAMP;
    case 1184: return ETL_K_TIMESTAMPTZ;
    case 1185: return ETL_K_ARRAY | ETL_K_TIMESTAMPTZ;
    case 2950: return ETL_K_UUID;
    case 2951: return ETL_K_ARRAY | ETL_K_UUID;
    case 114: case 3802: return ETL_K_JSON;
    case 199: case 3807: return ETL_K_ARRAY | ETL_K_JSON;
    case 26: return ETL_K_U32;
    case 1028: return ETL_K_ARRAY | ETL_K_U32;
    default: break;
  }
  static const uint32_t other_arrays[] = {
      143, 271, 629, 651, 719, 775, 1006, 1008, 1010, 1011, 1012, 1013, 1017, 1018, 1019, 1020, 1027, 1034, 1040,
      1041, 1187, 1263, 1270, 1561, 1563, 2201, 2207, 2208, 2209, 2210, 2211, 2949, 3221, 3643, 3644, 3645, 3735,
      3770, 3905, 3907, 3909, 3911, 3913, 3927, 4073, 4090, 4097, 4192, 5039, 6150, 6151, 6152, 6153, 6155, 6157};
  for (uint32_t a : other_arrays)
    if (a == oid) return ETL_K_ARRAY | ETL_K_STRING;
  return ETL_K_STRING;
}
bool kind_supported_on_device(uint32_t k) {
  if (k & ETL_K_ARRAY) k &= ~(uint32_t)ETL_K_ARRAY;   // arrays: element kinds below
  switch (k) {
    case ETL_K_BOOL: case ETL_K_STRING: case ETL_K_I16: case ETL_K_I32: case ETL_K_U32: case ETL_K_I64:
    case ETL_K_NUMERIC: case ETL_K_DATE: case ETL_K_TIME: case ETL_K_TIMESTAMP: case ETL_K_TIMESTAMPTZ:
    case ETL_K_UUID: case ETL_K_JSON: case ETL_K_BYTES: case ETL_K_F32: case ETL_K_F64:
      return true;
    default: return false;
  }
}
bool kind_has_heap(uint32_t k) { return k == ETL_K_NUMERIC || k == ETL_K_UUID || k == ETL_K_BYTES || (k & ETL_K_ARRAY); }

uint32_t error_kind_of(uint32_t code) {
  switch (code) {
    case ETL_E_NONE: return ETL_EK_NONE;
    case ETL_E_UUID: case ETL_E_BOOL: case ETL_E_NOT_NULL: return ETL_EK_INVALID_DATA;
    case ETL_E_JSON: return ETL_EK_DESERIALIZATION_ERROR;
    case ETL_E_TX_STATE: case ETL_E_MISSING_TABLE_STATE: return ETL_EK_INVALID_STATE;
    case ETL_E_COMMIT_LSN: return ETL_EK_VALIDATION_ERROR;
    case ETL_E_UNKNOWN_COLUMNS: return ETL_EK_CORRUPTED_TABLE_SCHEMA;
    case ETL_E_MISSING_TABLE_SCHEMA: return ETL_EK_MISSING_TABLE_SCHEMA;
    case ETL_E_MALFORMED_FRAME: return ETL_EK_SOURCE_ERROR;
    default: return ETL_EK_CONVERSION_ERROR;
  }
}

uint32_t rd32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint16_t rd16(const uint8_t* p) { return (uint16_t)((p[0] << 8) | p[1]); }

bool utf8_ok(const uint8_t* s, size_t n) {
  size_t i = 0;
  while (i < n) {
    uint8_t b = s[i];
    if (b < 0x80) { i++; continue; }
    if (b >= 0xC2 && b <= 0xDF) { if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return false; i += 2; }
    else if (b >= 0xE0 && b <= 0xEF) {
      if (i + 2 >= n) return false;
      uint8_t lo = b == 0xE0 ? 0xA0 : 0x80, hi = b == 0xED ? 0x9F : 0xBF;
      if (s[i + 1] < lo || s[i + 1] > hi || (s[i + 2] & 0xC0) != 0x80) return false;
      i += 3;
    } else if (b >= 0xF0 && b <= 0xF4) {
      if (i + 3 >= n) return false;
      uint8_t lo = b == 0xF0 ? 0x90 : 0x80, hi = b == 0xF4 ? 0x8F : 0xBF;
      if (s[i + 1] < lo || s[i + 1] > hi || (s[i + 2] & 0xC0) != 0x80 || (s[i + 3] & 0xC0) != 0x80) return false;
      i += 4;
    } else return false;
  }
  return true;
}

template <typename T>
struct DevBuf {  // growable device scratch
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = std::max<size_t>(n + n / 4, 256);
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

// ================================================================================================
struct etl_stager {
  uint8_t* buf = nullptr;   // pinned
  uint64_t cap = 0, len = 0;
  uint32_t stride = 2048;
  std::vector<uint64_t> anchors;
  std::vector<uint64_t> relations;
  size_t n_real_anchors = 0;   // anchors.size() before etl_stage_view padded the tail with `len`
  bool padded = false;
};

struct etl_dec_batch {
  etl_dec_ctx* ctx = nullptr;
  etl_dec_planes dev{};
  etl_dec_planes host{};
  bool has_host = false;
  void* dev_block = nullptr;   // single device allocation holding all planes
  void* host_block = nullptr;  // pinned host copy (borrowed from ctx->h_result)
  size_t block_bytes = 0;
  etl_dec_summary summary{};
  std::vector<RelVersion> schemas;
};

struct etl_dec_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool lines_launched = false;
  std::string last_error;
  std::map<uint32_t, StoredTable> tables;
  std::map<uint32_t, RelVersion> current;  // SharedTableCache Ready state (table_cache.rs:36-130)
  // scratch
  DevBuf<uint8_t> d_stream;
  DevBuf<uint64_t> d_anchors;
  DevBuf<uint32_t> d_seg_frames, d_act, d_act_blk;
  DevBuf<Summ> d_tile_summ, d_group_summ, d_group_prefix, d_total, d_tile_prefix, d_seg_summ;
  DevBuf<uint32_t> d_schema_by_batch;
  DevBuf<DevSchema> d_schemas;
  DevBuf<uint8_t> d_col_kind, d_col_flags;
  DevBuf<uint32_t> d_line_bad, d_dead, d_bin_count, d_bin_cursor, d_perm, d_bin_start, d_bin_row_base, d_row_chunk;
  DevBuf<CellDesc> d_desc;
  DevBuf<CopyPair> d_copies;
  DevBuf<LongCell> d_long;
  DevBuf<unsigned long long> d_scalars;  // [0] first_error key, [1..4] metrics
  DevBuf<uint64_t> d_rel_err_off;
  DevBuf<uint32_t> d_rel_err_code, d_rel_err_seq;
  void* h_result = nullptr; size_t h_result_cap = 0;  // pinned result staging, grow-only
  uint64_t pending_h2d_bytes = 0;
  Summ* h_total = nullptr;               // pinned
  unsigned long long* h_scalars = nullptr;  // pinned
  cudaEvent_t ev[6]{};
  cudaEvent_t evk[3]{};
  cudaStream_t side = nullptr;           // k_utf8_lines runs here, underneath the index / records passes
  cudaEvent_t ev_in = nullptr, ev_l0 = nullptr, ev_l1 = nullptr;
  // pending two-phase decode
  bool pending = false;
  DecodeParams P{};
  std::vector<RelVersion> pending_schemas;
  uint32_t pending_flags = 0;
  float pending_h2d_ms = 0, pending_index_ms = 0;
  uint32_t launches = 0;
  const uint8_t* pending_host_buf = nullptr;
};

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ctx->last_error = std::string(#call) + ": " + cudaGetErrorString(_e);              \
      return ETL_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

extern "C" {

uint32_t etl_dec_abi_version(void) { return ETL_DECODE_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------ stager
int etl_stage_create(uint64_t capacity_bytes, uint32_t anchor_stride, etl_stager** out) {
  if (!out || anchor_stride < 256 || anchor_stride > 32768 || (anchor_stride & (anchor_stride - 1))) return ETL_ERR_INVALID_ARG;
  etl_stager* s = new etl_stager();
  s->stride = anchor_stride;
  s->cap = capacity_bytes;
  // pinned when a CUDA device is usable, plain memory otherwise (the stager itself needs no GPU)
  if (cudaHostAlloc((void**)&s->buf, capacity_bytes ? capacity_bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    s->buf = (uint8_t*)malloc(capacity_bytes ? capacity_bytes : 1);
    if (!s->buf) { delete s; return ETL_ERR_ALLOC; }
    s->cap |= (1ull << 63);  // tag: malloc'ed
  }
  *out = s;
  return ETL_OK;
}
void etl_stage_destroy(etl_stager* s) {
  if (!s) return;
  if (s->cap >> 63) free(s->buf); else cudaFreeHost(s->buf);
  delete s;
}
void etl_stage_reset(etl_stager* s) { s->len = 0; s->anchors.clear(); s->relations.clear(); s->padded = false; s->n_real_anchors = 0; }

static inline void stage_note_frame(etl_stager* s, uint64_t off, const uint8_t* body, uint32_t body_len) {
  // anchors[k] = first frame starting at or after k*stride
  if (s->padded) { s->anchors.resize(s->n_real_anchors); s->padded = false; }
  while ((uint64_t)s->anchors.size() * s->stride <= off) s->anchors.push_back(off);
  if (body_len >= 26 && body[0] == 'w' && body[25] == 'R') s->relations.push_back(off);
}
int etl_stage_append(etl_stager* s, const uint8_t* body, uint32_t body_len) {
  uint64_t cap = s->cap & ~(1ull << 63);
  if (s->len + 5ull + body_len > cap) return ETL_ERR_ALLOC;
  uint64_t off = s->len;
  uint8_t* d = s->buf + off;
  uint32_t fl = body_len + 4;
  d[0] = 'd'; d[1] = (uint8_t)(fl >> 24); d[2] = (uint8_t)(fl >> 16); d[3] = (uint8_t)(fl >> 8); d[4] = (uint8_t)fl;
  memcpy(d + 5, body, body_len);
  stage_note_frame(s, off, body, body_len);
  s->len += 5ull + body_len;
  return ETL_OK;
}
int etl_stage_append_framed(etl_stager* s, const uint8_t* framed, uint64_t len) {
  uint64_t cap = s->cap & ~(1ull << 63);
  if (s->len + len > cap) return ETL_ERR_ALLOC;
  uint64_t base = s->len;
  memcpy(s->buf + base, framed, len);
  uint64_t pos = 0;
  while (pos + 5 <= len) {
    if (framed[pos] != 'd') break;
    uint32_t fl = rd32(framed + pos + 1);
    if (fl < 4 || pos + 1ull + fl > len) break;
    stage_note_frame(s, base + pos, framed + pos + 5, fl - 4);
    pos += 1ull + fl;
  }
  s->len += len;
  return pos == len ? ETL_OK : ETL_ERR_INVALID_ARG;
}
int etl_stage_view(const etl_stager* cs, etl_dec_input* out) {
  etl_stager* s = const_cast<etl_stager*>(cs);
  // blocks k*stride past the last frame start have no frame: anchors[k] = len
  if (!s->padded) { s->n_real_anchors = s->anchors.size(); s->padded = true; }
  s->anchors.resize(s->n_real_anchors);
  const uint64_t want = s->len ? (s->len + s->stride - 1) / s->stride : 0;
  while (s->anchors.size() < want) s->anchors.push_back(s->len);
  memset(out, 0, sizeof *out);
  out->host_buf = s->buf;
  out->len = s->len;
  out->anchors = s->anchors.data();
  out->n_anchors = s->anchors.size();
  out->anchor_stride = s->stride;
  out->relation_offsets = s->relations.data();
  out->n_relations = s->relations.size();
  return ETL_OK;
}

// ------------------------------------------------------------------------------------------------ ctx
int etl_dec_create(int device_id, etl_dec_ctx** out) {
  if (!out) return ETL_ERR_INVALID_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return ETL_ERR_NO_DEVICE; }
  if (device_id < 0 || device_id >= n) return ETL_ERR_INVALID_ARG;
  etl_dec_ctx* ctx = new etl_dec_ctx();
  ctx->device = device_id;
  if (cudaSetDevice(device_id) != cudaSuccess) { delete ctx; return ETL_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ETL_ERR_CUDA; }
  ctx->own_stream = true;
  for (auto& e : ctx->ev) cudaEventCreate(&e);
  for (auto& e : ctx->evk) cudaEventCreate(&e);
  if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ETL_ERR_CUDA; }
  {  // batch planes come from the stream-ordered pool: keep freed blocks for the next batch instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming); cudaEventCreate(&ctx->ev_l0); cudaEventCreate(&ctx->ev_l1);
  cudaHostAlloc((void**)&ctx->h_total, sizeof(Summ), cudaHostAllocDefault);
  cudaHostAlloc((void**)&ctx->h_scalars, 16 * sizeof(unsigned long long), cudaHostAllocDefault);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = ctx;
  return ETL_OK;
}
int etl_dec_set_stream(etl_dec_ctx* ctx, void* s) {
  if (!ctx) return ETL_ERR_INVALID_ARG;
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return ETL_OK;
}
void etl_dec_destroy(etl_dec_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->d_stream.release(); ctx->d_anchors.release(); ctx->d_seg_frames.release(); ctx->d_tile_summ.release();
  ctx->d_group_summ.release(); ctx->d_group_prefix.release(); ctx->d_tile_prefix.release(); ctx->d_line_bad.release(); ctx->d_seg_summ.release(); ctx->d_schema_by_batch.release(); ctx->d_total.release(); ctx->d_schemas.release();
  ctx->d_col_kind.release(); ctx->d_col_flags.release(); ctx->d_scalars.release(); ctx->d_rel_err_off.release();
  ctx->d_rel_err_code.release(); ctx->d_rel_err_seq.release();
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->h_total) cudaFreeHost(ctx->h_total);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->evk) if (e) cudaEventDestroy(e);
  if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
  if (ctx->ev_l0) cudaEventDestroy(ctx->ev_l0);
  if (ctx->ev_l1) cudaEventDestroy(ctx->ev_l1);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
const char* etl_dec_last_error(const etl_dec_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null ctx"; }

int etl_dec_put_table_schema(etl_dec_ctx* ctx, uint32_t table_id, uint64_t snapshot_id, const etl_column_schema* cols,
                             uint32_t n_cols) {
  if (!ctx || (n_cols && !cols)) return ETL_ERR_INVALID_ARG;
  StoredTable t;
  t.snapshot_id = snapshot_id;
  for (uint32_t i = 0; i < n_cols; i++) {
    StoredCol c;
    c.name = cols[i].name ? cols[i].name : "";
    c.type_oid = cols[i].type_oid; c.modifier = cols[i].modifier; c.ordinal = cols[i].ordinal_position;
    c.pk = cols[i].primary_key_ordinal_position; c.nullable = cols[i].nullable;
    t.cols.push_back(std::move(c));
  }
  ctx->tables[table_id] = std::move(t);
  return ETL_OK;
}
int etl_dec_reset_relations(etl_dec_ctx* ctx) {
  if (!ctx) return ETL_ERR_INVALID_ARG;
  ctx->current.clear();
  return ETL_OK;
}

// handle_relation_message (apply.rs:2012-2089): Relation body → masks → ReplicatedTableSchema.
// returns 0 ok, else etl_error_code; *seq receives the error step.
static uint32_t build_relation(etl_dec_ctx* ctx, const uint8_t* frame, uint64_t avail, uint64_t off, RelVersion* out,
                               uint32_t* seq) {
  *seq = 0;
  if (avail < 5 || frame[0] != 'd') return ETL_E_MALFORMED_FRAME;
  uint32_t fl = rd32(frame + 1);
  if (fl < 4 || 1ull + fl > avail) return ETL_E_MALFORMED_FRAME;
  const uint8_t* end = frame + 1 + fl;
  const uint8_t* p = frame + 31;  // after 'd' len 'w' hdr 'R'
  if (p + 4 > end) return ETL_E_MALFORMED_FRAME;
  uint32_t rel_id = rd32(p); p += 4;
  auto cstr = [&](const uint8_t** q) -> const uint8_t* {
    const uint8_t* s = *q;
    const uint8_t* z = (const uint8_t*)memchr(s, 0, (size_t)(end - s));
    if (!z) return nullptr;
    *q = z + 1;
    return s;
  };
  if (!cstr(&p) || !cstr(&p)) return ETL_E_MALFORMED_FRAME;
  if (p + 3 > end) return ETL_E_MALFORMED_FRAME;
  uint8_t replident = *p++;
  if (replident != 'd' && replident != 'n' && replident != 'f' && replident != 'i') return ETL_E_MALFORMED_FRAME;
  int16_t ncols = (int16_t)rd16(p); p += 2;
  struct RC { std::string name; uint8_t flags; };
  std::vector<RC> rcols;
  for (int i = 0; i < ncols; i++) {
    if (p + 1 > end) return ETL_E_MALFORMED_FRAME;
    uint8_t flags = *p++;
    const uint8_t* nm = cstr(&p);
    if (!nm) return ETL_E_MALFORMED_FRAME;
    size_t nl = (size_t)(p - 1 - nm);
    if (!utf8_ok(nm, nl)) return ETL_E_MALFORMED_FRAME;
    if (p + 8 > end) return ETL_E_MALFORMED_FRAME;
    p += 8;
    rcols.push_back(RC{std::string((const char*)nm, nl), flags});
  }
  *seq = 2;
  auto it = ctx->tables.find(rel_id);
  if (it == ctx->tables.end()) return ETL_E_MISSING_TABLE_SCHEMA;
  const StoredTable& t = it->second;
  std::vector<uint8_t> repl(t.cols.size(), 0), ident(t.cols.size(), 0);
  for (const RC& rc : rcols) {
    bool found = false;
    for (size_t k = 0; k < t.cols.size(); k++)
      if (t.cols[k].name == rc.name) {
        found = true; repl[k] = 1;
        if (replident == 'f' || (rc.flags & 1)) ident[k] = 1;  // event.rs:351-366
      }
    if (!found) return ETL_E_UNKNOWN_COLUMNS;                  // schema.rs:288-309
  }
  out->table_id = rel_id; out->snapshot_id = t.snapshot_id; out->effective_off = off; out->n_ident = 0;
  out->kind.clear(); out->flags.clear(); out->index.clear();
  for (size_t k = 0; k < t.cols.size(); k++) {
    if (!repl[k]) continue;
    out->kind.push_back((uint8_t)kind_for_oid(t.cols[k].type_oid));
    out->flags.push_back((uint8_t)((t.cols[k].nullable ? 1 : 0) | (ident[k] ? 2 : 0)));
    out->index.push_back((int32_t)k);
    if (ident[k]) out->n_ident++;
  }
  return 0;
}

static void free_batch_blocks(etl_dec_batch* b) {
  if (!b) return;
  cudaStream_t st = b->ctx ? b->ctx->stream : nullptr;
  if (b->dev_block) cudaFreeAsync(b->dev_block, st);
  b->dev_block = nullptr; b->host_block = nullptr;
}
void etl_dec_batch_free(etl_dec_batch* b) {
  if (!b) return;
  free_batch_blocks(b);
  delete b;
}

// Where the structure-blind UTF-8 pass (k_utf8_dead, HBM-bound) runs relative to the latency-bound passes.
// 0: side stream from the start of the index pass; 1: side stream from the start of the tuple passes
// (the index + record passes keep the memory system to themselves); 2: main stream after the tuple passes.
// ETL_DEAD_MODE / ETL_DEAD_CTAS are tuning knobs for measurement, not part of the ABI.
static int dead_mode() {
  static const int m = getenv("ETL_DEAD_SERIAL") ? 2 : (getenv("ETL_DEAD_MODE") ? atoi(getenv("ETL_DEAD_MODE")) : 1);
  return m;
}
static int dead_ctas(int dflt) {
  static const int c = getenv("ETL_DEAD_CTAS") ? atoi(getenv("ETL_DEAD_CTAS")) : 0;
  return c > 0 ? c : dflt;
}
static cudaError_t launch_dead_side(etl_dec_ctx* ctx, cudaStream_t st) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  cudaError_t e;
  if ((e = cudaEventRecord(ctx->ev_in, st)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(ctx->side, ctx->ev_in, 0)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(ctx->ev_l0, ctx->side)) != cudaSuccess) return e;
  k_utf8_dead<<<sms * dead_ctas(3), 256, 0, ctx->side>>>(ctx->P);
  if ((e = cudaEventRecord(ctx->ev_l1, ctx->side)) != cudaSuccess) return e;
  ctx->launches += 1;
  ctx->lines_launched = true;
  return cudaSuccess;
}

int etl_dec_decode_begin(etl_dec_ctx* ctx, const etl_dec_input* in, uint32_t flags, etl_dec_seam* seam_out) {
  if (!ctx || !in) return ETL_ERR_INVALID_ARG;
  ctx->pending = false;
  const uint32_t stride = in->anchor_stride;
  if (stride < 256 || stride > 32768 || (stride & (stride - 1))) { ctx->last_error = "anchor_stride must be a power of two in [256, 32768]"; return ETL_ERR_INVALID_ARG; }
  if (in->len && (!in->host_buf && !in->dev_buf)) { ctx->last_error = "no input buffer"; return ETL_ERR_INVALID_ARG; }
  if (in->dev_buf && (reinterpret_cast<uintptr_t>(in->dev_buf) & 15u)) { ctx->last_error = "dev_buf must be 16-byte aligned"; return ETL_ERR_INVALID_ARG; }
  const uint64_t n_anchors_expected = in->len ? (in->len + stride - 1) / stride : 0;
  if (in->n_anchors != n_anchors_expected || (in->n_anchors && !in->anchors && !in->dev_anchors)) { ctx->last_error = "anchors: expected ceil(len/stride) entries"; return ETL_ERR_INVALID_ARG; }
  if (in->n_relations && (!in->relation_offsets || !in->host_buf)) { ctx->last_error = "relation_offsets require host_buf"; return ETL_ERR_INVALID_ARG; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ctx->launches = 0;

  // ---- schema versions of this batch: carried-in (ascending table id) then Relation frames in order
  std::vector<RelVersion> vers;
  for (auto& kv : ctx->current) { RelVersion v = kv.second; v.effective_off = 0; vers.push_back(std::move(v)); }
  std::vector<uint64_t> rel_err_off; std::vector<uint32_t> rel_err_code, rel_err_seq;
  for (uint64_t i = 0; i < in->n_relations; i++) {
    uint64_t off = in->relation_offsets[i];
    if (off >= in->len) { ctx->last_error = "relation offset out of range"; return ETL_ERR_INVALID_ARG; }
    RelVersion v; uint32_t seq = 0;
    uint32_t code = build_relation(ctx, in->host_buf + off, in->len - off, off, &v, &seq);
    if (code) { rel_err_off.push_back(off); rel_err_code.push_back(code); rel_err_seq.push_back(seq); continue; }
    ctx->current[v.table_id] = v;  // note_ready apply.rs:2079
    vers.push_back(std::move(v));
  }
  for (const RelVersion& v : vers)
    for (uint8_t k : v.kind)
      if (!kind_supported_on_device(k)) {
        char msg[160];
        snprintf(msg, sizeof msg, "table %u: column decode class 0x%x (array types) has no device parser yet; refusing to decode", v.table_id, k);
        ctx->last_error = msg;
        return ETL_ERR_INVALID_ARG;
      }
  // device tables sorted by (table_id, effective_off)
  std::vector<uint32_t> order(vers.size());
  for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    if (vers[a].table_id != vers[b].table_id) return vers[a].table_id < vers[b].table_id;
    return vers[a].effective_off < vers[b].effective_off;
  });
  std::vector<DevSchema> ds; std::vector<uint8_t> ck, cf;
  for (uint32_t oi : order) {
    const RelVersion& v = vers[oi];
    DevSchema d{};
    d.table_id = v.table_id; d.n_cols = (uint32_t)v.kind.size(); d.n_ident = v.n_ident; d.col_base = (uint32_t)ck.size();
    d.effective_off = v.effective_off; d.batch_index = oi; d.has_heap = 0;
    for (uint8_t k : v.kind) if (kind_has_heap(k)) d.has_heap = 1;
    ck.insert(ck.end(), v.kind.begin(), v.kind.end());
    cf.insert(cf.end(), v.flags.begin(), v.flags.end());
    ds.push_back(d);
  }

  // ---- geometry
  DecodeParams& P = ctx->P;
  memset(&P, 0, sizeof P);
  P.len = in->len;
  P.n_anchors = (uint32_t)in->n_anchors;
  P.anchor_stride = stride;
  P.segs_per_tile = 32;  // one warp of anchor segments per tile
  P.n_tiles = (P.n_anchors + P.segs_per_tile - 1) / P.segs_per_tile;
  P.tiles_per_group = std::max<uint32_t>(1, kIndexThreads / P.segs_per_tile);
  P.n_groups = (P.n_tiles + P.tiles_per_group - 1) / P.tiles_per_group;

  // ---- uploads
  CK(cudaEventRecord(ctx->ev[0], st));
  ctx->pending_h2d_bytes = (in->dev_buf ? 0 : in->len) + (in->dev_anchors ? 0 : (in->n_anchors + 1) * 8) +
                           ds.size() * sizeof(DevSchema) + ck.size() * 2 + rel_err_off.size() * 16 + 5 * 8;
  if (in->dev_buf) P.buf = in->dev_buf;
  else {
    CK(ctx->d_stream.ensure(in->len + 64));
    if (in->len) CK(cudaMemcpyAsync(ctx->d_stream.p, in->host_buf, in->len, cudaMemcpyHostToDevice, st));
    P.buf = ctx->d_stream.p;
  }
  if (in->dev_anchors) P.anchors = in->dev_anchors;
  else {
    CK(ctx->d_anchors.ensure(in->n_anchors + 1));
    if (in->n_anchors) CK(cudaMemcpyAsync(ctx->d_anchors.p, in->anchors, in->n_anchors * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_anchors.p + in->n_anchors, &in->len, 8, cudaMemcpyHostToDevice, st));
    P.anchors = ctx->d_anchors.p;
  }
  CK(ctx->d_schemas.ensure(ds.size() + 1)); CK(ctx->d_col_kind.ensure(ck.size() + 1)); CK(ctx->d_col_flags.ensure(cf.size() + 1));
  if (!ds.empty()) CK(cudaMemcpyAsync(ctx->d_schemas.p, ds.data(), ds.size() * sizeof(DevSchema), cudaMemcpyHostToDevice, st));
  if (!ck.empty()) {
    CK(cudaMemcpyAsync(ctx->d_col_kind.p, ck.data(), ck.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_col_flags.p, cf.data(), cf.size(), cudaMemcpyHostToDevice, st));
  }
  P.schemas = ctx->d_schemas.p; P.n_schemas = (uint32_t)ds.size(); P.col_kind = ctx->d_col_kind.p; P.col_flags = ctx->d_col_flags.p;
  CK(ctx->d_rel_err_off.ensure(rel_err_off.size() + 1)); CK(ctx->d_rel_err_code.ensure(rel_err_off.size() + 1)); CK(ctx->d_rel_err_seq.ensure(rel_err_off.size() + 1));
  if (!rel_err_off.empty()) {
    CK(cudaMemcpyAsync(ctx->d_rel_err_off.p, rel_err_off.data(), rel_err_off.size() * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_rel_err_code.p, rel_err_code.data(), rel_err_code.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_rel_err_seq.p, rel_err_seq.data(), rel_err_seq.size() * 4, cudaMemcpyHostToDevice, st));
  }
  P.rel_error_off = ctx->d_rel_err_off.p; P.rel_error_code = ctx->d_rel_err_code.p; P.rel_error_seq = ctx->d_rel_err_seq.p;
  P.n_rel_errors = (uint32_t)rel_err_off.size();
  CK(ctx->d_seg_frames.ensure(P.n_anchors + 1)); CK(ctx->d_tile_summ.ensure(P.n_tiles + 1));
  CK(ctx->d_group_summ.ensure(P.n_groups + 1)); CK(ctx->d_group_prefix.ensure(P.n_groups + 1)); CK(ctx->d_total.ensure(1));
  CK(ctx->d_scalars.ensure(16)); CK(ctx->d_tile_prefix.ensure(P.n_tiles + 1));
  CK(ctx->d_seg_summ.ensure(P.n_anchors + 1)); P.seg_summ = ctx->d_seg_summ.p;
  {
    std::vector<uint32_t> sbb(ds.size() + 1, 0);
    for (uint32_t i = 0; i < ds.size(); i++) sbb[ds[i].batch_index] = i;
    CK(ctx->d_schema_by_batch.ensure(sbb.size()));
    CK(cudaMemcpyAsync(ctx->d_schema_by_batch.p, sbb.data(), sbb.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // sbb is a local
    P.schema_by_batch = ctx->d_schema_by_batch.p;
  }
  P.tile_prefix = ctx->d_tile_prefix.p;
  const size_t line_words = (in->len + 4095) / 4096 + 1;
  CK(ctx->d_line_bad.ensure(line_words)); CK(ctx->d_dead.ensure(P.n_anchors + 1));
  P.line_bad = ctx->d_line_bad.p; P.dead = ctx->d_dead.p;
  P.long_cap = P.n_anchors + 16;                      // a listed cell covers at least one whole segment
  CK(ctx->d_long.ensure(P.long_cap));
  P.long_cells = ctx->d_long.p; P.long_count = (unsigned int*)(ctx->d_scalars.p + 7);
  CK(cudaMemsetAsync(P.line_bad, 0, line_words * 4, st));
  P.seg_frames = ctx->d_seg_frames.p; P.tile_summ = ctx->d_tile_summ.p; P.group_summ = ctx->d_group_summ.p;
  P.group_prefix = ctx->d_group_prefix.p; P.total = ctx->d_total.p;
  P.first_error = ctx->d_scalars.p; P.metrics = ctx->d_scalars.p + 1;
  const uint32_t act_blocks = (P.n_anchors + kActThreads - 1) / kActThreads;
  CK(ctx->d_act.ensure(P.n_anchors + 1)); CK(ctx->d_act_blk.ensure(act_blocks + 1));
  P.act = ctx->d_act.p; P.act_blk = ctx->d_act_blk.p; P.n_act = (unsigned int*)(ctx->d_scalars.p + 12);   // [12] survives decode_finish's reset of [0..10]
  CK(cudaEventRecord(ctx->ev[1], st));

  ctx->lines_launched = false;

  // ---- pass A + B
  if (P.n_groups) {
    k_act_count<<<act_blocks, kActThreads, 0, st>>>(P);
    k_act_scan<<<1, kActThreads, 0, st>>>(P, act_blocks);
    k_act_scatter<<<act_blocks, kActThreads, 0, st>>>(P);
    if (dead_mode() == 0) CK(launch_dead_side(ctx, st));   // underneath everything that follows
    k_index<<<P.n_groups, P.tiles_per_group * P.segs_per_tile, 0, st>>>(P);
    k_scan<<<1, 512, 0, st>>>(P);
    k_tile_prefix<<<(P.n_tiles + 255) / 256, 256, 0, st>>>(P);
    ctx->launches += 6;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_total, P.total, sizeof(Summ), cudaMemcpyDeviceToHost, st));
  } else *ctx->h_total = summ_identity();
  CK(cudaEventRecord(ctx->ev[2], st));
  CK(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->pending_h2d_ms, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->pending_index_ms, ctx->ev[1], ctx->ev[2]);

  const Summ& T = *ctx->h_total;
  if (seam_out) {
    memset(seam_out, 0, sizeof *seam_out);
    seam_out->n_records = T.n_rec; seam_out->n_cells = T.n_cells; seam_out->heap_bytes = 0;
    seam_out->lsn = T.lsn; seam_out->ord = T.ord;
    seam_out->has_begin = (T.flags & S_HAS_B) ? 1 : 0; seam_out->closed = (T.flags & S_CLOSED) ? 1 : 0;
  }
  ctx->pending = true;
  ctx->pending_schemas = std::move(vers);
  ctx->pending_flags = flags;
  ctx->pending_host_buf = in->host_buf;
  return ETL_OK;
}

int etl_dec_decode_finish(etl_dec_ctx* ctx, const etl_stream_state* carry_in, uint64_t record_index_base,
                          etl_dec_batch** out) {
  if (!ctx || !out || !ctx->pending) { if (ctx) ctx->last_error = "decode_finish without decode_begin"; return ETL_ERR_INVALID_ARG; }
  ctx->pending = false;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DecodeParams& P = ctx->P;
  const Summ T = *ctx->h_total;
  etl_dec_batch* b = new etl_dec_batch();
  b->ctx = ctx;
  b->schemas = std::move(ctx->pending_schemas);

  // ---- one device block for all planes
  bool any_heap = false, any_array = false;
  for (const RelVersion& v : b->schemas) for (uint8_t k : v.kind) { any_heap = any_heap || kind_has_heap(k); any_array = any_array || (k & ETL_K_ARRAY); }
  // upper bound on Σ cell_heap_bound: numeric ≤ n/2+19, bytea ≤ n/2+7, uuid = 16 per decoded text cell.
  // Arrays reserve 16 + 44·n_elems + 1.5·len per cell: first guess 3·len, retried ×4 on overflow (≤ 48·len).
  const uint64_t nr = T.n_rec, nc = T.n_cells;
  uint64_t nh = any_heap ? (P.len / 2 + 24 * T.n_cells + 256) : 0;
  nh = (nh + 15) & ~15ull;
  const uint64_t scalar_heap = nh;
  if (any_array) nh += 3 * P.len + 4096;
  auto al = [](uint64_t x) { return (x + 255) & ~255ull; };
  uint64_t f_rec_off, f_kind, f_flags, f_rel, f_schema, f_start, f_commit, f_ord, f_cbase, f_tag, f_val, f_aux, f_heap;
  auto fill = [&](etl_dec_planes& pl, uint8_t* bs) {
    pl.n_records = nr; pl.n_cells = nc; pl.heap_bytes = nh;
    pl.rec_off = (uint64_t*)(bs + f_rec_off); pl.rec_kind = bs + f_kind; pl.rec_flags = bs + f_flags;
    pl.rec_rel = (uint32_t*)(bs + f_rel); pl.rec_schema = (int32_t*)(bs + f_schema); pl.rec_start_lsn = (uint64_t*)(bs + f_start);
    pl.rec_commit_lsn = (uint64_t*)(bs + f_commit); pl.rec_tx_ordinal = (uint64_t*)(bs + f_ord); pl.rec_cell_base = (uint64_t*)(bs + f_cbase);
    pl.cell_tag = bs + f_tag; pl.cell_val = (uint64_t*)(bs + f_val); pl.cell_aux = (uint32_t*)(bs + f_aux); pl.heap = bs + f_heap;
  };
  P.record_index_base = record_index_base;
  Summ carry = summ_identity();
  etl_stream_state cin{};
  if (carry_in) cin = *carry_in;
  if (cin.in_tx) { carry.flags = S_HAS_B; carry.lsn = cin.final_lsn; }
  carry.ord = cin.next_tx_ordinal;
  P.carry = carry;
  uint64_t heap_used = 0;
  for (int attempt = 0;; attempt++) {
    uint64_t cur = 0;
    auto take = [&](uint64_t bytes) { uint64_t o = cur; cur += al(bytes); return o; };
    f_rec_off = take(nr * 8); f_kind = take(nr); f_flags = take(nr); f_rel = take(nr * 4); f_schema = take(nr * 4);
    f_start = take(nr * 8); f_commit = take(nr * 8); f_ord = take(nr * 8); f_cbase = take((nr + 1) * 8);
    f_tag = take(nc); f_val = take(nc * 8); f_aux = take(nc * 4); f_heap = take(nh);
    b->block_bytes = cur ? cur : 256;
    CK(cudaMallocAsync(&b->dev_block, b->block_bytes, st));
    fill(b->dev, (uint8_t*)b->dev_block);
    P.rec_off = (uint64_t*)b->dev.rec_off; P.rec_kind = (uint8_t*)b->dev.rec_kind; P.rec_flags = (uint8_t*)b->dev.rec_flags;
    P.rec_rel = (uint32_t*)b->dev.rec_rel; P.rec_schema = (int32_t*)b->dev.rec_schema; P.rec_start_lsn = (uint64_t*)b->dev.rec_start_lsn;
    P.rec_commit_lsn = (uint64_t*)b->dev.rec_commit_lsn; P.rec_tx_ordinal = (uint64_t*)b->dev.rec_tx_ordinal;
    P.rec_cell_base = (uint64_t*)b->dev.rec_cell_base; P.cell_tag = (uint8_t*)b->dev.cell_tag; P.cell_val = (uint64_t*)b->dev.cell_val;
    P.cell_aux = (uint32_t*)b->dev.cell_aux; P.heap = (uint8_t*)b->dev.heap;
    P.heap_top = ctx->d_scalars.p + 5; P.heap_cap = nh;
    P.heap_overflow = (unsigned int*)(ctx->d_scalars.p + 9);
    P.arr_top = ctx->d_scalars.p + 10; P.arr_base = scalar_heap;

    // ---- pass C
    ctx->h_scalars[0] = ~0ull;
    for (int i = 1; i < 12; i++) ctx->h_scalars[i] = 0;
    CK(cudaMemcpyAsync(ctx->d_scalars.p, ctx->h_scalars, 12 * 8, cudaMemcpyHostToDevice, st));
    P.n_bins = (uint32_t)std::min<size_t>(kMaxBins, std::max<size_t>(16, 16 * b->schemas.size()));
    const size_t perm_cap = nr + 32ull * P.n_bins + 256;
    CK(ctx->d_bin_count.ensure(kMaxBins)); CK(ctx->d_bin_cursor.ensure(kMaxBins)); CK(ctx->d_perm.ensure(perm_cap));
    P.bin_count = ctx->d_bin_count.p; P.bin_cursor = ctx->d_bin_cursor.p; P.perm = ctx->d_perm.p;
    P.perm_len = (unsigned int*)(ctx->d_scalars.p + 11);
    CK(ctx->d_bin_start.ensure(kMaxBins)); CK(ctx->d_bin_row_base.ensure(kMaxBins));
    P.bin_start = ctx->d_bin_start.p; P.bin_row_base = ctx->d_bin_row_base.p; P.n_batch_schemas = (uint32_t)b->schemas.size();
    // descriptor rows: Σ over bins ceil(count/32)·slots(bin) ≤ slots/32 + Σ slots(bin) ≤ slots/32 + 32·Σ n_cols
    uint64_t sum_cols = 0, max_cols = 0;
    for (const RelVersion& v : b->schemas) { sum_cols += v.kind.size(); max_cols = std::max<uint64_t>(max_cols, v.kind.size()); }
    uint64_t rows_cap = T.slots / 32 + 32 * sum_cols + 64;
    if (16 * b->schemas.size() > (size_t)kMaxBins) rows_cap = ((uint64_t)nr * 2 * max_cols) / 32 + 64ull * kMaxBins / 16 * max_cols + 64;   // clamped bins take the widest schema
    if (rows_cap >= (1ull << 31)) { ctx->last_error = "batch too large for the descriptor plane"; delete b; return ETL_ERR_INVALID_ARG; }
    P.desc_row_cap = (uint32_t)rows_cap;
    CK(ctx->d_desc.ensure(rows_cap * 32)); CK(ctx->d_row_chunk.ensure(rows_cap));
    P.desc = ctx->d_desc.p; P.row_chunk = ctx->d_row_chunk.p; P.desc_rows = (unsigned int*)(ctx->d_scalars.p + 13);
    P.copy_cap = (uint32_t)std::min<uint64_t>(nc / 2 + 16, 0xFFFFFFFFull);
    CK(ctx->d_copies.ensure(P.copy_cap));
    P.copies = ctx->d_copies.p; P.copy_count = (unsigned int*)(ctx->d_scalars.p + 8);
    CK(cudaMemsetAsync(P.bin_count, 0, P.n_bins * 4, st));
    CK(cudaMemsetAsync(P.perm, 0xFF, perm_cap * 4, st));
    CK(cudaEventRecord(ctx->ev[3], st));
    if (P.n_tiles) {
      P.n_records = nr;
      k_frames<<<(P.n_anchors + 255) / 256, 256, 0, st>>>(P);
      cudaEventRecord(ctx->evk[0], st);
      if (dead_mode() == 1 && !ctx->lines_launched) CK(launch_dead_side(ctx, st));   // underneath the tuple passes only
      if (nr) {
        k_bin_scan<<<1, 1024, 0, st>>>(P);
        k_perm<<<(uint32_t)((nr + 255) / 256), 256, 0, st>>>(P);
        k_walk<<<(uint32_t)((nr + 32ull * P.n_bins + kWalkThreads - 1) / kWalkThreads) + 64u, kWalkThreads, 0, st>>>(P);
        cudaEventRecord(ctx->evk[2], st);
        k_cells<<<(uint32_t)((rows_cap * 32 + 255) / 256), 256, 0, st>>>(P);
        k_copy<<<128, 256, 0, st>>>(P);
      } else cudaEventRecord(ctx->evk[2], st);
      cudaEventRecord(ctx->evk[1], st);
      if (ctx->lines_launched) CK(cudaStreamWaitEvent(st, ctx->ev_l1, 0));   // join: the bitmap is complete
      else {                                          // ETL_DEAD_SERIAL: the same pass on the main stream (tuning knob)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        cudaEventRecord(ctx->ev_l0, st);
        k_utf8_dead<<<sms * dead_ctas(6), 256, 0, st>>>(P);
        cudaEventRecord(ctx->ev_l1, st);
        ctx->launches += 1;
      }
      if (nr) k_long_verdict<<<592, 256, 0, st>>>(P);
      ctx->launches += nr ? 7 : 1;
      CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync((void*)(b->dev.rec_cell_base + nr), &ctx->h_total->n_cells, 8, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[4], st));
    CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars.p, 13 * 8, cudaMemcpyDeviceToHost, st));
    heap_used = nh;
    if (!nh) break;
    CK(cudaStreamSynchronize(st));
    heap_used = std::min<uint64_t>(nh, ctx->h_scalars[10] ? scalar_heap + ctx->h_scalars[10] : ctx->h_scalars[5]);
    if (!ctx->h_scalars[9]) break;
    if (attempt >= 3) { ctx->last_error = "array heap reservation overflow after retries"; cudaFreeAsync(b->dev_block, st); delete b; return ETL_ERR_CUDA; }
    CK(cudaFreeAsync(b->dev_block, st));
    b->dev_block = nullptr;
    nh = scalar_heap + (nh - scalar_heap) * 4;
  }
  const uint64_t copy_bytes = f_heap + heap_used;
  b->dev.heap_bytes = heap_used;
  if (ctx->pending_flags & ETL_DECODE_RESULTS_TO_HOST) {
    if (ctx->h_result_cap < b->block_bytes) {
      if (ctx->h_result) cudaFreeHost(ctx->h_result);
      ctx->h_result = nullptr; ctx->h_result_cap = 0;
      size_t want = b->block_bytes + b->block_bytes / 8;
      CK(cudaHostAlloc(&ctx->h_result, want, cudaHostAllocDefault));
      ctx->h_result_cap = want;
    }
    b->host_block = ctx->h_result;
    CK(cudaMemcpyAsync(b->host_block, b->dev_block, copy_bytes, cudaMemcpyDeviceToHost, st));
    fill(b->host, (uint8_t*)b->host_block);
    b->host.heap_bytes = heap_used;
    b->has_host = true;
  }
  CK(cudaEventRecord(ctx->ev[5], st));
  CK(cudaStreamSynchronize(st));

  // ---- summary
  etl_dec_summary& S = b->summary;
  memset(&S, 0, sizeof S);
  float emit_ms = 0, d2h_ms = 0;
  cudaEventElapsedTime(&emit_ms, ctx->ev[3], ctx->ev[4]);
  cudaEventElapsedTime(&d2h_ms, ctx->ev[4], ctx->ev[5]);
  S.kernel_ms = ctx->pending_index_ms + emit_ms;
  S.index_ms = ctx->pending_index_ms; S.emit_ms = emit_ms;
  if (P.n_tiles) {
    cudaEventElapsedTime(&S.frames_ms, ctx->ev[3], ctx->evk[0]);
    cudaEventElapsedTime(&S.walk_ms, ctx->evk[0], ctx->evk[2]);    // k_bin_scan + k_perm + k_walk (structure)
    cudaEventElapsedTime(&S.cells_ms, ctx->evk[2], ctx->evk[1]);   // k_cells + k_copy
    cudaEventElapsedTime(&S.spans_ms, ctx->ev_l0, ctx->ev_l1);   // concurrent with index / records
  }
  S.h2d_ms = ctx->pending_h2d_ms; S.d2h_ms = d2h_ms;
  S.h2d_bytes = ctx->pending_h2d_bytes;
  // bytes k_utf8_dead streamed: the dead segments (h_scalars[12] = live segment count, left by k_act_scan)
  S.span_bytes = P.n_tiles ? std::min<uint64_t>(P.len, (uint64_t)(P.n_anchors - (uint32_t)ctx->h_scalars[12]) * P.anchor_stride) : 0;
  S.d2h_bytes = 5 * 8 + sizeof(Summ) + ((ctx->pending_flags & ETL_DECODE_RESULTS_TO_HOST) ? copy_bytes : 0);
  S.gpu_launches = ctx->launches;
  S.n_schemas = (uint32_t)b->schemas.size();
  unsigned long long key = ctx->h_scalars[0];
  if (key == ~0ull) { S.first_error.record_index = UINT64_MAX; }
  else {
    S.first_error.record_index = (key >> 24) - 0;  // global index
    S.first_error.seq = (uint32_t)((key >> 6) & 0x3FFFFu);
    S.first_error.code = (uint32_t)(key & 63u);
    S.first_error.kind = error_kind_of(S.first_error.code);
  }
  S.insert_bytes = ctx->h_scalars[1]; S.update_bytes = ctx->h_scalars[2]; S.delete_bytes = ctx->h_scalars[3]; S.n_events = ctx->h_scalars[4];
  Summ endst = fold(carry, T);
  S.carry_out.in_tx = ((endst.flags & S_HAS_B) && !(endst.flags & S_CLOSED)) ? 1 : 0;
  S.carry_out.final_lsn = endst.lsn;
  S.carry_out.next_tx_ordinal = endst.ord;
  *out = b;
  return ETL_OK;
}

int etl_dec_decode(etl_dec_ctx* ctx, const etl_dec_input* in, uint32_t flags, etl_dec_batch** out) {
  int rc = etl_dec_decode_begin(ctx, in, flags, nullptr);
  if (rc) return rc;
  return etl_dec_decode_finish(ctx, &in->carry_in, 0, out);
}

int etl_dec_batch_planes(const etl_dec_batch* b, int host, etl_dec_planes* out) {
  if (!b || !out) return ETL_ERR_INVALID_ARG;
  if (host && !b->has_host) return ETL_ERR_INVALID_ARG;
  *out = host ? b->host : b->dev;
  return ETL_OK;
}
int etl_dec_batch_summary(const etl_dec_batch* b, etl_dec_summary* out) {
  if (!b || !out) return ETL_ERR_INVALID_ARG;
  *out = b->summary;
  return ETL_OK;
}
int etl_dec_batch_schema(const etl_dec_batch* b, uint32_t i, etl_dec_schema_info* out) {
  if (!b || !out || i >= b->schemas.size()) return ETL_ERR_INVALID_ARG;
  const RelVersion& v = b->schemas[i];
  out->table_id = v.table_id; out->n_cols = (uint32_t)v.kind.size(); out->n_identity = v.n_ident; out->_pad = 0;
  out->snapshot_id = v.snapshot_id; out->effective_off = v.effective_off;
  out->col_kind = v.kind.data(); out->col_flags = v.flags.data(); out->col_index = v.index.data();
  return ETL_OK;
}

}  // extern "C"
