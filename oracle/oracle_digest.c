/* oracle_digest.c — canonical digest of a decoded batch (the plane layout of include/etl_decode.h).
 * TEST INFRASTRUCTURE ONLY (see oracle.h): bench.py's post-timing parity leg and tests/ use it to compare the
 * CUDA path with the oracle at the BASELINE sizes, where the per-cell Python comparator of tests/canon.py is
 * too slow.  Same canonical form as canon.assert_planes_equal: record fields and fixed-width cells by value,
 * var-width payloads (numeric / uuid / bytes / arrays) dereferenced from the heap so that heap placement does
 * not matter; string and json cells are (stream offset, length) spans of the same staged stream.
 * Four independent 64-bit multiplicative hashes (256 bits) over the canonical byte sequence. */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

typedef struct dg { uint64_t h[4]; } dg;
static const uint64_t K[4] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0xD6E8FEB86659FD93ull};
static inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline void dg_u64(dg* d, uint64_t v) {
  for (int i = 0; i < 4; i++) { uint64_t x = (d->h[i] ^ v) * K[i]; d->h[i] = rotl(x, 23 + 4 * i) + K[(i + 1) & 3]; }
}
static void dg_bytes(dg* d, const uint8_t* p, uint64_t n) {
  dg_u64(d, n);
  uint64_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t v; memcpy(&v, p + i, 8); dg_u64(d, v); }
  if (i < n) { uint64_t v = 0; memcpy(&v, p + i, n - i); dg_u64(d, v); }
}
static void dg_cell(dg* d, uint32_t tag, uint64_t val, uint32_t aux, const uint8_t* heap, int in_array) {
  dg_u64(d, tag);
  switch (tag) {
    case ETL_CELL_NUMERIC: {
      etl_numeric_hdr h; memcpy(&h, heap + val, sizeof h);
      dg_u64(d, h.kind | ((uint64_t)h.sign << 8) | ((uint64_t)(uint16_t)h.weight << 16) | ((uint64_t)h.scale << 32) | ((uint64_t)(in_array ? 0 : h.pushed_groups) << 48));
      dg_bytes(d, heap + val + 8, (uint64_t)aux * 2);
      break;
    }
    case ETL_CELL_UUID: dg_bytes(d, heap + val, 16); break;
    case ETL_CELL_BYTES: dg_bytes(d, heap + val, aux); break;
    case ETL_CELL_ARRAY: {
      etl_array_hdr ah; memcpy(&ah, heap + val, sizeof ah);
      dg_u64(d, ah.elem_kind | ((uint64_t)ah.n_elems << 8));
      for (uint32_t k = 0; k < ah.n_elems; k++) {
        etl_array_elem e; memcpy(&e, heap + val + 8 + 16ull * k, sizeof e);
        if (e.tag == ETL_CELL_STRING || e.tag == ETL_CELL_JSON) { dg_u64(d, e.tag); dg_bytes(d, heap + e.val, e.aux); }
        else dg_cell(d, e.tag, e.val, e.aux, heap, 1);
      }
      break;
    }
    default: dg_u64(d, val); dg_u64(d, aux); break;
  }
}

/* digest of records [0, n_valid) of `p` (n_valid = n_records, or first_error.record_index) */
void orc_planes_digest(const etl_dec_planes* p, uint64_t n_valid, uint64_t out[4]) {
  dg d; for (int i = 0; i < 4; i++) d.h[i] = K[i] ^ (uint64_t)(i + 1);
  dg_u64(&d, n_valid);
  for (uint64_t r = 0; r < n_valid; r++) {
    dg_u64(&d, p->rec_off[r]);
    dg_u64(&d, p->rec_kind[r] | ((uint64_t)p->rec_flags[r] << 8) | ((uint64_t)p->rec_rel[r] << 16));
    dg_u64(&d, (uint64_t)(uint32_t)p->rec_schema[r] | ((uint64_t)p->rec_tuple_bytes[r] << 32));
    dg_u64(&d, p->rec_start_lsn[r]); dg_u64(&d, p->rec_commit_lsn[r]); dg_u64(&d, p->rec_tx_ordinal[r]);
    dg_u64(&d, p->rec_cell_base[r]); dg_u64(&d, p->rec_heap_hint[r]);
  }
  dg_u64(&d, p->rec_cell_base[n_valid]);
  const uint64_t m = p->rec_cell_base[n_valid];
  for (uint64_t c = 0; c < m; c++) dg_cell(&d, p->cell_tag[c], p->cell_val[c], p->cell_aux[c], p->heap, 0);
  for (int i = 0; i < 4; i++) out[i] = d.h[i];
}
/* the same over an oracle batch */
void orc_batch_digest(const orc_batch* b, uint64_t out[4]) {
  etl_dec_planes p; memset(&p, 0, sizeof p);
  p.n_records = b->n_records; p.n_cells = b->n_cells; p.heap_bytes = b->heap_bytes;
  p.rec_off = b->rec_off; p.rec_kind = b->rec_kind; p.rec_flags = b->rec_flags; p.rec_rel = b->rec_rel; p.rec_schema = b->rec_schema;
  p.rec_start_lsn = b->rec_start_lsn; p.rec_commit_lsn = b->rec_commit_lsn; p.rec_tx_ordinal = b->rec_tx_ordinal;
  p.rec_cell_base = b->rec_cell_base; p.rec_tuple_bytes = b->rec_tuple_bytes; p.rec_heap_hint = b->rec_heap_hint;
  p.cell_tag = b->cell_tag; p.cell_val = b->cell_val; p.cell_aux = b->cell_aux; p.heap = b->heap;
  const uint64_t nv = b->first_error.record_index == UINT64_MAX ? b->n_records : b->first_error.record_index;
  orc_planes_digest(&p, nv, out);
}

/* ---- COPY rows: canonical digest of decoded rows, strings by CONTENT (the device keeps a field
 * that needs no unescaping as a span of the staged stream and copies the others into its heap; the oracle keeps
 * every unescaped field in its own text buffer). */
static void dg_copy_cell(dg* d, uint32_t tag, uint64_t val, uint32_t aux, const uint8_t* text_src, const uint8_t* heap) {
  if (tag == ETL_CELL_STRING || tag == ETL_CELL_JSON) { dg_u64(d, tag); dg_bytes(d, text_src + val, aux); }
  else dg_cell(d, tag, val, aux, heap, 1);   /* in_array = 1: pushed_groups is not part of a COPY row's contract */
}
/* device planes: n_valid_rows x n_cols cells; string / json val = offset into `stream`, or bit 63 set = offset into `heap` */
void orc_copy_planes_digest(const uint8_t* tags, const uint64_t* vals, const uint32_t* auxs, uint64_t n_valid_rows, uint32_t n_cols,
                            const uint8_t* stream, const uint8_t* heap, uint64_t out[4]) {
  dg d; for (int i = 0; i < 4; i++) d.h[i] = K[i] ^ (uint64_t)(i + 11);
  dg_u64(&d, n_valid_rows); dg_u64(&d, n_cols);
  for (uint64_t c = 0; c < n_valid_rows * n_cols; c++) {
    const uint64_t v = vals[c];
    const int spanned = tags[c] == ETL_CELL_STRING || tags[c] == ETL_CELL_JSON;   /* only these carry the in-heap flag: bit 63 of an int cell is its sign */
    const int in_heap = spanned && (v >> 63);
    dg_copy_cell(&d, tags[c], spanned ? (v & ~(1ull << 63)) : v, auxs[c], in_heap ? heap : stream, heap);
  }
  for (int i = 0; i < 4; i++) out[i] = d.h[i];
}
/* oracle side: parse rows [0, n_rows) given by row_off (n_rows + 1 offsets into buf) one by one with orc_parse_copy_row
 * (table_row.rs:25-165), stop at the first failing row; digest of the rows before it. */
void orc_copy_rows_digest(const uint32_t* type_oids, uint32_t n_cols, const uint8_t* buf, const uint64_t* row_off, uint64_t n_rows,
                          uint64_t out[4], uint64_t* err_row, uint32_t* err_col, uint32_t* err_code) {
  dg d; for (int i = 0; i < 4; i++) d.h[i] = K[i] ^ (uint64_t)(i + 11);
  *err_row = UINT64_MAX; *err_col = 0xFFFFFFFFu; *err_code = 0;
  uint64_t cap = 1 << 16;
  uint8_t* tags = 0; uint64_t* vals = 0; uint32_t* auxs = 0; uint8_t* text = 0; uint8_t* heap = 0;
  tags = (uint8_t*)malloc(n_cols + 1); vals = (uint64_t*)malloc((n_cols + 1) * 8); auxs = (uint32_t*)malloc((n_cols + 1) * 4);
  text = (uint8_t*)malloc(cap); heap = (uint8_t*)malloc(cap * 24);
  /* first pass: find the first failing row so that the row count is hashed first (as the device side does) */
  uint64_t n_ok = n_rows;
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) { dg_u64(&d, n_ok); dg_u64(&d, n_cols); }
    for (uint64_t r = 0; r < (pass ? n_ok : n_rows); r++) {
      const uint64_t len = row_off[r + 1] - row_off[r];
      if (len + 64 > cap) { cap = (len + 64) * 2; text = (uint8_t*)realloc(text, cap); heap = (uint8_t*)realloc(heap, cap * 24); }
      uint32_t nv = 0, ec = 0; uint64_t tl = 0, hl = 0;
      const uint32_t e = orc_parse_copy_row(type_oids, n_cols, buf + row_off[r], len, tags, vals, auxs, &nv, text, cap, &tl, heap, cap * 24, &hl, &ec);
      if (e) { if (!pass) { n_ok = r; *err_row = r; *err_col = ec; *err_code = e; } break; }
      if (pass) for (uint32_t c = 0; c < n_cols; c++) dg_copy_cell(&d, tags[c], vals[c], auxs[c], text, heap);
    }
  }
  free(tags); free(vals); free(auxs); free(text); free(heap);
  for (int i = 0; i < 4; i++) out[i] = d.h[i];
}
