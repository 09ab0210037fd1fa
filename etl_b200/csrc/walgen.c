/*
 * walgen.c — deterministic synthetic pgoutput stream generator for the BASELINE.json workloads
 * (SURVEY.md §8d).  Host-side tooling for tests and bench.py; not part of the decode path.
 *
 * A stream is produced as independent SEGMENTS (own PCG32 stream, own LSN range, begins with the
 * Relation messages of its connection epoch, ends on a Commit) so segments can be generated in
 * parallel threads and each GPU rank can generate only the byte range it owns.  Text spellings
 * are exactly what the reference's pinned session produces (datestyle=ISO, timezone=UTC,
 * extra_float_digits=3 — crates/etl-config/src/shared/connection.rs:22-26).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

enum {
  WG_SEQ_INT8 = 1, WG_INT4_FULL, WG_INT4_RANGE, WG_INT8_FULL, WG_BOOL, WG_TEXT_LOGNORMAL, WG_TEXT_UNIFORM,
  WG_TIMESTAMPTZ, WG_NUMERIC, WG_JSONB, WG_TOAST_TEXT, WG_UUID, WG_DATE, WG_FLOAT8, WG_BYTEA, WG_TIMESTAMP,
  WG_TIME, WG_INT2, WG_FLOAT4, WG_OID,
};

typedef struct wg_col {
  uint32_t type_oid;
  uint8_t nullable, is_pk, gen, _pad;
  uint32_t p0, p1;
  char name[32];
} wg_col;

typedef struct wg_table {
  uint32_t rel_id;
  uint8_t replident; /* 'd' | 'f' */
  uint8_t _pad;
  uint16_t n_cols;
  const wg_col* cols;
  char name[32];
} wg_table;

typedef struct wg_cfg {
  uint64_t seed;
  uint32_t n_tables, _pad0;
  const wg_table* tables;
  uint64_t n_msgs;        /* DML messages in this segment (0 = unbounded) */
  uint64_t target_bytes;  /* stop at the first Commit at or past this many bytes (0 = unbounded) */
  uint32_t pct_insert, pct_update, pct_delete; /* per 10000, must sum to 10000 */
  uint32_t pct_key_change;     /* of updates on 'd' tables, per 10000 */
  uint32_t tx_mean;            /* mean DML per transaction (geometric) ; */
  uint32_t tx_fixed;           /* if non-zero: exactly this many DML per transaction */
  uint32_t null_pct;           /* per 10000 on nullable columns */
  uint32_t nonascii_pct;       /* per 10000 of text cells */
  uint32_t toast_row_pct;      /* per 10000 rows carry a TOAST-sized value in WG_TOAST_TEXT columns */
  uint32_t toast_unchanged_pct;/* per 10000 of updates leave WG_TOAST_TEXT columns 'u' */
  uint32_t keepalive_every;    /* a keepalive frame every N DML messages (0 = none) */
  uint32_t schema_bump_ppm;    /* per million DML: re-send the Relation with a flipped replica identity */
  uint32_t relations_once;     /* 1: ONE stream — only segment 0 carries the Relation messages (segments > 0 continue the connection) */
  uint32_t _pad1;
} wg_cfg;

typedef struct wg_stats {
  uint64_t bytes, frames, dml, inserts, updates, deletes, txs, relations, cells;
} wg_stats;

/* ---------------------------------------------------------------- PCG32 */
typedef struct { uint64_t state, inc; } rng_t;
static uint32_t rnd(rng_t* r) {
  uint64_t old = r->state;
  r->state = old * 6364136223846793005ULL + r->inc;
  uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
  return (xs >> rot) | (xs << ((-rot) & 31));
}
static void rng_seed(rng_t* r, uint64_t seed, uint64_t seq) {
  r->state = 0; r->inc = (seq << 1u) | 1u; rnd(r); r->state += seed; rnd(r);
}
static uint32_t rnd_below(rng_t* r, uint32_t n) { return n ? (uint32_t)(((uint64_t)rnd(r) * n) >> 32) : 0; }
static uint64_t rnd64(rng_t* r) { return ((uint64_t)rnd(r) << 32) | rnd(r); }
static double rnd_unit(rng_t* r) { return (rnd(r) + 0.5) / 4294967296.0; }
static double rnd_normal(rng_t* r) { return sqrt(-2.0 * log(rnd_unit(r))) * cos(6.283185307179586 * rnd_unit(r)); }

/* ---------------------------------------------------------------- writer */
typedef struct { uint8_t* p; uint64_t len, cap; int overflow; } out_t;
static void put(out_t* o, const void* s, uint64_t n) {
  if (o->len + n > o->cap) { o->overflow = 1; return; }
  memcpy(o->p + o->len, s, n); o->len += n;
}
static void put8(out_t* o, uint8_t v) { put(o, &v, 1); }
static void put16(out_t* o, uint16_t v) { uint8_t b[2] = {(uint8_t)(v >> 8), (uint8_t)v}; put(o, b, 2); }
static void put32(out_t* o, uint32_t v) { uint8_t b[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v}; put(o, b, 4); }
static void put64(out_t* o, uint64_t v) { put32(o, (uint32_t)(v >> 32)); put32(o, (uint32_t)v); }
static void patch32(out_t* o, uint64_t at, uint32_t v) { if (at + 4 <= o->cap) { o->p[at] = (uint8_t)(v >> 24); o->p[at + 1] = (uint8_t)(v >> 16); o->p[at + 2] = (uint8_t)(v >> 8); o->p[at + 3] = (uint8_t)v; } }
static void patch64(out_t* o, uint64_t at, uint64_t v) { patch32(o, at, (uint32_t)(v >> 32)); patch32(o, at + 4, (uint32_t)v); }

typedef struct { out_t* o; uint64_t lsn; int64_t clock; uint64_t frame_at; wg_stats* st; } wr_t;
/* open a CopyData frame carrying an XLogData message; returns so the caller appends the pgoutput body */
static void frame_open(wr_t* w) {
  w->frame_at = w->o->len;
  put8(w->o, 'd'); put32(w->o, 0);
  put8(w->o, 'w'); put64(w->o, w->lsn); put64(w->o, w->lsn + 0x1000); put64(w->o, (uint64_t)w->clock);
}
static void frame_close(wr_t* w) {
  uint64_t flen = w->o->len - w->frame_at - 1;
  patch32(w->o, w->frame_at + 1, (uint32_t)flen);
  w->lsn += (w->o->len - w->frame_at - 30) + 8;
  w->clock += 3;
  w->st->frames++;
}

/* ---------------------------------------------------------------- value text generators */
static const char ALNUM[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789      ";
static void gen_text(out_t* o, rng_t* r, uint32_t len, int nonascii) {
  /* length-prefixed 't' cell body is written by the caller; this emits exactly len bytes */
  uint64_t start = o->len;
  if (o->len + len > o->cap) { o->overflow = 1; return; }
  uint8_t* d = o->p + o->len;
  uint32_t i = 0;
  while (i + 4 <= len) { uint32_t x = rnd(r); d[i] = ALNUM[x & 63]; d[i + 1] = ALNUM[(x >> 6) & 63]; d[i + 2] = ALNUM[(x >> 12) & 63]; d[i + 3] = ALNUM[(x >> 18) & 63]; i += 4; }
  while (i < len) d[i++] = ALNUM[rnd(r) & 63];
  o->len += len;
  if (nonascii && len >= 4) {
    static const char* mb[] = {"\xc3\xa9", "\xe2\x9c\x93", "\xf0\x9f\xa4\x94", "\xc3\xbc"};
    const char* s = mb[rnd_below(r, 4)];
    uint32_t l = (uint32_t)strlen(s);
    uint32_t at = rnd_below(r, len - l + 1);
    memcpy(o->p + start + at, s, l);
  }
}
static int fmt_i64(char* b, int64_t v) { return sprintf(b, "%lld", (long long)v); }
static void civil_from_days(int64_t z, int* y, int* m, int* d) {
  z += 719468;
  int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  int64_t doe = z - era * 146097;
  int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  int64_t yy = yoe + era * 400;
  int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  int64_t mp = (5 * doy + 2) / 153;
  *d = (int)(doy - (153 * mp + 2) / 5 + 1);
  *m = (int)(mp < 10 ? mp + 3 : mp - 9);
  *y = (int)(yy + (*m <= 2));
}
static int fmt_ts(char* b, int64_t us_since_1970, int with_tz, int with_date, int with_time) {
  int64_t secs = us_since_1970 / 1000000, us = us_since_1970 % 1000000;
  int64_t days = secs / 86400, sod = secs % 86400;
  int y, m, d, n = 0;
  civil_from_days(days, &y, &m, &d);
  if (with_date) n += sprintf(b + n, "%04d-%02d-%02d", y, m, d);
  if (with_date && with_time) b[n++] = ' ';
  if (with_time) {
    n += sprintf(b + n, "%02d:%02d:%02d", (int)(sod / 3600), (int)(sod / 60 % 60), (int)(sod % 60));
    if (us) { n += sprintf(b + n, ".%06d", (int)us); while (b[n - 1] == '0') n--; }
  }
  if (with_tz) n += sprintf(b + n, "+00");
  return n;
}
static int gen_numeric(char* b, rng_t* r) {
  if (rnd_below(r, 100) == 0) return sprintf(b, "NaN");
  int n = 0;
  if (rnd_below(r, 10) < 3) b[n++] = '-';
  int ip = 1 + (int)rnd_below(r, 28), sc = (int)rnd_below(r, 11);
  b[n++] = (char)('1' + rnd_below(r, 9));
  for (int i = 1; i < ip; i++) b[n++] = (char)('0' + rnd_below(r, 10));
  if (sc) { b[n++] = '.'; for (int i = 0; i < sc; i++) b[n++] = (char)('0' + rnd_below(r, 10)); }
  return n;
}
static int gen_jsonb(char* b, rng_t* r) {
  int nk = 3 + (int)rnd_below(r, 4), n = 0;
  b[n++] = '{';
  for (int k = 0; k < nk; k++) {
    if (k) { b[n++] = ','; b[n++] = ' '; }
    n += sprintf(b + n, "\"k%d\": ", k);
    switch (rnd_below(r, 4)) {
      case 0: n += sprintf(b + n, "%d", (int)rnd_below(r, 1000000)); break;
      case 1: { b[n++] = '"'; int l = 1 + (int)rnd_below(r, 8); for (int i = 0; i < l; i++) b[n++] = ALNUM[rnd(r) % 52]; b[n++] = '"'; break; }
      case 2: n += sprintf(b + n, rnd_below(r, 2) ? "true" : "false"); break;
      default: n += sprintf(b + n, "%d.%02d", (int)rnd_below(r, 10000), (int)rnd_below(r, 100)); break;
    }
  }
  b[n++] = '}';
  return n;
}
static uint32_t lognormal_len(rng_t* r, uint32_t median, uint32_t maxlen) {
  double v = exp(log((double)(median ? median : 1)) + 0.8 * rnd_normal(r));
  if (v > maxlen) v = maxlen;
  return (uint32_t)v;
}

/* writes one tuple cell ('n' or 't'+len+bytes) for column c */
static void gen_cell(wr_t* w, rng_t* r, const wg_cfg* cfg, const wg_col* c, uint64_t row_id, int toast_row) {
  out_t* o = w->o;
  w->st->cells++;
  if (c->nullable && !c->is_pk && rnd_below(r, 10000) < cfg->null_pct) { put8(o, 'n'); return; }
  char b[192];
  int n = -1;
  uint32_t tl;
  switch (c->gen) {
    case WG_SEQ_INT8: n = fmt_i64(b, (int64_t)row_id); break;
    case WG_INT4_FULL: n = fmt_i64(b, (int32_t)rnd(r)); break;
    case WG_INT4_RANGE: n = fmt_i64(b, (int64_t)c->p0 + rnd_below(r, c->p1 - c->p0 + 1)); break;
    case WG_INT2: n = fmt_i64(b, (int16_t)rnd(r)); break;
    case WG_OID: n = sprintf(b, "%u", rnd(r)); break;
    case WG_INT8_FULL: n = fmt_i64(b, (int64_t)rnd64(r)); break;
    case WG_BOOL: b[0] = rnd_below(r, 2) ? 't' : 'f'; n = 1; break;
    case WG_TIMESTAMPTZ: n = fmt_ts(b, 1577836800000000LL + (int64_t)(rnd64(r) % 315532800000000ULL), 1, 1, 1); break;
    case WG_TIMESTAMP: n = fmt_ts(b, 1577836800000000LL + (int64_t)(rnd64(r) % 315532800000000ULL), 0, 1, 1); break;
    case WG_DATE: n = fmt_ts(b, 1577836800000000LL + (int64_t)(rnd64(r) % 315532800000000ULL), 0, 1, 0); break;
    case WG_TIME: n = fmt_ts(b, (int64_t)(rnd64(r) % 86400000000ULL), 0, 0, 1); break;
    case WG_NUMERIC: n = gen_numeric(b, r); break;
    case WG_JSONB: n = gen_jsonb(b, r); break;
    case WG_FLOAT8: { double v = (rnd_unit(r) - 0.5) * pow(10.0, (double)rnd_below(r, 20) - 5.0); n = sprintf(b, "%.17g", v); break; }
    case WG_FLOAT4: { float v = (float)((rnd_unit(r) - 0.5) * pow(10.0, (double)rnd_below(r, 12) - 3.0)); n = sprintf(b, "%.9g", (double)v); break; }
    case WG_UUID: { uint64_t a = rnd64(r), c2 = rnd64(r); n = sprintf(b, "%08x-%04x-%04x-%04x-%012llx", (uint32_t)(a >> 32), (uint32_t)(a >> 16) & 0xffff, (uint32_t)a & 0xffff, (uint32_t)(c2 >> 48), (unsigned long long)(c2 & 0xffffffffffffULL)); break; }
    case WG_BYTEA: { int l = (int)rnd_below(r, 24); b[0] = '\\'; b[1] = 'x'; n = 2; for (int i = 0; i < l; i++) n += sprintf(b + n, "%02x", rnd(r) & 0xff); break; }
    case WG_TEXT_UNIFORM: tl = c->p0 + rnd_below(r, c->p1 - c->p0 + 1); goto text;
    case WG_TEXT_LOGNORMAL: tl = lognormal_len(r, c->p0, c->p1); goto text;
    case WG_TOAST_TEXT: tl = toast_row ? c->p0 + rnd_below(r, c->p1 - c->p0 + 1) : lognormal_len(r, 24, 256); goto text;
    default: n = 0; break;
  }
  put8(o, 't'); put32(o, (uint32_t)n); put(o, b, (uint64_t)n);
  return;
text:
  put8(o, 't'); put32(o, tl);
  gen_text(o, r, tl, rnd_below(r, 10000) < cfg->nonascii_pct);
}

static void emit_relation(wr_t* w, const wg_table* t, uint8_t replident) {
  out_t* o = w->o;
  frame_open(w);
  put8(o, 'R'); put32(o, t->rel_id);
  put(o, "public", 7); put(o, t->name, strlen(t->name) + 1);
  put8(o, replident); put16(o, t->n_cols);
  for (int i = 0; i < t->n_cols; i++) {
    const wg_col* c = &t->cols[i];
    put8(o, (uint8_t)(c->is_pk ? 1 : 0));
    put(o, c->name, strlen(c->name) + 1);
    put32(o, c->type_oid); put32(o, 0xffffffffu);
  }
  frame_close(w);
  w->st->relations++;
}

static void emit_tuple(wr_t* w, rng_t* r, const wg_cfg* cfg, const wg_table* t, uint64_t row_id, int toast_row,
                       int key_only, int unchanged_toast) {
  out_t* o = w->o;
  put16(o, t->n_cols);
  for (int i = 0; i < t->n_cols; i++) {
    const wg_col* c = &t->cols[i];
    if (key_only && !c->is_pk) { put8(o, 'n'); w->st->cells++; continue; } /* full-width key tuple */
    if (unchanged_toast && c->gen == WG_TOAST_TEXT) { put8(o, 'u'); w->st->cells++; continue; }
    gen_cell(w, r, cfg, c, row_id, toast_row);
  }
}

/* Generates segment `seg`. Returns bytes written, or (uint64_t)-1 if `cap` was too small. */
uint64_t wg_generate_segment(const wg_cfg* cfg, uint64_t seg, uint8_t* out, uint64_t cap, wg_stats* stats) {
  wg_stats st; memset(&st, 0, sizeof st);
  out_t o = {out, 0, cap, 0};
  rng_t r; rng_seed(&r, cfg->seed, seg + 1);
  wr_t w = {&o, 0x0100000000ULL + (seg << 40), 631152000000000LL + (int64_t)seg * 1000000, 0, &st};
  uint8_t sent[4096]; uint8_t ident[4096];
  memset(sent, 0, sizeof sent);
  for (uint32_t i = 0; i < cfg->n_tables && i < 4096; i++) ident[i] = cfg->tables[i].replident;
  if (cfg->relations_once && seg > 0) memset(sent, 1, sizeof sent);
  uint64_t row_id = seg * 1000000007ULL + 1;
  uint64_t dml = 0;
  int done = 0;
  while (!done) {
    /* ---- Begin */
    uint32_t tx_n = cfg->tx_fixed;
    if (!tx_n) { double u = rnd_unit(&r); double m = cfg->tx_mean ? cfg->tx_mean : 1; tx_n = 1 + (uint32_t)(log(u) / log(1.0 - 1.0 / (m + 1e-9) + 1e-12)); if (tx_n > 100000) tx_n = 100000; }
    frame_open(&w);
    uint64_t begin_lsn_at = o.len + 1;
    put8(&o, 'B'); put64(&o, 0); put64(&o, (uint64_t)w.clock); put32(&o, (uint32_t)(1000 + st.txs + seg * 100000));
    frame_close(&w);
    for (uint32_t k = 0; k < tx_n; k++) {
      if (cfg->n_msgs && dml >= cfg->n_msgs) break;
      uint32_t ti = cfg->n_tables > 1 ? rnd_below(&r, cfg->n_tables) : 0;
      const wg_table* t = &cfg->tables[ti];
      if (cfg->schema_bump_ppm && ti < 4096 && sent[ti] && rnd_below(&r, 1000000) < cfg->schema_bump_ppm) {
        ident[ti] = (ident[ti] == 'f') ? 'd' : 'f';
        sent[ti] = 0;
      }
      if (ti < 4096 && !sent[ti]) { emit_relation(&w, t, ident[ti]); sent[ti] = 1; }
      uint8_t ri = ti < 4096 ? ident[ti] : t->replident;
      uint32_t op = rnd_below(&r, 10000);
      int toast_row = rnd_below(&r, 10000) < cfg->toast_row_pct;
      frame_open(&w);
      if (op < cfg->pct_insert) {
        put8(&o, 'I'); put32(&o, t->rel_id); put8(&o, 'N');
        emit_tuple(&w, &r, cfg, t, row_id++, toast_row, 0, 0);
        st.inserts++;
      } else if (op < cfg->pct_insert + cfg->pct_update) {
        put8(&o, 'U'); put32(&o, t->rel_id);
        int unchanged = rnd_below(&r, 10000) < cfg->toast_unchanged_pct;
        uint64_t id = row_id - 1 - rnd_below(&r, 1000);
        if (ri == 'f') { put8(&o, 'O'); emit_tuple(&w, &r, cfg, t, id, unchanged ? 1 : toast_row, 0, 0); }
        else if (rnd_below(&r, 10000) < cfg->pct_key_change) { put8(&o, 'K'); emit_tuple(&w, &r, cfg, t, id, 0, 1, 0); id = row_id++; }
        put8(&o, 'N');
        emit_tuple(&w, &r, cfg, t, id, toast_row, 0, unchanged);
        st.updates++;
      } else {
        put8(&o, 'D'); put32(&o, t->rel_id);
        uint64_t id = row_id - 1 - rnd_below(&r, 1000);
        if (ri == 'f') { put8(&o, 'O'); emit_tuple(&w, &r, cfg, t, id, toast_row, 0, 0); }
        else { put8(&o, 'K'); emit_tuple(&w, &r, cfg, t, id, 0, 1, 0); }
        st.deletes++;
      }
      frame_close(&w);
      dml++;
      if (cfg->keepalive_every && dml % cfg->keepalive_every == 0) {
        put8(&o, 'd'); put32(&o, 22); put8(&o, 'k'); put64(&o, w.lsn); put64(&o, (uint64_t)w.clock); put8(&o, 0);
        st.frames++;
      }
    }
    /* ---- Commit: Begin.final_lsn = Commit.commit_lsn = wal_start of the commit frame */
    uint64_t commit_lsn = w.lsn;
    patch64(&o, begin_lsn_at, commit_lsn);
    frame_open(&w);
    put8(&o, 'C'); put8(&o, 0); put64(&o, commit_lsn); put64(&o, commit_lsn + 0x38); put64(&o, (uint64_t)w.clock);
    frame_close(&w);
    st.txs++;
    if (cfg->n_msgs && dml >= cfg->n_msgs) done = 1;
    if (cfg->target_bytes && o.len >= cfg->target_bytes) done = 1;
    if (!cfg->n_msgs && !cfg->target_bytes) done = 1;
    if (o.overflow) break;
  }
  st.bytes = o.len; st.dml = dml;
  if (stats) *stats = st;
  return o.overflow ? (uint64_t)-1 : o.len;
}
