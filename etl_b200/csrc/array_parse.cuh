// array_parse.cuh — the exact per-kind dispatch (text.rs:28-173) and the array splitter (text.rs:184-249).
// Kept apart from the kernels so that tests/emul/host_parsers.cpp can compile exactly these sources for the
// host and fuzz them against the oracle; wal_kernels.cuh includes this file where the code used to live.
#pragma once
#include <stdint.h>

#include "cell_parsers.cuh"
#include "float_parse.cuh"

namespace etl {

// text.rs:28-173 dispatch for one text cell (bytes already UTF-8 validated). `soff` = absolute
// stream offset of the value bytes.
__device__ __forceinline__ uint32_t parse_text_cell_impl(uint32_t kind, const uint8_t* s, uint32_t n, uint64_t soff,
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
    case ETL_K_F32: return parse_float(s, n, true, o);
    case ETL_K_F64: return parse_float(s, n, false, o);
    default: return ETL_E_MALFORMED_FRAME;  // unsupported decode class: rejected on the host before launch
  }
}

// out-of-line entry for k_heavy's cold kinds: heap position by value, so the caller's state stays in registers
__device__ __noinline__ uint32_t parse_text_cell(uint32_t kind, const uint8_t* s, uint32_t n, uint64_t soff, uint8_t* heap,
                                                 uint64_t hpos, CellOut& o) {
  HeapCursor hc{heap, hpos};
  return parse_text_cell_impl(kind, s, n, soff, hc, o);
}

// ---- arrays (text.rs:184-249): one-dimensional split — `"` toggles quoting, `\` escapes the next
// char, `,` splits outside quotes, an unquoted case-insensitive NULL is a null element.  Elements are
// unescaped into the heap and parsed there with the exact scalar parsers.  Thread-serial: arrays are
// off the named hot configurations; what matters is that accept/reject and every value match.
// Returns 0xFFFFFFFE when the heap reservation does not fit (host retries with a larger heap).
struct ArrHeap { uint8_t* heap; unsigned long long* arr_top; uint64_t arr_base, heap_cap; };
__device__ __noinline__ uint32_t parse_array_cell(const ArrHeap P, uint32_t ekind, const uint8_t* s, uint32_t n,
                                                  int tz_fmt /*0 n/a, 1 %#z, 2 %:z*/, CellOut& o) {
  if (n < 2) return ETL_E_ARRAY_SHORT;
  if (s[0] != '{' || s[n - 1] != '}') return ETL_E_ARRAY_BRACES;
  const uint8_t* p = s + 1;
  const uint32_t m = n - 2;
  // pass 1: element count
  uint32_t ne = 0;
  {
    bool in_q = false, in_e = false;
    uint32_t commas = 0;
    for (uint32_t i = 0; i < m; i++) {
      const uint32_t ch = p[i];
      if (in_e) { in_e = false; continue; }
      if (ch == '"') in_q = !in_q;
      else if (ch == '\\') in_e = true;
      else if (ch == ',' && !in_q) commas++;
    }
    ne = m ? commas + 1 : 0;
  }
  const uint64_t need = 16ull + 44ull * ne + (uint64_t)m + m / 2u;   // hdr + elems + unescaped text + numeric/bytes payloads
  // arrays bump-allocate in their own region [arr_base, heap_cap) so the scalar region's bound stays exact
  const uint64_t base = P.arr_base + atomicAdd(P.arr_top, (unsigned long long)((need + 7ull) & ~7ull));
  if (base + need > P.heap_cap) return 0xFFFFFFFEu;
  etl_array_hdr hdr;
  hdr.elem_kind = (uint8_t)ekind; hdr._pad[0] = hdr._pad[1] = hdr._pad[2] = 0; hdr.n_elems = ne;
  *reinterpret_cast<etl_array_hdr*>(P.heap + base) = hdr;
  etl_array_elem* elems = reinterpret_cast<etl_array_elem*>(P.heap + base + 8);
  HeapCursor hc{P.heap, base + 8 + 16ull * ne};
  // pass 2
  bool in_q = false, in_e = false, quoted = false;
  uint32_t i = 0, k = 0;
  bool done = (m == 0);
  uint32_t err = 0;
  while (!done && !err) {
    const uint64_t voff = hc.pos;
    uint8_t* val = P.heap + voff;
    uint32_t vl = 0;
    for (;;) {
      if (i >= m) { done = true; break; }
      const uint32_t ch = p[i];
      if (in_e) { val[vl++] = (uint8_t)ch; in_e = false; i++; continue; }
      if (ch == '"') { if (!in_q) quoted = true; in_q = !in_q; i++; continue; }
      if (ch == '\\') { in_e = true; i++; continue; }
      if (ch == ',' && !in_q) { i++; break; }
      val[vl++] = (uint8_t)ch; i++;
    }
    hc.pos += (vl + 7u) & ~7u;
    etl_array_elem e;
    e.val = 0; e.aux = 0; e.tag = ETL_CELL_NULL; e._pad[0] = e._pad[1] = e._pad[2] = 0;
    if (!(!quoted && ieq(val, vl, "null", 4))) {
      CellOut eo;
      eo.val = 0; eo.aux = 0; eo.tag = 0;
      if (ekind == ETL_K_STRING) { eo.tag = ETL_CELL_STRING; eo.val = voff; eo.aux = vl; }
      else if (ekind == ETL_K_JSON) { if (json_valid(val, vl)) { eo.tag = ETL_CELL_JSON; eo.val = voff; eo.aux = vl; } else err = ETL_E_JSON; }
      else if (ekind == ETL_K_TIMESTAMPTZ) { if (!parse_timestamptz_fmt(val, vl, tz_fmt == 1, eo)) err = ETL_E_DATETIME; }
      else err = parse_text_cell_impl(ekind, val, vl, voff, hc, eo);
      e.val = eo.val; e.aux = eo.aux; e.tag = (uint8_t)eo.tag;
    }
    if (!err && k < ne) elems[k++] = e;
    quoted = false;
  }
  if (err) return err;
  o.tag = ETL_CELL_ARRAY; o.val = base; o.aux = ne;
  return 0;
}
__device__ __noinline__ uint32_t parse_array_any(const ArrHeap P, uint32_t kind, const uint8_t* s, uint32_t n, CellOut& o) {
  const uint32_t ek = kind & ~(uint32_t)ETL_K_ARRAY;
  if (ek == ETL_K_TIMESTAMPTZ) {                       // text.rs:117-140: whole-array retry with the second format
    const uint32_t e = parse_array_cell(P, ek, s, n, 1, o);
    if (e == 0 || e == 0xFFFFFFFEu) return e;
    return parse_array_cell(P, ek, s, n, 2, o);
  }
  return parse_array_cell(P, ek, s, n, 0, o);
}

}  // namespace etl
