/* oracle_copy.c — CPU restatement of the COPY-text initial-sync row parser.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Groundwork for the next hot-path row: no device path binds
 * to this yet; the restatement is pinned to the reference's own tests (tests/test_oracle_copy_rows.py).
 *
 * Follows crates/etl/src/conversions/table_row.rs:25-165 step by step:
 *   - the whole row must be UTF-8 (:33);
 *   - fields end at TAB, the row at LF; a backslash escapes the next char: b f n r t v map to the control
 *     characters, `\N` stays the two characters `\N`, anything else is the character itself (:46-71);
 *   - after the LF the scan continues to the end of the input; whatever follows the last LF without a
 *     terminator of its own is dropped when the input ends (:88-96) — but only if an LF was seen at all,
 *     otherwise "Row data not properly terminated";
 *   - a field whose unescaped text is exactly `\N` is Cell::Null (:116-121), everything else goes through
 *     parse_cell_from_postgres_text (:126);
 *   - one field too many → column-count error as soon as it is met (:103-113); too few → at the end (:150-160).
 */
#include <stdlib.h>
#include <string.h>

#include "oracle_internal.h"

uint32_t orc_parse_copy_row(const uint32_t* type_oids, uint32_t n_cols, const uint8_t* row, uint64_t len,
                            uint8_t* tags, uint64_t* vals, uint32_t* auxs, uint32_t* n_values,
                            uint8_t* text_out, uint64_t text_cap, uint64_t* text_len,
                            uint8_t* heap_out, uint64_t heap_cap, uint64_t* heap_len, uint32_t* err_col) {
  *n_values = 0; *text_len = 0; *heap_len = 0; *err_col = 0xFFFFFFFFu;
  if (!orc_utf8_valid(row, len)) return ETL_E_UTF8;
  orc_heap heap = {0, 0, 0};
  uint8_t* text = (uint8_t*)malloc(len + 16);   /* unescaped field values, back to back */
  uint64_t tl = 0, field0 = 0;
  uint32_t col = 0, err = 0;
  int in_escape = 0, terminated = 0, done = 0;
  uint64_t i = 0;
  while (!done && !err) {
    for (;;) {
      if (i >= len) {
        if (!terminated) err = ORC_E_COPY_NOT_TERMINATED;
        done = 1;
        break;
      }
      const uint8_t c = row[i++];                /* every structural character is ASCII; other bytes copy through */
      if (in_escape) {
        if (c == 'N') { text[tl++] = '\\'; text[tl++] = 'N'; }
        else if (c == 'b') text[tl++] = 8;
        else if (c == 'f') text[tl++] = 12;
        else if (c == 'n') text[tl++] = '\n';
        else if (c == 'r') text[tl++] = '\r';
        else if (c == 't') text[tl++] = '\t';
        else if (c == 'v') text[tl++] = 11;
        else text[tl++] = c;
        in_escape = 0;
      } else if (c == '\t') break;
      else if (c == '\n') { terminated = 1; break; }
      else if (c == '\\') in_escape = 1;
      else text[tl++] = c;
    }
    if (done || err) break;
    if (col >= n_cols) { err = ORC_E_COPY_COLUMN_COUNT; *err_col = col; break; }
    const uint64_t flen = tl - field0;
    orc_cell cell; memset(&cell, 0, sizeof cell);
    if (flen == 2 && text[field0] == '\\' && text[field0 + 1] == 'N') cell.tag = ETL_CELL_NULL;
    else {
      err = orc_parse_text(orc_kind_for_oid(type_oids[col]), text + field0, flen, field0, &heap, &cell);
      if (err) { *err_col = col; break; }
    }
    tags[col] = cell.tag; vals[col] = cell.val; auxs[col] = cell.aux;
    col++;
    field0 = tl;
  }
  if (!err && col < n_cols) { err = ORC_E_COPY_COLUMN_COUNT; *err_col = col; }
  *n_values = col;
  *text_len = tl; *heap_len = heap.len;
  if (text_out) memcpy(text_out, text, tl < text_cap ? tl : text_cap);
  if (heap_out && heap.len) memcpy(heap_out, heap.data, heap.len < heap_cap ? heap.len : heap_cap);
  free(text); free(heap.data);
  return err;
}
