/* oracle_internal.h — shared between oracle_cells.c and oracle_stream.c. TEST INFRASTRUCTURE ONLY. */
#ifndef ETL_ORACLE_INTERNAL_H
#define ETL_ORACLE_INTERNAL_H
#include "oracle.h"

typedef struct orc_heap { uint8_t* data; uint64_t len, cap; } orc_heap;
typedef struct orc_cell { uint64_t val; uint32_t aux; uint8_t tag; } orc_cell;

uint64_t orc_heap_alloc(orc_heap* h, uint64_t n, uint64_t align);
int orc_utf8_valid(const uint8_t* s, uint64_t n);
int orc_json_valid(const uint8_t* s, uint64_t n);
/* text.rs:28 on already-UTF-8-validated text; kind = ETL_K_* */
uint32_t orc_parse_text(uint32_t kind, const uint8_t* s, uint64_t n, uint64_t stream_off,
                        orc_heap* heap, orc_cell* c);
#endif
