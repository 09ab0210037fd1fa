// oid_classes.h — type oid → decode class, as parse_cell_from_postgres_text dispatches
// (crates/etl/src/conversions/text.rs:28-173 after convert_type_oid_to_type / is_array_type,
// crates/etl-postgres/src/types/utils.rs:7-16).
//
// Written from PostgreSQL's catalogue (pg_type: every built-in type that has an array type, as
// `element oid, array oid` pairs) rather than from the reference's match arms, so that this table and the
// oracle's switch (oracle/oracle_cells.c) are two independent statements of the same mapping;
// tests/test_oid_classes.py checks both against a golden list extracted from text.rs itself.
#pragma once
#include <stdint.h>

#include "etl_decode.h"

namespace etl_oid {
struct PgArrayPair { uint32_t elem, array; uint8_t cls; };   // cls = ETL_K_* of the element, 0 = no typed parser (text)
// pg_type.dat, types known to postgres-types' `Type::from_oid` that have typarray != 0
static const PgArrayPair kPairs[] = {
    {16, 1000, ETL_K_BOOL},        {17, 1001, ETL_K_BYTES},      {18, 1002, ETL_K_STRING},     {19, 1003, ETL_K_STRING},
    {20, 1016, ETL_K_I64},         {21, 1005, ETL_K_I16},        {22, 1006, 0},                {23, 1007, ETL_K_I32},
    {24, 1008, 0},                 {25, 1009, ETL_K_STRING},     {26, 1028, ETL_K_U32},        {27, 1010, 0},
    {28, 1011, 0},                 {29, 1012, 0},                {30, 1013, 0},                {114, 199, ETL_K_JSON},
    {142, 143, 0},                 {600, 1017, 0},               {601, 1018, 0},               {602, 1019, 0},
    {603, 1020, 0},                {604, 1027, 0},               {628, 629, 0},                {650, 651, 0},
    {700, 1021, ETL_K_F32},        {701, 1022, ETL_K_F64},       {718, 719, 0},                {774, 775, 0},
    {790, 791, ETL_K_STRING},      {829, 1040, 0},               {869, 1041, 0},               {1033, 1034, 0},
    {1042, 1014, ETL_K_STRING},    {1043, 1015, ETL_K_STRING},   {1082, 1182, ETL_K_DATE},     {1083, 1183, ETL_K_TIME},
    {1114, 1115, ETL_K_TIMESTAMP}, {1184, 1185, ETL_K_TIMESTAMPTZ}, {1186, 1187, 0},           {1266, 1270, 0},
    {1560, 1561, 0},               {1562, 1563, 0},              {1700, 1231, ETL_K_NUMERIC},  {1790, 2201, 0},
    {2202, 2207, 0},               {2203, 2208, 0},              {2204, 2209, 0},              {2205, 2210, 0},
    {2206, 2211, 0},               {2275, 1263, 0},              {2950, 2951, ETL_K_UUID},     {2970, 2949, 0},
    {3220, 3221, 0},               {3614, 3643, 0},              {3642, 3644, 0},              {3615, 3645, 0},
    {3734, 3735, 0},               {3769, 3770, 0},              {3802, 3807, ETL_K_JSON},     {3904, 3905, 0},
    {3906, 3907, 0},               {3908, 3909, 0},              {3910, 3911, 0},              {3912, 3913, 0},
    {3926, 3927, 0},               {4072, 4073, 0},              {4089, 4090, 0},              {4096, 4097, 0},
    {4191, 4192, 0},               {4451, 6150, 0},              {4532, 6151, 0},              {4533, 6152, 0},
    {4534, 6153, 0},               {4535, 6155, 0},              {4536, 6157, 0},              {5038, 5039, 0},
    {5069, 271, 0},
};
}  // namespace etl_oid

// ETL_K_* for a column of this type; unknown oids decode as text (utils.rs:11-13 falls back to TEXT)
static inline uint32_t etl_oid_decode_class(uint32_t oid) {
  for (const etl_oid::PgArrayPair& p : etl_oid::kPairs) {
    if (p.elem == oid) return p.cls ? p.cls : (uint32_t)ETL_K_STRING;
    if (p.array == oid) return (uint32_t)ETL_K_ARRAY | (p.cls ? p.cls : (uint32_t)ETL_K_STRING);
  }
  return ETL_K_STRING;
}
