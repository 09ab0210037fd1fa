// arrow_emit.cu — columnar emitter: the rows of ONE replicated-schema version of a decoded batch as
// Arrow-layout column buffers, built on the device from the cell plane.
//
// Replaces the per-row walk of the destinations' encoders — build_array_for_field / build_primitive_array /
// build_boolean_array / build_string_array / build_binary_array / build_uuid_array and the cell_to_* converters of
// crates/etl-destinations/src/iceberg/encoding.rs:61-330 (the DuckLake and BigQuery encoders walk the same
// Vec<TableRow>) — for the column types whose Arrow value is a function of the decoded cell alone:
//   Bool → Boolean (bit-packed) · I16/I32 → Int32 · I64/U32 → Int64 (:200-221) · F32 · F64 · Date → Date32 days (:257-262)
//   Time → Time64 µs (:270-275) · Timestamp / TimestampTz → Timestamp µs (:284-301) · Uuid → FixedSizeBinary(16) (:313-318)
//   String → Utf8 (int32 offsets) · Bytes → LargeBinary (int64 offsets) (:245-250).
// A cell of another variant in such a column becomes null, exactly as the converters return None.  Numeric, Json and
// Array columns go through cell_to_string in the reference (formatting of PgNumeric / serde_json / arrays): they are
// reported as ETL_ARROW_UNSUPPORTED and stay on the shim's row path.
// Pure gather / scan / copy kernels over planes that are already in HBM: HBM-bound, no parsing.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "etl_decode.h"

namespace {

constexpr int kSelThreads = 1024;

struct SelParams {
  const uint8_t* rec_kind; const uint8_t* rec_flags; const int32_t* rec_schema; const uint64_t* rec_cell_base;
  uint64_t n_records;
  int32_t schema; uint32_t row_kinds; uint32_t n_cols;
  uint32_t* blk;          // per-block counts → exclusive offsets
  uint64_t* row_cell0;    // out: first cell of the row's image, per selected row
  uint64_t* row_rec;      // out: record index per selected row
  unsigned long long* n_rows;
};
// which image of record r is a row of this batch? returns false or the cell offset of the image inside the record
__device__ __forceinline__ bool row_of(const SelParams& S, uint64_t r, uint64_t* cell0) {
  if (r >= S.n_records || S.rec_schema[r] != S.schema) return false;
  const uint32_t k = S.rec_kind[r], f = S.rec_flags[r];
  if (!(f & ETL_RF_EVENT)) return false;
  const uint64_t c0 = S.rec_cell_base[r];
  if (k == 'I' && (S.row_kinds & 1u)) { *cell0 = c0; return true; }
  if (k == 'U' && (S.row_kinds & 2u) && !(f & ETL_RF_NEW_PARTIAL)) {      // UpdatedTableRow::Full only: a partial row has holes
    *cell0 = S.rec_cell_base[r + 1] - S.n_cols;                           // the new image is the record's last n_cols cells
    return true;
  }
  if (k == 'D' && (S.row_kinds & 4u) && (f & ETL_RF_OLD_FULL)) { *cell0 = c0; return true; }
  return false;
}
__global__ void __launch_bounds__(kSelThreads) k_sel_count(SelParams S) {
  uint64_t c0;
  const int c = __syncthreads_count(row_of(S, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, &c0));
  if (threadIdx.x == 0) S.blk[blockIdx.x] = (uint32_t)c;
}
__global__ void __launch_bounds__(kSelThreads) k_blk_scan(uint32_t* blk, uint32_t nb, unsigned long long* total) {
  __shared__ uint32_t sh[kSelThreads];
  const uint32_t per = (nb + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = threadIdx.x * per, hi = min(lo + per, nb);
  uint32_t acc = 0;
  for (uint32_t i = lo; i < hi; i++) acc += blk[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
    uint32_t v = sh[threadIdx.x];
    if (threadIdx.x >= d) v += sh[threadIdx.x - d];
    __syncthreads();
    sh[threadIdx.x] = v;
    __syncthreads();
  }
  uint32_t run = threadIdx.x ? sh[threadIdx.x - 1] : 0u;
  for (uint32_t i = lo; i < hi; i++) { const uint32_t c = blk[i]; blk[i] = run; run += c; }
  if (threadIdx.x == blockDim.x - 1) *total = sh[blockDim.x - 1];
}
__global__ void __launch_bounds__(kSelThreads) k_sel_scatter(SelParams S) {
  __shared__ uint32_t warp_cnt[kSelThreads / 32];
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t c0 = 0;
  const bool sel = row_of(S, r, &c0);
  const unsigned bal = __ballot_sync(0xffffffffu, sel);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[wid] = __popc(bal);
  __syncthreads();
  uint32_t before = 0;
  for (int k = 0; k < wid; k++) before += warp_cnt[k];
  if (sel) {
    const uint64_t at = (uint64_t)S.blk[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1u));
    S.row_cell0[at] = c0; S.row_rec[at] = r;
  }
}

struct ColParams {
  const uint8_t* cell_tag; const uint64_t* cell_val; const uint32_t* cell_aux; const uint8_t* heap; const uint8_t* stream;
  const uint64_t* row_cell0; uint64_t n_rows; uint32_t col; uint32_t arrow_type;
  uint32_t* validity;     // one word per 32 rows
  void* values;           // fixed width
  uint32_t* lens;         // var width: byte length per row (→ scanned into offsets)
};
// iceberg/encoding.rs:200-318: value of a cell for the column's Arrow type, or "null"
__device__ __forceinline__ bool fixed_value(uint32_t at, uint32_t tag, uint64_t val, uint32_t aux, int64_t* out) {
  switch (at) {
    case ETL_ARROW_BOOLEAN: if (tag != ETL_CELL_BOOL) return false; *out = (int64_t)(val & 1u); return true;
    case ETL_ARROW_INT32: if (tag != ETL_CELL_I16 && tag != ETL_CELL_I32) return false; *out = (int64_t)val; return true;
    case ETL_ARROW_INT64: if (tag != ETL_CELL_I64 && tag != ETL_CELL_U32) return false; *out = tag == ETL_CELL_U32 ? (int64_t)(uint32_t)val : (int64_t)val; return true;
    case ETL_ARROW_FLOAT32: if (tag != ETL_CELL_F32) return false; *out = (int64_t)(uint32_t)val; return true;
    case ETL_ARROW_FLOAT64: if (tag != ETL_CELL_F64) return false; *out = (int64_t)val; return true;
    case ETL_ARROW_DATE32: if (tag != ETL_CELL_DATE) return false; *out = (int64_t)val; return true;
    case ETL_ARROW_TIME64_US: if (tag != ETL_CELL_TIME) return false; *out = (int64_t)val * 1000000ll + (int64_t)(aux / 1000u); return true;
    case ETL_ARROW_TIMESTAMP_US: if (tag != ETL_CELL_TIMESTAMP) return false; *out = (int64_t)val * 1000000ll + (int64_t)(aux / 1000u); return true;
    case ETL_ARROW_TIMESTAMPTZ_US: if (tag != ETL_CELL_TIMESTAMPTZ) return false; *out = (int64_t)val * 1000000ll + (int64_t)(aux / 1000u); return true;
    default: return false;
  }
}
// one thread per row of one column: value / length + the validity word of its warp
__global__ void __launch_bounds__(256) k_col_fixed(ColParams C) {
  const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = false;
  int64_t v = 0;
  uint32_t len = 0;
  if (row < C.n_rows) {
    const uint64_t cell = C.row_cell0[row] + C.col;
    const uint32_t tag = C.cell_tag[cell];
    const uint64_t val = C.cell_val[cell];
    const uint32_t aux = C.cell_aux[cell];
    switch (C.arrow_type) {
      case ETL_ARROW_UTF8: valid = tag == ETL_CELL_STRING; len = valid ? aux : 0u; break;
      case ETL_ARROW_LARGE_BINARY: valid = tag == ETL_CELL_BYTES; len = valid ? aux : 0u; break;
      case ETL_ARROW_UUID: valid = tag == ETL_CELL_UUID; break;
      default: valid = fixed_value(C.arrow_type, tag, val, aux, &v); break;
    }
    switch (C.arrow_type) {
      case ETL_ARROW_INT32: case ETL_ARROW_DATE32: case ETL_ARROW_FLOAT32: static_cast<int32_t*>(C.values)[row] = valid ? (int32_t)v : 0; break;
      case ETL_ARROW_INT64: case ETL_ARROW_FLOAT64: case ETL_ARROW_TIME64_US: case ETL_ARROW_TIMESTAMP_US: case ETL_ARROW_TIMESTAMPTZ_US:
        static_cast<int64_t*>(C.values)[row] = valid ? v : 0; break;
      case ETL_ARROW_UUID: {
        uint64_t a = 0, b = 0;
        if (valid) { const uint64_t* s = reinterpret_cast<const uint64_t*>(C.heap + val); a = s[0]; b = s[1]; }   // heap reservations are 8-byte aligned
        static_cast<uint64_t*>(C.values)[2 * row] = a; static_cast<uint64_t*>(C.values)[2 * row + 1] = b;
        break;
      }
      case ETL_ARROW_UTF8: case ETL_ARROW_LARGE_BINARY: C.lens[row] = len; break;
      default: break;
    }
  }
  const unsigned vb = __ballot_sync(0xffffffffu, valid);
  const unsigned bb = __ballot_sync(0xffffffffu, valid && v != 0);
  if ((threadIdx.x & 31) == 0 && (row >> 5) < ((C.n_rows + 31) >> 5)) {
    C.validity[row >> 5] = vb;
    if (C.arrow_type == ETL_ARROW_BOOLEAN) static_cast<uint32_t*>(C.values)[row >> 5] = bb;   // Boolean values are bit-packed too
  }
}
// exclusive scan of lens → int32 / int64 offsets (n + 1 entries): block sums, scan of the sums, final pass
__global__ void __launch_bounds__(1024) k_len_blocks(const uint32_t* lens, uint64_t n, unsigned long long* blk) {
  __shared__ unsigned long long sh[32];
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = i < n ? lens[i] : 0ull;
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = sh[threadIdx.x];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (threadIdx.x == 0) blk[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(1024) k_blk_scan64(unsigned long long* blk, uint32_t nb, unsigned long long* total) {
  __shared__ unsigned long long sh[1024];
  const uint32_t per = (nb + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = threadIdx.x * per, hi = min(lo + per, nb);
  unsigned long long acc = 0;
  for (uint32_t i = lo; i < hi; i++) acc += blk[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
    unsigned long long v = sh[threadIdx.x];
    if (threadIdx.x >= d) v += sh[threadIdx.x - d];
    __syncthreads();
    sh[threadIdx.x] = v;
    __syncthreads();
  }
  unsigned long long run = threadIdx.x ? sh[threadIdx.x - 1] : 0ull;
  for (uint32_t i = lo; i < hi; i++) { const unsigned long long c = blk[i]; blk[i] = run; run += c; }
  if (threadIdx.x == blockDim.x - 1) *total = sh[blockDim.x - 1];
}
template <typename OffT>
__global__ void __launch_bounds__(1024) k_offsets(const uint32_t* lens, uint64_t n, const unsigned long long* blk, OffT* offs) {
  __shared__ unsigned long long sh[1024];
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long mine = i < n ? lens[i] : 0ull;
  sh[threadIdx.x] = mine;
  __syncthreads();
  for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
    unsigned long long v = sh[threadIdx.x];
    if (threadIdx.x >= d) v += sh[threadIdx.x - d];
    __syncthreads();
    sh[threadIdx.x] = v;
    __syncthreads();
  }
  const unsigned long long excl = blk[blockIdx.x] + sh[threadIdx.x] - mine;
  if (i < n) offs[i] = (OffT)excl;
  if (i + 1 == n) offs[n] = (OffT)(excl + mine);
}
// warp per row: copy the row's bytes to their place in the column's data buffer (16 B per lane where aligned)
template <typename OffT>
__global__ void __launch_bounds__(256) k_gather(ColParams C, const OffT* offs, uint8_t* data) {
  const uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= C.n_rows) return;
  const uint64_t cell = C.row_cell0[row] + C.col;
  const uint32_t tag = C.cell_tag[cell];
  const bool str = C.arrow_type == ETL_ARROW_UTF8;
  if (tag != (str ? (uint32_t)ETL_CELL_STRING : (uint32_t)ETL_CELL_BYTES)) return;
  const uint8_t* src = (str ? C.stream : C.heap) + C.cell_val[cell];
  const uint32_t n = C.cell_aux[cell];
  uint8_t* dst = data + (uint64_t)offs[row];
  for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}

uint32_t arrow_type_of(uint32_t k) {
  switch (k) {
    case ETL_K_BOOL: return ETL_ARROW_BOOLEAN;
    case ETL_K_I16: case ETL_K_I32: return ETL_ARROW_INT32;
    case ETL_K_I64: case ETL_K_U32: return ETL_ARROW_INT64;
    case ETL_K_F32: return ETL_ARROW_FLOAT32;
    case ETL_K_F64: return ETL_ARROW_FLOAT64;
    case ETL_K_STRING: return ETL_ARROW_UTF8;
    case ETL_K_BYTES: return ETL_ARROW_LARGE_BINARY;
    case ETL_K_DATE: return ETL_ARROW_DATE32;
    case ETL_K_TIME: return ETL_ARROW_TIME64_US;
    case ETL_K_TIMESTAMP: return ETL_ARROW_TIMESTAMP_US;
    case ETL_K_TIMESTAMPTZ: return ETL_ARROW_TIMESTAMPTZ_US;
    case ETL_K_UUID: return ETL_ARROW_UUID;
    default: return ETL_ARROW_UNSUPPORTED;    // numeric / json / arrays: cell_to_string formatting stays with the shim
  }
}
uint32_t value_width(uint32_t at) {
  switch (at) {
    case ETL_ARROW_INT32: case ETL_ARROW_DATE32: case ETL_ARROW_FLOAT32: return 4;
    case ETL_ARROW_INT64: case ETL_ARROW_FLOAT64: case ETL_ARROW_TIME64_US: case ETL_ARROW_TIMESTAMP_US: case ETL_ARROW_TIMESTAMPTZ_US: return 8;
    case ETL_ARROW_UUID: return 16;
    default: return 0;
  }
}
struct Col {
  uint32_t arrow_type = 0;
  uint64_t validity_off = 0, values_off = 0, offsets_off = 0, data_off = 0, data_bytes = 0, values_bytes = 0, offsets_bytes = 0;
};

}  // namespace

struct etl_arrow_batch {
  uint64_t n_rows = 0;
  std::vector<Col> cols;
  uint8_t* dev = nullptr;     // one device allocation: row_rec | per column validity, values / offsets, data
  uint8_t* host = nullptr;    // pinned host image (to_host)
  uint64_t bytes = 0, row_rec_off = 0;
  std::string error;
};

extern "C" {

int etl_dec_arrow_emit(const etl_dec_batch* batch, uint32_t schema_index, uint32_t row_kinds, int to_host, etl_arrow_batch** out) {
  if (!batch || !out) return ETL_ERR_INVALID_ARG;
  const uint8_t* dev_stream = etl_dec_batch_device_stream(batch);
  etl_dec_planes P;
  etl_dec_summary S;
  etl_dec_schema_info sc;
  if (etl_dec_batch_planes(batch, 0, &P) != ETL_OK || etl_dec_batch_summary(batch, &S) != ETL_OK) return ETL_ERR_INVALID_ARG;
  if (etl_dec_batch_schema(batch, schema_index, &sc) != ETL_OK) return ETL_ERR_INVALID_ARG;
  cudaStream_t st = cudaStreamPerThread;
  etl_arrow_batch* A = new etl_arrow_batch();
  auto fail = [&](int rc) { if (A->dev) cudaFree(A->dev); if (A->host) cudaFreeHost(A->host); delete A; return rc; };
#define CKA(call) do { if ((call) != cudaSuccess) { cudaGetLastError(); return fail(ETL_ERR_CUDA); } } while (0)
  // rows of the valid prefix only
  const uint64_t n_valid = S.first_error.record_index == UINT64_MAX ? P.n_records : std::min<uint64_t>(P.n_records, S.first_error.record_index - S.record_index_base);
  const uint32_t nb = (uint32_t)((n_valid + kSelThreads - 1) / kSelThreads);
  uint32_t* d_blk = nullptr; uint64_t* d_cell0 = nullptr; uint64_t* d_rec = nullptr; unsigned long long* d_n = nullptr;
  CKA(cudaMalloc(&d_blk, (nb + 1) * 4ull)); CKA(cudaMalloc(&d_cell0, (n_valid + 1) * 8)); CKA(cudaMalloc(&d_rec, (n_valid + 1) * 8)); CKA(cudaMalloc(&d_n, 16));
  auto free_tmp = [&]() { cudaFree(d_blk); cudaFree(d_cell0); cudaFree(d_rec); cudaFree(d_n); };
  SelParams Sp{P.rec_kind, P.rec_flags, P.rec_schema, P.rec_cell_base, n_valid, (int32_t)schema_index, row_kinds, sc.n_cols, d_blk, d_cell0, d_rec, d_n};
  unsigned long long n_rows = 0;
  if (nb) {
    k_sel_count<<<nb, kSelThreads, 0, st>>>(Sp);
    k_blk_scan<<<1, kSelThreads, 0, st>>>(d_blk, nb, d_n);
    k_sel_scatter<<<nb, kSelThreads, 0, st>>>(Sp);
    if (cudaMemcpyAsync(&n_rows, d_n, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { free_tmp(); return fail(ETL_ERR_CUDA); }
  }
  A->n_rows = n_rows;
  // pass 1: validity + fixed values + lengths (into a scratch), per column; var-width sizes need a sync before the data buffers exist
  const uint64_t vbytes = ((n_rows + 31) / 32 * 4 + 63) & ~63ull;
  uint64_t cur = ((n_rows * 8) + 63) & ~63ull;        // row_rec first
  A->row_rec_off = 0;
  A->cols.resize(sc.n_cols);
  uint32_t n_var = 0;
  for (uint32_t c = 0; c < sc.n_cols; c++) {
    Col& col = A->cols[c];
    col.arrow_type = arrow_type_of(sc.col_kind[c]);
    if (col.arrow_type == ETL_ARROW_UNSUPPORTED) continue;
    col.validity_off = cur; cur += vbytes;
    if (col.arrow_type == ETL_ARROW_BOOLEAN) { col.values_off = cur; col.values_bytes = vbytes; cur += vbytes; }
    else if (value_width(col.arrow_type)) { col.values_off = cur; col.values_bytes = (n_rows * value_width(col.arrow_type) + 63) & ~63ull; cur += col.values_bytes; }
    else { col.offsets_off = cur; col.offsets_bytes = ((n_rows + 1) * (col.arrow_type == ETL_ARROW_UTF8 ? 4 : 8) + 63) & ~63ull; cur += col.offsets_bytes; n_var++; }
  }
  const uint64_t fixed_bytes = cur;
  uint32_t* d_lens = nullptr; unsigned long long* d_lblk = nullptr; unsigned long long* d_tot = nullptr;
  const uint32_t lb = (uint32_t)((n_rows + 1023) / 1024);
  if (n_var) { if (cudaMalloc(&d_lens, (size_t)n_var * (n_rows + 1) * 4) != cudaSuccess || cudaMalloc(&d_lblk, (size_t)n_var * (lb + 1) * 8) != cudaSuccess || cudaMalloc(&d_tot, n_var * 8 + 8) != cudaSuccess) { free_tmp(); return fail(ETL_ERR_ALLOC); } }
  uint8_t* d_fixed = nullptr;
  if (cudaMalloc(&d_fixed, fixed_bytes + 64) != cudaSuccess) { free_tmp(); cudaFree(d_lens); cudaFree(d_lblk); cudaFree(d_tot); return fail(ETL_ERR_ALLOC); }
  cudaMemsetAsync(d_fixed, 0, fixed_bytes + 64, st);
  if (n_rows) cudaMemcpyAsync(d_fixed, d_rec, n_rows * 8, cudaMemcpyDeviceToDevice, st);
  std::vector<unsigned long long> totals(n_var + 1, 0);
  uint32_t vi = 0;
  for (uint32_t c = 0; c < sc.n_cols && n_rows; c++) {
    Col& col = A->cols[c];
    if (col.arrow_type == ETL_ARROW_UNSUPPORTED) continue;
    ColParams Cp{P.cell_tag, P.cell_val, P.cell_aux, P.heap, dev_stream, d_cell0, n_rows, c, col.arrow_type,
                 (uint32_t*)(d_fixed + col.validity_off), col.values_bytes ? (void*)(d_fixed + col.values_off) : nullptr, nullptr};
    const bool var = col.arrow_type == ETL_ARROW_UTF8 || col.arrow_type == ETL_ARROW_LARGE_BINARY;
    if (var) Cp.lens = d_lens + (size_t)vi * (n_rows + 1);
    k_col_fixed<<<(uint32_t)((n_rows + 255) / 256), 256, 0, st>>>(Cp);
    if (var) {
      k_len_blocks<<<lb, 1024, 0, st>>>(Cp.lens, n_rows, d_lblk + (size_t)vi * (lb + 1));
      k_blk_scan64<<<1, 1024, 0, st>>>(d_lblk + (size_t)vi * (lb + 1), lb, d_tot + vi);
      if (col.arrow_type == ETL_ARROW_UTF8) k_offsets<int32_t><<<lb, 1024, 0, st>>>(Cp.lens, n_rows, d_lblk + (size_t)vi * (lb + 1), (int32_t*)(d_fixed + col.offsets_off));
      else k_offsets<int64_t><<<lb, 1024, 0, st>>>(Cp.lens, n_rows, d_lblk + (size_t)vi * (lb + 1), (int64_t*)(d_fixed + col.offsets_off));
      vi++;
    }
  }
  if (n_var && n_rows) cudaMemcpyAsync(totals.data(), d_tot, n_var * 8, cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { free_tmp(); cudaFree(d_lens); cudaFree(d_lblk); cudaFree(d_tot); cudaFree(d_fixed); return fail(ETL_ERR_CUDA); }
  // pass 2: data buffers
  vi = 0;
  for (uint32_t c = 0; c < sc.n_cols; c++) {
    Col& col = A->cols[c];
    if (col.arrow_type != ETL_ARROW_UTF8 && col.arrow_type != ETL_ARROW_LARGE_BINARY) continue;
    col.data_off = cur; col.data_bytes = n_rows ? totals[vi] : 0; cur += (col.data_bytes + 63) & ~63ull;
    if (col.arrow_type == ETL_ARROW_UTF8 && col.data_bytes > 0x7FFFFFFFull) { A->error = "Utf8 column exceeds 2 GiB: split the batch"; }
    vi++;
  }
  A->bytes = cur + 64;
  bool ok = A->error.empty() && cudaMalloc(&A->dev, A->bytes) == cudaSuccess;
  if (ok) ok = cudaMemcpyAsync(A->dev, d_fixed, fixed_bytes, cudaMemcpyDeviceToDevice, st) == cudaSuccess;
  for (uint32_t c = 0; ok && c < sc.n_cols && n_rows; c++) {
    Col& col = A->cols[c];
    if (col.arrow_type != ETL_ARROW_UTF8 && col.arrow_type != ETL_ARROW_LARGE_BINARY) continue;
    ColParams Cp{P.cell_tag, P.cell_val, P.cell_aux, P.heap, dev_stream, d_cell0, n_rows, c, col.arrow_type, nullptr, nullptr, nullptr};
    const uint32_t grid = (uint32_t)((n_rows * 32 + 255) / 256);
    if (col.arrow_type == ETL_ARROW_UTF8) k_gather<int32_t><<<grid, 256, 0, st>>>(Cp, (const int32_t*)(A->dev + col.offsets_off), A->dev + col.data_off);
    else k_gather<int64_t><<<grid, 256, 0, st>>>(Cp, (const int64_t*)(A->dev + col.offsets_off), A->dev + col.data_off);
  }
  if (ok && to_host) {
    ok = cudaHostAlloc((void**)&A->host, A->bytes, cudaHostAllocDefault) == cudaSuccess;
    if (ok) ok = cudaMemcpyAsync(A->host, A->dev, A->bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  }
  if (ok) ok = cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  free_tmp(); cudaFree(d_lens); cudaFree(d_lblk); cudaFree(d_tot); cudaFree(d_fixed);
  if (!ok) return fail(A->error.empty() ? ETL_ERR_CUDA : ETL_ERR_INVALID_ARG);
  *out = A;
  return ETL_OK;
#undef CKA
}
uint64_t etl_dec_arrow_rows(const etl_arrow_batch* a) { return a ? a->n_rows : 0; }
uint32_t etl_dec_arrow_cols(const etl_arrow_batch* a) { return a ? (uint32_t)a->cols.size() : 0; }
const uint64_t* etl_dec_arrow_row_records(const etl_arrow_batch* a, int host) {
  if (!a) return nullptr;
  const uint8_t* base = host ? a->host : a->dev;
  return base ? reinterpret_cast<const uint64_t*>(base + a->row_rec_off) : nullptr;
}
int etl_dec_arrow_column(const etl_arrow_batch* a, uint32_t c, int host, etl_arrow_column* out) {
  if (!a || !out || c >= a->cols.size()) return ETL_ERR_INVALID_ARG;
  const uint8_t* base = host ? a->host : a->dev;
  if (!base) return ETL_ERR_INVALID_ARG;
  const Col& col = a->cols[c];
  memset(out, 0, sizeof *out);
  out->arrow_type = col.arrow_type;
  if (col.arrow_type == ETL_ARROW_UNSUPPORTED) return ETL_OK;
  out->validity = base + col.validity_off;
  if (col.values_bytes) out->values = base + col.values_off;
  if (col.offsets_bytes) { out->offsets = base + col.offsets_off; out->data = base + col.data_off; out->data_bytes = col.data_bytes; }
  return ETL_OK;
}
void etl_dec_arrow_free(etl_arrow_batch* a) {
  if (!a) return;
  if (a->dev) cudaFree(a->dev);
  if (a->host) cudaFreeHost(a->host);
  delete a;
}

}  // extern "C"
