// float_parse.cuh — correctly rounded decimal → f32 / f64 on the device (text.rs:61-68).
//
// The reference calls Rust's core::num::dec2flt (str::parse::<f32/f64>): the grammar below, then
// Eisel-Lemire with the 128-bit power-of-five table (always exact for ≤ 19 significant digits —
// Mushtak & Lemire, "Fast number parsing without fallback"), and for longer inputs the bracket
// w / w+1 with Nigel Tao's exact "simple decimal conversion" as the tie-breaker.  Same algorithm
// here, integer arithmetic only (the tables are generated exactly by tools/gen_float_tables.py).
#pragma once
#include <stdint.h>

#include "cell_parsers.cuh"
#include "float_tables.cuh"

namespace etl {

struct FloatFmt {
  int mant_bits;        // explicit mantissa bits: 52 / 23
  int min_exp;          // -1023 / -127
  int inf_power;        // 0x7FF / 0xFF
  int min_rte, max_rte; // round-to-even exponent window: [-4, 23] / [-17, 10]
  int smallest_q, largest_q;  // [-342, 308] / [-64, 38]
};
__device__ __forceinline__ FloatFmt fmt_of(bool f32) {
  return f32 ? FloatFmt{23, -127, 0xFF, -17, 10, -64, 38} : FloatFmt{52, -1023, 0x7FF, -4, 23, -342, 308};
}
struct AdjMant { uint64_t mant; int32_t pow2; };

// Eisel-Lemire (fast_float compute_float / Rust dec2flt::lemire::compute_float)
__device__ __noinline__ AdjMant lemire(int64_t q, uint64_t w, const FloatFmt& F) {
  AdjMant zero{0, 0}, inf{0, F.inf_power};
  if (w == 0 || q < F.smallest_q) return zero;
  if (q > F.largest_q) return inf;
  const int lz = __clzll((long long)w);
  w <<= lz;
  const uint64_t thi = kPow5Table[q - kPow5MinQ][0], tlo = kPow5Table[q - kPow5MinQ][1];
  const int precision = F.mant_bits + 3;
  const uint64_t mask = precision < 64 ? (0xFFFFFFFFFFFFFFFFull >> precision) : 0xFFFFFFFFFFFFFFFFull;
  uint64_t lo = w * thi, hi = __umul64hi(w, thi);
  if ((hi & mask) == mask) {
    const uint64_t shi = __umul64hi(w, tlo);
    lo += shi;
    if (shi > lo) hi++;
  }
  const int upperbit = (int)(hi >> 63);
  const int shift = upperbit + 64 - F.mant_bits - 3;
  uint64_t mant = hi >> shift;
  int32_t pow2 = (int32_t)((((int64_t)(152170 + 65536) * q) >> 16) + 63) + upperbit - lz - F.min_exp;
  if (pow2 <= 0) {  // subnormal
    if (-pow2 + 1 >= 64) return zero;
    mant >>= (-pow2 + 1);
    mant += (mant & 1);
    mant >>= 1;
    pow2 = (mant < (1ull << F.mant_bits)) ? 0 : 1;
    return AdjMant{mant, pow2};
  }
  if (lo <= 1 && q >= F.min_rte && q <= F.max_rte && (mant & 3) == 1 && (mant << shift) == hi) mant &= ~1ull;
  mant += (mant & 1);
  mant >>= 1;
  if (mant >= (2ull << F.mant_bits)) { mant = 1ull << F.mant_bits; pow2++; }
  mant &= ~(1ull << F.mant_bits);
  if (pow2 >= F.inf_power) return inf;
  return AdjMant{mant, pow2};
}

// ---- exact fallback: simple decimal conversion (Rust dec2flt::decimal + slow.rs)
constexpr int kDecMaxDigits = 768;
struct Decimal {
  uint32_t num_digits;
  int32_t decimal_point;
  bool truncated;
  uint8_t digits[kDecMaxDigits];
};
__device__ __forceinline__ void dec_trim(Decimal& d) { while (d.num_digits != 0 && d.digits[d.num_digits - 1] == 0) d.num_digits--; }
__device__ __noinline__ uint32_t dec_new_digits_left_shift(const Decimal& d, uint32_t shift) {
  shift &= 63;
  const uint32_t num_new = kPow2Digits[shift];
  const uint32_t a = kPow5DigitOff[shift], b = kPow5DigitOff[shift + 1];
  for (uint32_t i = 0; i < b - a; i++) {
    const uint32_t p5 = kPow5Digits[a + i];
    if (i >= d.num_digits) return num_new - 1;
    if (d.digits[i] == p5) continue;
    return d.digits[i] < p5 ? num_new - 1 : num_new;
  }
  return num_new;
}
__device__ __noinline__ void dec_left_shift(Decimal& d, uint32_t shift) {
  if (d.num_digits == 0) return;
  const uint32_t num_new = dec_new_digits_left_shift(d, shift);
  uint32_t read = d.num_digits, write = d.num_digits + num_new;
  uint64_t n = 0;
  while (read != 0) {
    read--; write--;
    n += (uint64_t)d.digits[read] << shift;
    const uint64_t q = n / 10, r = n - 10 * q;
    if (write < (uint32_t)kDecMaxDigits) d.digits[write] = (uint8_t)r; else if (r > 0) d.truncated = true;
    n = q;
  }
  while (n > 0) {
    write--;
    const uint64_t q = n / 10, r = n - 10 * q;
    if (write < (uint32_t)kDecMaxDigits) d.digits[write] = (uint8_t)r; else if (r > 0) d.truncated = true;
    n = q;
  }
  d.num_digits += num_new;
  if (d.num_digits > (uint32_t)kDecMaxDigits) d.num_digits = kDecMaxDigits;
  d.decimal_point += (int32_t)num_new;
  dec_trim(d);
}
__device__ __noinline__ void dec_right_shift(Decimal& d, uint32_t shift) {
  uint32_t read = 0, write = 0;
  uint64_t n = 0;
  while ((n >> shift) == 0) {
    if (read < d.num_digits) { n = 10 * n + d.digits[read]; read++; }
    else if (n == 0) return;
    else { while ((n >> shift) == 0) { n *= 10; read++; } break; }
  }
  d.decimal_point -= (int32_t)read - 1;
  if (d.decimal_point < -2047) { d.num_digits = 0; d.decimal_point = 0; d.truncated = false; return; }
  const uint64_t mask = (1ull << shift) - 1;
  while (read < d.num_digits) {
    const uint8_t nd = (uint8_t)(n >> shift);
    n = 10 * (n & mask) + d.digits[read];
    read++;
    d.digits[write++] = nd;
  }
  while (n > 0) {
    const uint8_t nd = (uint8_t)(n >> shift);
    n = 10 * (n & mask);
    if (write < (uint32_t)kDecMaxDigits) d.digits[write++] = nd; else if (nd > 0) d.truncated = true;
  }
  d.num_digits = write;
  dec_trim(d);
}
__device__ __noinline__ uint64_t dec_round(const Decimal& d) {
  if (d.num_digits == 0 || d.decimal_point < 0) return 0;
  if (d.decimal_point > 18) return 0xFFFFFFFFFFFFFFFFull;
  const uint32_t dp = (uint32_t)d.decimal_point;
  uint64_t n = 0;
  for (uint32_t i = 0; i < dp; i++) { n *= 10; if (i < d.num_digits) n += d.digits[i]; }
  bool up = false;
  if (dp < d.num_digits) {
    up = d.digits[dp] >= 5;
    if (d.digits[dp] == 5 && dp + 1 == d.num_digits) up = d.truncated || (dp != 0 && (d.digits[dp - 1] & 1));
  }
  return up ? n + 1 : n;
}
__device__ __noinline__ AdjMant dec_to_binary(Decimal& d, const FloatFmt& F) {
  const AdjMant zero{0, 0}, inf{0, F.inf_power};
  if (d.num_digits == 0 || d.decimal_point < -324) return zero;
  if (d.decimal_point >= 310) return inf;
  const uint8_t powers[19] = {0, 3, 6, 9, 13, 16, 19, 23, 26, 29, 33, 36, 39, 43, 46, 49, 53, 56, 59};
  int32_t exp2 = 0;
  while (d.decimal_point > 0) {
    const uint32_t nn = (uint32_t)d.decimal_point;
    const uint32_t shift = nn < 19 ? powers[nn] : 60;
    dec_right_shift(d, shift);
    if (d.decimal_point < -2047) return zero;
    exp2 += (int32_t)shift;
  }
  while (d.decimal_point <= 0) {
    uint32_t shift;
    if (d.decimal_point == 0) {
      const uint8_t d0 = d.digits[0];
      if (d0 >= 5) break;
      shift = (d0 < 2) ? 2 : 1;
    } else {
      const uint32_t nn = (uint32_t)(-d.decimal_point);
      shift = nn < 19 ? powers[nn] : 60;
    }
    dec_left_shift(d, shift);
    if (d.decimal_point > 2047) return inf;
    exp2 -= (int32_t)shift;
  }
  exp2 -= 1;
  const int32_t min_exp = F.min_exp + 1;   // -1022 / -126
  while (min_exp > exp2) {
    uint32_t nn = (uint32_t)(min_exp - exp2);
    if (nn > 60) nn = 60;
    dec_right_shift(d, nn);
    exp2 += (int32_t)nn;
  }
  if (exp2 - F.min_exp >= F.inf_power) return inf;
  dec_left_shift(d, (uint32_t)F.mant_bits + 1);
  uint64_t mant = dec_round(d);
  if (mant >= (1ull << (F.mant_bits + 1))) {
    dec_right_shift(d, 1);
    exp2 += 1;
    mant = dec_round(d);
    if (exp2 - F.min_exp >= F.inf_power) return inf;
  }
  int32_t pow2 = exp2 - F.min_exp;
  if (mant < (1ull << F.mant_bits)) pow2 -= 1;
  mant &= (1ull << F.mant_bits) - 1;
  return AdjMant{mant, pow2};
}

// Rust FromStr for f32 / f64.  Returns 0 or ETL_E_PARSE_FLOAT; o.val = IEEE bits.
__device__ __noinline__ uint32_t parse_float(const uint8_t* s, uint32_t n, bool f32, CellOut& o) {
  o.tag = f32 ? ETL_CELL_F32 : ETL_CELL_F64; o.aux = 0;
  const FloatFmt F = fmt_of(f32);
  if (n == 0) return ETL_E_PARSE_FLOAT;
  uint32_t i = 0;
  bool neg = false;
  if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
  if (i == n) return ETL_E_PARSE_FLOAT;
  const uint64_t sign = neg ? (f32 ? 0x80000000ull : 0x8000000000000000ull) : 0ull;
  // ---- grammar
  const uint32_t int0 = i;
  while (i < n && is_digit(s[i])) i++;
  const uint32_t n_int = i - int0;
  uint32_t frac0 = i, n_frac = 0;
  if (i < n && s[i] == '.') { i++; frac0 = i; while (i < n && is_digit(s[i])) i++; n_frac = i - frac0; }
  if (n_int + n_frac == 0) {
    const uint8_t* t = s + int0; const uint32_t tn = n - int0;
    if (i != int0) return ETL_E_PARSE_FLOAT;               // "." alone
    if (ieq(t, tn, "inf", 3) || ieq(t, tn, "infinity", 8)) { o.val = sign | (f32 ? 0x7F800000ull : 0x7FF0000000000000ull); return 0; }
    if (ieq(t, tn, "nan", 3)) { o.val = sign | (f32 ? 0x7FC00000ull : 0x7FF8000000000000ull); return 0; }
    return ETL_E_PARSE_FLOAT;
  }
  int64_t exp_number = 0;
  if (i < n && (s[i] == 'e' || s[i] == 'E')) {
    i++;
    bool eneg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) { eneg = s[i] == '-'; i++; }
    if (!(i < n && is_digit(s[i]))) return ETL_E_PARSE_FLOAT;
    while (i < n && is_digit(s[i])) { if (exp_number < 0x10000) exp_number = 10 * exp_number + (s[i] - '0'); i++; }
    if (eneg) exp_number = -exp_number;
  }
  if (i != n) return ETL_E_PARSE_FLOAT;
  // ---- first 19 significant digits
  uint64_t w = 0;
  uint32_t taken = 0, sig = 0;   // sig = significant digits seen (after leading zeros)
  for (uint32_t k = 0; k < n_int + n_frac; k++) {
    const uint32_t c = (k < n_int) ? s[int0 + k] : s[frac0 + (k - n_int)];
    const uint32_t dgt = c - '0';
    if (sig == 0 && dgt == 0) continue;
    sig++;
    if (taken < 19) { w = w * 10 + dgt; taken++; }   // digits past 19: sig > taken sends the value to the w / w+1 check below
  }
  if (sig == 0) { o.val = sign; return 0; }                 // ±0
  const int64_t q = exp_number - (int64_t)n_frac + (int64_t)(sig - taken);
  AdjMant am = lemire(q, w, F);
  if (sig > 19) {
    const AdjMant am2 = lemire(q, w + 1, F);
    if (am.mant != am2.mant || am.pow2 != am2.pow2) {
      // exact tie-break (only when the 19-digit bracket straddles a rounding boundary)
      Decimal d;
      d.num_digits = 0; d.decimal_point = 0; d.truncated = false;
      bool seen = false;
      int32_t lead_zeros = 0;
      for (uint32_t k = 0; k < n_int + n_frac; k++) {
        const uint32_t c = (k < n_int) ? s[int0 + k] : s[frac0 + (k - n_int)];
        const uint32_t dgt = c - '0';
        if (!seen && dgt == 0) { lead_zeros++; continue; }
        seen = true;
        if (d.num_digits < (uint32_t)kDecMaxDigits) d.digits[d.num_digits++] = (uint8_t)dgt;
        else if (dgt) d.truncated = true;
      }
      d.decimal_point = (int32_t)n_int - lead_zeros;
      int64_t dp = (int64_t)d.decimal_point + exp_number;
      if (dp > 100000) dp = 100000; if (dp < -100000) dp = -100000;
      d.decimal_point = (int32_t)dp;
      dec_trim(d);
      am = dec_to_binary(d, F);
    }
  }
  o.val = sign | ((uint64_t)am.pow2 << F.mant_bits) | am.mant;
  return 0;
}

}  // namespace etl
