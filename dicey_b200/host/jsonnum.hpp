// jsonnum.hpp -- how nlohmann::json 3.5.0 prints a double (the Tm / MatchTm / Penalty fields of the
// `dicey search` output, silica.h:139-176): the Grisu2 algorithm (Loitsch, PLDI 2010) with
// alpha = -60, gamma = -32 and cached powers every 8 decimal exponents, then plain decimal notation
// for decimal exponents in (-4, 15] and d.ddde[+-]XX otherwise; integers get a trailing ".0".
// Grisu2 is not always the shortest representation, so a shortest-digits printer (std::to_chars)
// would differ in rare cases; tests/golden/jsonfloat.* pins 31 k values against nlohmann itself.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

namespace dhost {
namespace grisu {

struct Fp {
  uint64_t f;
  int e;
};
struct Pow10 {
  uint64_t f;
  int e, k;
};
static const Pow10 kPow10[] = {
#include "pow10_table.inc"
};

inline Fp mul(const Fp& x, const Fp& y) {   // upper 64 bits of the 128-bit product, rounded to nearest
  const unsigned __int128 p = (unsigned __int128)x.f * y.f + ((unsigned __int128)1 << 63);
  return Fp{(uint64_t)(p >> 64), x.e + y.e + 64};
}
inline Fp normalize(Fp x) {
  while ((x.f >> 63) == 0) { x.f <<= 1; --x.e; }
  return x;
}
inline Fp align_to(const Fp& x, int e) { return Fp{x.f << (x.e - e), e}; }

inline Pow10 cached_power(int e) {   // alpha <= c.e + e + 64 <= gamma
  const int f = -60 - e - 1;
  const int k = (f * 78913) / (1 << 18) + static_cast<int>(f > 0);
  const int index = (300 + k + 7) / 8;
  return kPow10[index];
}

inline void round_weed(char* buf, int len, uint64_t dist, uint64_t delta, uint64_t rest, uint64_t ten_k) {
  while (rest < dist && delta - rest >= ten_k && (rest + ten_k < dist || dist - rest > rest + ten_k - dist)) {
    --buf[len - 1];
    rest += ten_k;
  }
}

inline void digit_gen(char* buffer, int& length, int& decimal_exponent, Fp M_minus, Fp w, Fp M_plus) {
  uint64_t delta = M_plus.f - M_minus.f;
  uint64_t dist = M_plus.f - w.f;
  const Fp one{uint64_t{1} << -M_plus.e, M_plus.e};
  uint32_t p1 = static_cast<uint32_t>(M_plus.f >> -one.e);
  uint64_t p2 = M_plus.f & (one.f - 1);
  uint32_t pow10 = 1;
  int k = 1;
  while (k < 10 && p1 >= pow10 * 10ull) { pow10 *= 10; ++k; }
  int n = k;
  while (n > 0) {
    const uint32_t d = p1 / pow10;
    const uint32_t r = p1 % pow10;
    buffer[length++] = static_cast<char>('0' + d);
    p1 = r;
    --n;
    const uint64_t rest = (uint64_t{p1} << -one.e) + p2;
    if (rest <= delta) {
      decimal_exponent += n;
      const uint64_t ten_n = uint64_t{pow10} << -one.e;
      round_weed(buffer, length, dist, delta, rest, ten_n);
      return;
    }
    pow10 /= 10;
  }
  int m = 0;
  for (;;) {
    p2 *= 10;
    const uint64_t d = p2 >> -one.e;
    const uint64_t r = p2 & (one.f - 1);
    buffer[length++] = static_cast<char>('0' + d);
    p2 = r;
    ++m;
    delta *= 10;
    dist *= 10;
    if (p2 <= delta) break;
  }
  decimal_exponent -= m;
  round_weed(buffer, length, dist, delta, p2, one.f);
}

// digits and decimal exponent of a finite, positive double
inline void shortest(char* buf, int& len, int& decimal_exponent, double value) {
  uint64_t bits;
  memcpy(&bits, &value, 8);
  const uint64_t E = bits >> 52, F = bits & ((uint64_t{1} << 52) - 1);
  const bool denormal = E == 0;
  const Fp v = denormal ? Fp{F, 1 - 1075} : Fp{F + (uint64_t{1} << 52), static_cast<int>(E) - 1075};
  const bool lower_closer = F == 0 && E > 1;
  const Fp m_plus{2 * v.f + 1, v.e - 1};
  const Fp m_minus = lower_closer ? Fp{4 * v.f - 1, v.e - 2} : Fp{2 * v.f - 1, v.e - 1};
  const Fp w_plus = normalize(m_plus);
  const Fp w_minus = align_to(m_minus, w_plus.e);
  const Fp w = normalize(v);
  const Pow10 c = cached_power(w_plus.e);
  const Fp ck{c.f, c.e};
  const Fp W = mul(w, ck), Wm = mul(w_minus, ck), Wp = mul(w_plus, ck);
  const Fp M_minus{Wm.f + 1, Wm.e}, M_plus{Wp.f - 1, Wp.e};
  decimal_exponent = -c.k;
  len = 0;
  digit_gen(buf, len, decimal_exponent, M_minus, W, M_plus);
}

}  // namespace grisu

// nlohmann::json(double).dump()
inline std::string json_double(double value) {
  if (!std::isfinite(value)) return "null";
  std::string out;
  if (std::signbit(value)) { value = -value; out += '-'; }
  if (value == 0) return out + "0.0";
  char buf[32];
  int len = 0, dexp = 0;
  grisu::shortest(buf, len, dexp, value);
  const int min_exp = -4, max_exp = 15;   // std::numeric_limits<double>::digits10
  const int k = len, n = len + dexp;
  if (k <= n && n <= max_exp) {            // digits[000].0
    out.append(buf, k);
    out.append((size_t)(n - k), '0');
    out += ".0";
    return out;
  }
  if (0 < n && n <= max_exp) {             // dig.its
    out.append(buf, n);
    out += '.';
    out.append(buf + n, k - n);
    return out;
  }
  if (min_exp < n && n <= 0) {             // 0.[000]digits
    out += "0.";
    out.append((size_t)(-n), '0');
    out.append(buf, k);
    return out;
  }
  out += buf[0];                           // d[.igits]e+XX
  if (k > 1) { out += '.'; out.append(buf + 1, k - 1); }
  out += 'e';
  int e = n - 1;
  out += e < 0 ? '-' : '+';
  if (e < 0) e = -e;
  if (e < 10) { out += '0'; out += (char)('0' + e); }
  else if (e < 100) { out += (char)('0' + e / 10); out += (char)('0' + e % 10); }
  else { out += (char)('0' + e / 100); e %= 100; out += (char)('0' + e / 10); out += (char)('0' + e % 10); }
  return out;
}

}  // namespace dhost
