// dg_thal.cuh -- the melting-temperature gate of `dicey search` (reference src/silica.h:508-519):
// primer3's nearest-neighbour thermodynamic alignment of a primer against a genomic site, duplex
// mode thal_end1, temperature only (reference src/thal.h: thal() :2408-2655 with initMatrix :821-835,
// LSH :854-960, RSH :963-1076, Ss/Hs :1078-1122, maxTM :1123-1161, calc_bulge_internal :1200-1334,
// fillMatrix :1504-1551, traceback :2134-2179, the Tm line of drawDimer :2196-2204).
//
// SURVEY.md 8(f) rank 1.  Restated as one inline function over plain arrays so that the same source
// runs per thread on the GPU (dg_thal.cu) and on the host (tests/hostsim); results are compared
// bit for bit with the reference's own thal() through oracle/_ref/dicey_ref.  Exactness rests on
// evaluating every sum, product and quotient in the reference's order in IEEE double without
// fused multiply-add (the device translation unit is built with -fmad=false); the only
// transcendental terms (salt correction, R ln(c)) depend on the run's conditions alone and are
// computed once on the host.
#pragma once
#include <stdint.h>

#include "dg_core.cuh"

namespace dg {

constexpr double kThalInf = 999999.0;            // _INFINITY
constexpr double kThalMinEntropyCutoff = -2500.0;
constexpr double kThalMinEntropy = -3224.0;
constexpr double kThalTempK = 310.15;            // TEMP_KELVIN
constexpr double kThalAbsZero = 273.15;
constexpr double kThalInitH = 200.0, kThalInitS = -5.7;   // duplex initiation
constexpr double kThalIlas = (-300 / 310.15);    // internal-loop asymmetry (entropy); the enthalpy term is 0
constexpr int kThalMaxLen = 60;                  // THAL_MAX_ALIGN: the shorter sequence (thal.h:64); the shared-memory forms take both sides <= this
constexpr int kThalMaxSeq = 10000;               // THAL_MAX_SEQ: the longer sequence (thal.h:75); thal_end1_tm_wide takes one side up to this
constexpr int kThalMaxLoop = 30;

// Nearest-neighbour tables exactly as the reference holds them after get_thermodynamic_values()
// (5 x 5 x ... over A C G T N), plus the two condition-dependent constants.
struct ThalParams {
  double stackS[625], stackH[625];
  double int2S[625], int2H[625];
  double dangS3[125], dangH3[125], dangS5[125], dangH5[125];
  double intlS[30], bulgeS[30], intlH[30], bulgeH[30];
  double tstackS[625], tstackH[625];
  double tstack2S[625], tstack2H[625];
  double atpS[25], atpH[25];
  double salt;        // saltCorrectS(mv, dv, dntp)
  double rc[2];       // R ln(c / 1e9) for two self-complementary oligos, R ln(c / 4e9) otherwise
};

DG_HD bool thal_fin(double x) { return x < kThalInf / 2; }
DG_HD int thal_i4(int a, int b, int c, int d) { return ((a * 5 + b) * 5 + c) * 5 + d; }
DG_HD int thal_i3(int a, int b, int c) { return (a * 5 + b) * 5 + c; }
DG_HD int thal_bp(int a, int b) { return (a < 4 && b < 4 && a + b == 3) ? 1 : 0; }   // A-T, C-G
DG_HD int thal_code(uint8_t c) {
  if (c >= 'a' && c <= 'z') c -= 32;
  return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}

struct ThalWork {
  const ThalParams* p;
  const uint8_t* n1;   // 0 .. len1 + 1, sentinels 4
  const uint8_t* n2;   // the second sequence reversed, same framing
  int len1, len2;
  double* ds;          // entropy DP table, len1 x len2
  double* dh;          // enthalpy DP table
  double rc;
  long stride;         // distance between consecutive table cells (1 on the host; on the device the
                       // tables of the threads of a launch are interleaved so that a warp working on
                       // the same cell touches consecutive addresses)
  DG_HD double& S(int i, int j) const { return ds[(long)((j) + (i - 1) * len2 - 1) * stride]; }
  DG_HD double& H(int i, int j) const { return dh[(long)((j) + (i - 1) * len2 - 1) * stride]; }
};

// The terminal stack / dangling-end term on one side of the pair (x .. z): LSH looks left
// (x = n2[j], y = n2[j-1], z = n1[i], w = n1[i-1]), RSH right (x = n1[i], y = n1[i+1], z = n2[j],
// w = n2[j+1]); both pick, among terminal mismatch and dangling ends, the variant with the highest
// melting temperature.  as / ah: AT penalty of the pair.  Returns through (os, oh).
DG_HD void thal_end_term(const ThalWork& w, int x, int y, int z, int ww, double as, double ah, bool open, double& os, double& oh) {
  const ThalParams& p = *w.p;
  double S1, H1, T1, G1, S2, H2, T2, G2;
  T1 = -kThalInf;
  S1 = as + p.tstack2S[thal_i4(x, y, z, ww)];
  H1 = ah + p.tstack2H[thal_i4(x, y, z, ww)];
  G1 = H1 - kThalTempK * S1;
  if (!thal_fin(H1) || G1 > 0) { H1 = kThalInf; S1 = -1.0; G1 = 1.0; }
  const double d3h = p.dangH3[thal_i3(x, y, z)], d5h = p.dangH5[thal_i3(x, z, ww)];
  int variant = 0;
  if (open && thal_fin(d3h) && thal_fin(d5h)) variant = 3;
  else if (open && thal_fin(d3h)) variant = 1;
  else if (open && thal_fin(d5h)) variant = 2;
  if (variant) {
    if (variant == 3) {
      S2 = as + p.dangS3[thal_i3(x, y, z)] + p.dangS5[thal_i3(x, z, ww)];
      H2 = ah + d3h + d5h;
    } else if (variant == 1) {
      S2 = as + p.dangS3[thal_i3(x, y, z)];
      H2 = ah + d3h;
    } else {
      S2 = as + p.dangS5[thal_i3(x, z, ww)];
      H2 = ah + d5h;
    }
    G2 = H2 - kThalTempK * S2;
    if (!thal_fin(H2) || G2 > 0) { H2 = kThalInf; S2 = -1.0; G2 = 1.0; }
    T2 = (H2 + kThalInitH) / (S2 + kThalInitS + w.rc);
    if (thal_fin(H1) && G1 < 0) {
      T1 = (H1 + kThalInitH) / (S1 + kThalInitS + w.rc);
      if (T1 < T2 && G2 < 0) { S1 = S2; H1 = H2; T1 = T2; }
    } else if (G2 < 0) {
      S1 = S2; H1 = H2; T1 = T2;
    }
  }
  S2 = as;
  H2 = ah;
  T2 = (H2 + kThalInitH) / (S2 + kThalInitS + w.rc);
  if (thal_fin(H1) && !(T1 < T2)) { os = S1; oh = H1; }
  else { os = S2; oh = H2; }
}

// RSH(i, j): (-1, inf) when the bases do not pair.
DG_HD void thal_right(const ThalWork& w, int i, int j, double& os, double& oh) {
  const int a = w.n1[i], b = w.n2[j];
  if (!thal_bp(a, b)) { os = -1.0; oh = kThalInf; return; }
  const int a2 = w.n1[i + 1], b2 = w.n2[j + 1];
  thal_end_term(w, a, a2, b, b2, w.p->atpS[a * 5 + b], w.p->atpH[a * 5 + b], thal_bp(a2, b2) == 0, os, oh);
}
// LSH(i, j): a non-pair resets the cell and leaves (os, oh) as the caller set them.
DG_HD void thal_left(const ThalWork& w, int i, int j, double& os, double& oh) {
  const int a = w.n1[i], b = w.n2[j];
  if (!thal_bp(a, b)) { w.S(i, j) = -1.0; w.H(i, j) = kThalInf; return; }
  const int a2 = w.n1[i - 1], b2 = w.n2[j - 1];
  thal_end_term(w, b, b2, a, a2, w.p->atpS[a * 5 + b], w.p->atpH[a * 5 + b], thal_bp(a2, b2) != 1, os, oh);
}

// calc_bulge_internal(i, j, ii, jj), first half: entropy / enthalpy of the structure that reaches
// (ii, jj) through the loop closed by (i, j) on the left and (ii, jj) on the right.
DG_HD void thal_loop_value(const ThalWork& w, int i, int j, int ii, int jj, double& S, double& H) {
  const ThalParams& p = *w.p;
  const uint8_t *n1 = w.n1, *n2 = w.n2;
  const int l1 = ii - i - 1, l2 = jj - j - 1, ls = l1 + l2 - 1;
  S = -1.0;
  H = kThalInf;
  if ((l1 == 0 && l2 > 0) || (l2 == 0 && l1 > 0)) {
    if (l2 == 1 || l1 == 1) {      // a bulge of one base keeps the stack across it
      H = p.bulgeH[ls] + p.stackH[thal_i4(n1[i], n1[ii], n2[j], n2[jj])];
      S = p.bulgeS[ls] + p.stackS[thal_i4(n1[i], n1[ii], n2[j], n2[jj])];
      if (H > 0 || S > 0) { H = kThalInf; S = -1.0; }
      H += w.H(i, j);
      S += w.S(i, j);
      if (!thal_fin(H)) { H = kThalInf; S = -1.0; }
    } else {
      H = p.bulgeH[ls] + p.atpH[n1[i] * 5 + n2[j]] + p.atpH[n1[ii] * 5 + n2[jj]];
      H += w.H(i, j);
      S = p.bulgeS[ls] + p.atpS[n1[i] * 5 + n2[j]] + p.atpS[n1[ii] * 5 + n2[jj]];
      S += w.S(i, j);
      if (!thal_fin(H)) { H = kThalInf; S = -1.0; }
      if (H > 0 && S > 0) { H = kThalInf; S = -1.0; }
    }
  } else if (l1 == 1 && l2 == 1) {
    S = p.int2S[thal_i4(n1[i], n1[i + 1], n2[j], n2[j + 1])] + p.int2S[thal_i4(n2[jj], n2[jj - 1], n1[ii], n1[ii - 1])];
    S += w.S(i, j);
    H = p.int2H[thal_i4(n1[i], n1[i + 1], n2[j], n2[j + 1])] + p.int2H[thal_i4(n2[jj], n2[jj - 1], n1[ii], n1[ii - 1])];
    H += w.H(i, j);
    if (!thal_fin(H)) { H = kThalInf; S = -1.0; }
    if (H > 0 && S > 0) { H = kThalInf; S = -1.0; }
  } else {
    const int asym = l1 > l2 ? l1 - l2 : l2 - l1;
    H = p.intlH[ls] + p.tstackH[thal_i4(n1[i], n1[i + 1], n2[j], n2[j + 1])] +
        p.tstackH[thal_i4(n2[jj], n2[jj - 1], n1[ii], n1[ii - 1])] + (0.0 * asym);
    H += w.H(i, j);
    S = p.intlS[ls] + p.tstackS[thal_i4(n1[i], n1[i + 1], n2[j], n2[j + 1])] +
        p.tstackS[thal_i4(n2[jj], n2[jj - 1], n1[ii], n1[ii - 1])] + (kThalIlas * asym);
    S += w.S(i, j);
    if (!thal_fin(H)) { H = kThalInf; S = -1.0; }
    if (H > 0 && S > 0) { H = kThalInf; S = -1.0; }
  }
}

// The free energy calc_bulge_internal compares structures by; (rs, rh) = RSH of the right end.
DG_HD double thal_loop_energy(double S, double H, double rs, double rh) { return H + rh - kThalTempK * (S + rs); }

// calc_bulge_internal as the reference calls it: the candidate replaces (os, oh) when it is more
// stable than the cell's current value (always, during traceback).
DG_HD void thal_loop(const ThalWork& w, int i, int j, int ii, int jj, bool traceback, double& os, double& oh) {
  double S, H, rs, rh;
  thal_loop_value(w, i, j, ii, jj, S, H);
  thal_right(w, ii, jj, rs, rh);
  const double G1 = thal_loop_energy(S, H, rs, rh);
  const double G2 = thal_loop_energy(w.S(ii, jj), w.H(ii, jj), rs, rh);
  if (G1 < G2 || traceback) { os = S; oh = H; }
}

// maxTM(i, j): keep the cell, or extend the stack from (i-1, j-1), whichever melts higher.
// (rs, rh) = RSH(i, j).  Returns the cell's new value.
DG_HD void thal_stack_value(const ThalWork& w, int i, int j, double rs, double rh, double& oS, double& oH) {
  const ThalParams& p = *w.p;
  double S0 = w.S(i, j), H0 = w.H(i, j), S1, H1, T1;
  oS = S0;
  oH = H0;
  const double T0 = (H0 + kThalInitH + rh) / (S0 + kThalInitS + rs + w.rc);
  const int k = thal_i4(w.n1[i - 1], w.n1[i], w.n2[j - 1], w.n2[j]);
  if (thal_fin(w.H(i - 1, j - 1)) && thal_fin(p.stackH[k])) {
    S1 = (w.S(i - 1, j - 1) + p.stackS[k]);
    H1 = (w.H(i - 1, j - 1) + p.stackH[k]);
    T1 = (H1 + kThalInitH + rh) / (S1 + kThalInitS + rs + w.rc);
  } else {
    S1 = -1.0;
    H1 = kThalInf;
    T1 = (H1 + kThalInitH) / (S1 + kThalInitS + w.rc);
  }
  if (S1 < kThalMinEntropyCutoff) { S1 = kThalMinEntropy; H1 = 0.0; }
  if (S0 < kThalMinEntropyCutoff) { S0 = kThalMinEntropy; H0 = 0.0; }
  if (T1 > T0) { oS = S1; oH = H1; }
  else if (T0 >= T1) { oS = S0; oH = H0; }
}
DG_HD void thal_stack_rs(const ThalWork& w, int i, int j, double rs, double rh) {
  double S, H;
  thal_stack_value(w, i, j, rs, rh, S, H);
  w.S(i, j) = S;
  w.H(i, j) = H;
}
DG_HD void thal_stack(const ThalWork& w, int i, int j) {
  double rs, rh;
  thal_right(w, i, j, rs, rh);
  thal_stack_rs(w, i, j, rs, rh);
}

DG_HD bool thal_symmetric(const uint8_t* s, int len) {   // symmetry_thermo
  if (len % 2 == 1) return false;
  for (int i = 0; i < len / 2; ++i) {
    const int a = thal_code(s[i]), b = thal_code(s[len - 1 - i]);
    if ((a < 4 || b < 4) && a + b != 3) return false;   // any A/C/G/T on either side must face its complement
  }
  return true;
}

// thal(oligo1, oligo2, thal_end1, temponly): returns false where the reference returns false
// (a sequence longer than the DP tables); *tm is o->temp.  num1 / num2: len + 2 bytes each,
// ds / dh: len1 * len2 cells each, `stride` doubles apart.
// thal_end1_tm_any: the same for any pair of lengths the reference accepts (thal.h:2440-2451: at most
// one side longer than THAL_MAX_ALIGN, neither longer than THAL_MAX_SEQ); the caller sizes the work areas.
DG_HD bool thal_lengths_ok(int len1, int len2) {
  return len1 > 0 && len2 > 0 && !(len1 > kThalMaxLen && len2 > kThalMaxLen) && len1 <= kThalMaxSeq && len2 <= kThalMaxSeq;
}
DG_HD bool thal_end1_tm_any(const ThalParams* p, const uint8_t* o1, int len1, const uint8_t* o2, int len2, uint8_t* num1,
                            uint8_t* num2, double* ds, double* dh, double* tm, long stride = 1);
DG_HD bool thal_end1_tm(const ThalParams* p, const uint8_t* o1, int len1, const uint8_t* o2, int len2, uint8_t* num1,
                        uint8_t* num2, double* ds, double* dh, double* tm, long stride = 1) {
  *tm = -kThalInf;   // THAL_ERROR_SCORE
  if (len1 <= 0 || len2 <= 0) { *tm = 0.0; return false; }
  if (len1 > kThalMaxLen || len2 > kThalMaxLen) return false;   // (the work areas of this form's callers)
  return thal_end1_tm_any(p, o1, len1, o2, len2, num1, num2, ds, dh, tm, stride);
}
DG_HD bool thal_end1_tm_any(const ThalParams* p, const uint8_t* o1, int len1, const uint8_t* o2, int len2, uint8_t* num1,
                            uint8_t* num2, double* ds, double* dh, double* tm, long stride) {
  *tm = -kThalInf;   // THAL_ERROR_SCORE
  if (len1 <= 0 || len2 <= 0) { *tm = 0.0; return false; }
  if (!thal_lengths_ok(len1, len2)) return false;
  ThalWork w;
  w.p = p; w.n1 = num1; w.n2 = num2; w.len1 = len1; w.len2 = len2; w.ds = ds; w.dh = dh; w.stride = stride;
  w.rc = (thal_symmetric(o1, len1) && thal_symmetric(o2, len2)) ? p->rc[0] : p->rc[1];
  for (int i = 1; i <= len1; ++i) num1[i] = (uint8_t)thal_code(o1[i - 1]);
  for (int j = 1; j <= len2; ++j) num2[j] = (uint8_t)thal_code(o2[len2 - j]);   // 3' -> 5'
  num1[0] = num1[len1 + 1] = num2[0] = num2[len2 + 1] = 4;
  // initMatrix
  for (int i = 1; i <= len1; ++i)
    for (int j = 1; j <= len2; ++j) {
      const bool pair = thal_bp(num1[i], num2[j]) != 0;
      w.H(i, j) = pair ? 0.0 : kThalInf;
      w.S(i, j) = pair ? kThalMinEntropy : -1.0;
    }
  // fillMatrix
  for (int i = 1; i <= len1; ++i)
    for (int j = 1; j <= len2; ++j) {
      if (!thal_fin(w.H(i, j))) continue;
      double s = -1.0, h = kThalInf;
      thal_left(w, i, j, s, h);
      if (thal_fin(h)) { w.S(i, j) = s; w.H(i, j) = h; }
      if (i > 1 && j > 1) {
        thal_stack(w, i, j);
        for (int d = 3; d <= kThalMaxLoop + 2; ++d) {
          int ii = i - 1;
          int jj = -ii - d + (j + i);
          if (jj < 1) { ii -= (1 - jj); jj = 1; }
          for (; ii > 0 && jj < j; --ii, ++jj) {
            if (!thal_fin(w.H(ii, jj))) continue;
            s = -1.0; h = kThalInf;
            thal_loop(w, ii, jj, i, j, false, s, h);
            if (s < kThalMinEntropyCutoff) { s = kThalMinEntropy; h = 0.0; }
            if (thal_fin(h)) { w.H(i, j) = h; w.S(i, j) = s; }
          }
        }
      }
    }
  // the most stable structure ending at the 3' end of the first sequence (thal_end1)
  int bestI = len1, bestJ = 0;
  double bestG = kThalInf;
  for (int j = 1; j <= len2; ++j) {
    double s, h;
    thal_right(w, len1, j, s, h);
    s = s + 0.000001;
    h = h + 0.000001;
    const double G1 = (w.H(len1, j) + h + kThalInitH) - kThalTempK * (w.S(len1, j) + s + kThalInitS);
    if (G1 < bestG) { bestG = G1; bestJ = j; }
  }
  if (!thal_fin(bestG)) bestI = bestJ = 1;
  double rs, rh;
  thal_right(w, bestI, bestJ, rs, rh);
  const double dH = w.H(bestI, bestJ) + rh + kThalInitH;
  const double dS = (w.S(bestI, bestJ) + rs + kThalInitS);
  if (!thal_fin(w.H(bestI, bestJ))) { *tm = 0.0; return true; }
  // traceback: only the number of paired bases enters the temperature
  int paired = 2, i = bestI, j = bestJ;   // ps1[i-1] and ps2[j-1]
  for (int guard = 0; guard < 4 * (len1 + len2); ++guard) {
    double s = -1.0, h = kThalInf;
    thal_left(w, i, j, s, h);
    if (w.S(i, j) == s && w.H(i, j) == h) break;
    bool done = false;
    if (i > 1 && j > 1) {
      const int k = thal_i4(num1[i - 1], num1[i], num2[j - 1], num2[j]);
      if (w.S(i, j) == p->stackS[k] + w.S(i - 1, j - 1) && w.H(i, j) == p->stackH[k] + w.H(i - 1, j - 1)) {
        --i; --j;
        paired += 2;
        done = true;
      }
    }
    for (int d = 3; !done && d <= kThalMaxLoop + 2; ++d) {
      int ii = i - 1;
      int jj = -ii - d + (j + i);
      if (jj < 1) { ii -= (1 - jj); jj = 1; }
      for (; !done && ii > 0 && jj < j; --ii, ++jj) {
        s = -1.0; h = kThalInf;
        thal_loop(w, ii, jj, i, j, true, s, h);
        if (w.S(i, j) == s && w.H(i, j) == h) {
          i = ii; j = jj;
          paired += 2;
          done = true;
          break;
        }
      }
    }
    if (!done) break;   // (the reference would spin here; never observed)
  }
  const int N = (paired / 2) - 1;
  *tm = ((dH) / (dS + (N * p->salt) + w.rc)) - kThalAbsZero;
  return true;
}


// thal_loop_value with everything that depends on the right-hand cell (ii, jj) alone looked up
// once (the lanes form evaluates many left-hand partners (i, j) against one right-hand cell), and
// the 4-base table indices composed from per-position base-pair codes:
//   a1[i] = n1[i] * 5 + n1[i + 1]   ra[i] = n1[i] * 5 + n1[i - 1]
//   b[j]  = n2[j] * 5 + n2[j + 1]   rb[j] = n2[j] * 5 + n2[j - 1]
// Same additions in the same order as thal_loop_value.
struct ThalRight {
  double tsS, tsH, i2S, i2H, atS, atH;   // tstack / int2 at (n2[jj], n2[jj-1], n1[ii], n1[ii-1]); AT penalty of (ii, jj)
  int mix;                               // n1[ii] * 25 + n2[jj]
};
DG_HD void thal_right_consts(const ThalWork& w, const uint8_t* ra, const uint8_t* rb, int ii, int jj, ThalRight& c) {
  const ThalParams& p = *w.p;
  const int k = rb[jj] * 25 + ra[ii];
  c.tsS = p.tstackS[k]; c.tsH = p.tstackH[k];
  c.i2S = p.int2S[k]; c.i2H = p.int2H[k];
  c.atS = p.atpS[w.n1[ii] * 5 + w.n2[jj]]; c.atH = p.atpH[w.n1[ii] * 5 + w.n2[jj]];
  c.mix = w.n1[ii] * 25 + w.n2[jj];
}
DG_HD void thal_loop_value_right(const ThalWork& w, const uint8_t* a1, const uint8_t* b, const ThalRight& c, int i, int j, int ii, int jj,
                                 double& S, double& H) {
  const ThalParams& p = *w.p;
  const int l1 = ii - i - 1, l2 = jj - j - 1, ls = l1 + l2 - 1;
  const double cS = w.S(i, j), cH = w.H(i, j);
  if (l1 == 0 || l2 == 0) {        // (not both: the caller's d >= 3)
    if (l2 == 1 || l1 == 1) {
      const int k = w.n1[i] * 125 + w.n2[j] * 5 + c.mix;
      H = p.bulgeH[ls] + p.stackH[k];
      S = p.bulgeS[ls] + p.stackS[k];
      if (H > 0 || S > 0) { H = kThalInf; S = -1.0; }
      H += cH;
      S += cS;
      if (!thal_fin(H)) { H = kThalInf; S = -1.0; }
    } else {
      const int k = w.n1[i] * 5 + w.n2[j];
      H = p.bulgeH[ls] + p.atpH[k] + c.atH;
      H += cH;
      S = p.bulgeS[ls] + p.atpS[k] + c.atS;
      S += cS;
      if (!thal_fin(H)) { H = kThalInf; S = -1.0; }
      if (H > 0 && S > 0) { H = kThalInf; S = -1.0; }
    }
  } else if (l1 == 1 && l2 == 1) {
    const int k = a1[i] * 25 + b[j];
    S = p.int2S[k] + c.i2S;
    S += cS;
    H = p.int2H[k] + c.i2H;
    H += cH;
    if (!thal_fin(H)) { H = kThalInf; S = -1.0; }
    if (H > 0 && S > 0) { H = kThalInf; S = -1.0; }
  } else {
    const int k = a1[i] * 25 + b[j];
    const int asym = l1 > l2 ? l1 - l2 : l2 - l1;
    H = p.intlH[ls] + p.tstackH[k] + c.tsH + (0.0 * asym);
    H += cH;
    S = p.intlS[ls] + p.tstackS[k] + c.tsS + (kThalIlas * asym);
    S += cS;
    if (!thal_fin(H)) { H = kThalInf; S = -1.0; }
    if (H > 0 && S > 0) { H = kThalInf; S = -1.0; }
  }
}

// ------------------------------------------------------------------------------------------
// The same computation arranged for a group of cooperating lanes (a warp on the device, one lane
// in the host build that the CPU tests run):
//   * only cells whose bases pair are ever finite, so the fill walks a row-major list of paired
//     cells instead of the whole table;
//   * LSH (thal_left) and RSH (thal_right) depend on the sequences alone: LSH seeds the table for
//     all paired cells at once, RSH is evaluated once per cell;
//   * a cell reads only cells above and to the left of it, so the cells of one row are independent:
//     the lanes split into one group per paired cell of the row and the groups work side by side;
//   * for one cell the reference scans its loop partners (ii, jj) in a fixed order and keeps a
//     candidate when its free energy is strictly below that of the cell's current value, which
//     then becomes the candidate -- a running strict minimum.  Its outcome is the first partner
//     in scan order that attains the overall minimum, if that is below the starting value; the
//     lanes of a group evaluate partners independently and an arg-min over (energy, scan rank)
//     picks it.  One thing would break that equivalence and makes the function return 2 ("use
//     the sequential form"): a candidate entropy below the -2500 cutoff (the reference then
//     stores a substitute value), which no pair of sequences within the length limit produces;
//   * the traceback looks for the first partner in scan order whose value reproduces the cell:
//     the same arg-min with a constant energy.
// Warp concept: static n (lanes), lane, sync(), ballot(p), lanemask_lt(), all(p), any(p),
// group_mask(first_lane, lanes, member), argmin(mask, g, order, S, H) -- after it every lane of
// the group `mask` holds the group's (g, order) minimum (ties: lower order) and that lane's S, H.
struct ThalOneLane {
  static constexpr int n = 1;
  int lane = 0;
  DG_HD void sync() const {}
  DG_HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
  DG_HD unsigned lanemask_lt() const { return 0u; }
  DG_HD bool all(bool p) const { return p; }
  DG_HD bool any(bool p) const { return p; }
  DG_HD unsigned group_mask(int, int, bool) const { return 1u; }
  DG_HD void argmin(unsigned, double&, uint32_t&, double&, double&) const {}
};

constexpr uint32_t kThalNone = 0xFFFFFFFFu;
DG_HD int thal_popc(unsigned x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// Work areas: num1 / num2 len + 2 bytes; tab 2 * len1 * len2 doubles (S, H interleaved);
// plist len1 * len2 entries; rstart len1 + 2 entries; codes 4 * 64 bytes.
// Returns 0 where the reference's thal() fails, 1 with *tm set, 2 for "use thal_end1_tm".
template <class Warp>
DG_HD int thal_end1_tm_lanes(Warp& wp, const ThalParams* p, const uint8_t* o1, int len1, const uint8_t* o2, int len2, uint8_t* num1,
                             uint8_t* num2, double* tab, uint16_t* plist, uint16_t* rstart, uint8_t* codes, double* tm) {
  constexpr int n = Warp::n;
  const int lane = wp.lane;
  const unsigned everyone = wp.group_mask(0, n, true);
  *tm = -kThalInf;
  if (len1 <= 0 || len2 <= 0) { *tm = 0.0; return 0; }
  if (len1 > kThalMaxLen || len2 > kThalMaxLen) return 0;
  for (int i = 1 + lane; i <= len1; i += n) num1[i] = (uint8_t)thal_code(o1[i - 1]);
  for (int j = 1 + lane; j <= len2; j += n) num2[j] = (uint8_t)thal_code(o2[len2 - j]);
  if (lane == 0) num1[0] = num1[len1 + 1] = num2[0] = num2[len2 + 1] = 4;
  wp.sync();
  uint8_t *a1 = codes, *ra = codes + 64, *b = codes + 128, *rb = codes + 192;
  for (int i = 1 + lane; i <= len1; i += n) {
    a1[i] = (uint8_t)(num1[i] * 5 + num1[i + 1]);
    ra[i] = (uint8_t)(num1[i] * 5 + num1[i - 1]);
  }
  for (int j = 1 + lane; j <= len2; j += n) {
    b[j] = (uint8_t)(num2[j] * 5 + num2[j + 1]);
    rb[j] = (uint8_t)(num2[j] * 5 + num2[j - 1]);
  }
  bool sym = (len1 % 2 == 0) && (len2 % 2 == 0);
  if (sym) {
    bool mine = true;
    for (int t = lane; t < len1 / 2; t += n) {
      const int a = num1[1 + t], b = num1[len1 - t];
      if ((a < 4 || b < 4) && a + b != 3) mine = false;
    }
    for (int t = lane; t < len2 / 2; t += n) {
      const int a = num2[1 + t], b = num2[len2 - t];
      if ((a < 4 || b < 4) && a + b != 3) mine = false;
    }
    sym = wp.all(mine);
  }
  ThalWork w;
  w.p = p; w.n1 = num1; w.n2 = num2; w.len1 = len1; w.len2 = len2; w.ds = tab; w.dh = tab + 1; w.stride = 2;
  w.rc = sym ? p->rc[0] : p->rc[1];
  // initMatrix + the list of paired cells, row-major
  int np = 0;
  for (int i = 1; i <= len1; ++i) {
    if (lane == 0) rstart[i] = (uint16_t)np;
    for (int jb = 1; jb <= len2; jb += n) {
      const int j = jb + lane;
      const bool in = j <= len2;
      const bool pr = in && thal_bp(num1[i], num2[j]) != 0;
      if (in) { w.H(i, j) = pr ? 0.0 : kThalInf; w.S(i, j) = pr ? kThalMinEntropy : -1.0; }
      const unsigned m = wp.ballot(pr);
      if (pr) plist[np + thal_popc(m & wp.lanemask_lt())] = (uint16_t)((i << 8) | j);
      np += thal_popc(m);
    }
  }
  if (lane == 0) rstart[len1 + 1] = (uint16_t)np;
  wp.sync();
  for (int e = lane; e < np; e += n) {   // LSH of every paired cell
    const int i = plist[e] >> 8, j = plist[e] & 0xff;
    double s = -1.0, h = kThalInf;
    thal_left(w, i, j, s, h);
    if (thal_fin(h)) { w.S(i, j) = s; w.H(i, j) = h; }
  }
  wp.sync();
  // fillMatrix, one row of paired cells at a time
  bool odd = false;
  for (int i = 2; i <= len1; ++i) {
    const int r0 = rstart[i], r1 = rstart[i + 1];
    const int lo = rstart[i > kThalMaxLoop + 1 ? i - (kThalMaxLoop + 1) : 1], hi = r0;
    for (int c0 = r0; c0 < r1; c0 += n) {
      const int cells = r1 - c0 < n ? r1 - c0 : n;
      const int gsize = n / cells, g = lane / gsize, gl = lane - g * gsize;
      const bool member = g < cells;
      const unsigned gmask = wp.group_mask(g * gsize, gsize, member);
      const int j = member ? (plist[c0 + g] & 0xff) : 0;
      const bool live = member && j > 1 && thal_fin(w.H(i, j));
      double cs = -1.0, ch = kThalInf, rs = -1.0, rh = kThalInf, G2 = 0.0;
      double bestG = 1e300, bS = -1.0, bH = kThalInf;
      uint32_t bestO = kThalNone;
      if (live) {
        thal_right(w, i, j, rs, rh);
        thal_stack_value(w, i, j, rs, rh, cs, ch);
        G2 = thal_loop_energy(cs, ch, rs, rh);
        ThalRight rc;
        thal_right_consts(w, ra, rb, i, j, rc);
        for (int e = lo + gl; e < hi; e += gsize) {
          const int ii = plist[e] >> 8, jj = plist[e] & 0xff;
          const int d = (i - ii) + (j - jj);
          if (jj >= j || d < 3 || d > kThalMaxLoop + 2) continue;
          if (!thal_fin(w.H(ii, jj))) continue;
          double S, H;
          thal_loop_value_right(w, a1, b, rc, ii, jj, i, j, S, H);
          if (!thal_fin(H)) continue;                    // never stored by the reference
          if (S < kThalMinEntropyCutoff) odd = true;
          const double G1 = thal_loop_energy(S, H, rs, rh) + 0.0;   // (+ 0.0: one zero for the bitwise arg-min)
          const uint32_t order = ((uint32_t)d << 6) | (uint32_t)(i - 1 - ii);
          if (bestO == kThalNone || G1 < bestG || (G1 == bestG && order < bestO)) { bestG = G1; bestO = order; bS = S; bH = H; }
        }
      }
      wp.sync();
      wp.argmin(gmask, bestG, bestO, bS, bH);
      if (live && gl == 0) {
        const bool take = bestO != kThalNone && bestG < G2;
        w.S(i, j) = take ? bS : cs;
        w.H(i, j) = take ? bH : ch;
      }
      wp.sync();
    }
  }
  if (wp.any(odd)) return 2;
  // the most stable structure ending at the 3' end of the first sequence
  int bestI = len1, bestJ = 0;
  {
    double bestG = 1e300, dS_ = 0.0, dH_ = 0.0;
    uint32_t bestO = kThalNone;
    for (int j = 1 + lane; j <= len2; j += n) {
      double s, h;
      thal_right(w, len1, j, s, h);
      s = s + 0.000001;
      h = h + 0.000001;
      const double G1 = ((w.H(len1, j) + h + kThalInitH) - kThalTempK * (w.S(len1, j) + s + kThalInitS)) + 0.0;
      if (G1 < kThalInf && (bestO == kThalNone || G1 < bestG)) { bestG = G1; bestO = (uint32_t)j; }
    }
    wp.sync();
    wp.argmin(everyone, bestG, bestO, dS_, dH_);
    if (bestO != kThalNone && thal_fin(bestG)) bestJ = (int)bestO;
    else bestI = bestJ = 1;
  }
  double rs, rh;
  thal_right(w, bestI, bestJ, rs, rh);
  const double dH = w.H(bestI, bestJ) + rh + kThalInitH;
  const double dS = (w.S(bestI, bestJ) + rs + kThalInitS);
  if (!thal_fin(w.H(bestI, bestJ))) { *tm = 0.0; return 1; }
  // traceback: only the number of paired bases enters the temperature
  int paired = 2, i = bestI, j = bestJ;
  for (int guard = 0; guard < 4 * (len1 + len2); ++guard) {
    double s = -1.0, h = kThalInf;
    if (thal_bp(num1[i], num2[j])) thal_left(w, i, j, s, h);   // (a finite cell always pairs)
    const double cs = w.S(i, j), ch = w.H(i, j);
    if (cs == s && ch == h) break;
    if (i > 1 && j > 1) {
      const int k = thal_i4(num1[i - 1], num1[i], num2[j - 1], num2[j]);
      if (cs == p->stackS[k] + w.S(i - 1, j - 1) && ch == p->stackH[k] + w.H(i - 1, j - 1)) {
        --i; --j;
        paired += 2;
        continue;
      }
    }
    double g = 0.0, dS_ = 0.0, dH_ = 0.0;
    uint32_t bestO = kThalNone;
    const int lo = rstart[i > kThalMaxLoop + 1 ? i - (kThalMaxLoop + 1) : 1], hi = rstart[i];
    ThalRight rc;
    thal_right_consts(w, ra, rb, i, j, rc);
    for (int e = lo + lane; e < hi; e += n) {
      const int ii = plist[e] >> 8, jj = plist[e] & 0xff;
      const int d = (i - ii) + (j - jj);
      if (jj >= j || d < 3 || d > kThalMaxLoop + 2) continue;
      double S, H;
      thal_loop_value_right(w, a1, b, rc, ii, jj, i, j, S, H);
      const uint32_t order = ((uint32_t)d << 6) | (uint32_t)(i - 1 - ii);
      if (cs == S && ch == H && order < bestO) bestO = order;
    }
    wp.sync();
    wp.argmin(everyone, g, bestO, dS_, dH_);
    if (bestO == kThalNone) break;   // (the reference would spin here; never observed)
    const int d = (int)(bestO >> 6), ii = i - 1 - (int)(bestO & 63);
    j = j - (d - (i - ii));
    i = ii;
    paired += 2;
  }
  const int N = (paired / 2) - 1;
  *tm = ((dH) / (dS + (N * p->salt) + w.rc)) - kThalAbsZero;
  return 1;
}

// ------------------------------------------------------------------------------------------
// thal_end1_tm_wide: the lane-cooperative computation for the pairs the reference accepts with ONE
// side longer than THAL_MAX_ALIGN (thal.h:58, :2440-2451: "one of the sequences must be <= 60, the
// other can be longer", up to THAL_MAX_SEQ = 10 000).  The table no longer fits shared memory, so
// the work areas are the caller's (global memory on the device) and two things change against
// thal_end1_tm_lanes, the arithmetic and every decision staying the same:
//   * no list of paired cells: the live cells of a row are found 32 columns at a time with a ballot,
//     and the lanes split into one group per live cell of that chunk;
//   * the loop partners of a cell are enumerated by coordinates -- (ii, jj) = (i - di, j - dj),
//     3 <= di + dj <= maxLoop + 2 -- and the table says which of them are finite (a cell is finite
//     exactly when its bases pair).  The reference scans them by ascending d = di + dj and, within
//     one d, ascending di: `order` below is that rank, and the arg-min over (energy, order) picks the
//     partner the sequential scan would have kept (argument: see thal_end1_tm_lanes).
// Work areas: num1 / a1 / ra len1 + 2 bytes each, num2 / b / rb len2 + 2 bytes each, tab
// 2 * len1 * len2 doubles (S, H interleaved).  Returns 0 where the reference's thal() fails, 1 with
// *tm set, 2 for "use thal_end1_tm_any" (the entropy cutoff, as in thal_end1_tm_lanes).
DG_HD int thal_nth_bit(unsigned m, int k) {   // position of the k-th (0-based) set bit of m
#ifdef __CUDA_ARCH__
  return (int)__fns(m, 0, k + 1);
#else
  for (int t = 0; t < k; ++t) m &= m - 1;
  return __builtin_ctz(m);
#endif
}

template <class Warp>
DG_HD int thal_end1_tm_wide(Warp& wp, const ThalParams* p, const uint8_t* o1, int len1, const uint8_t* o2, int len2, uint8_t* num1,
                            uint8_t* num2, double* tab, uint8_t* a1, uint8_t* ra, uint8_t* b, uint8_t* rb, double* tm) {
  constexpr int n = Warp::n;
  const int lane = wp.lane;
  const unsigned everyone = wp.group_mask(0, n, true);
  *tm = -kThalInf;
  if (len1 <= 0 || len2 <= 0) { *tm = 0.0; return 0; }
  if (!thal_lengths_ok(len1, len2)) return 0;
  for (int i = 1 + lane; i <= len1; i += n) num1[i] = (uint8_t)thal_code(o1[i - 1]);
  for (int j = 1 + lane; j <= len2; j += n) num2[j] = (uint8_t)thal_code(o2[len2 - j]);
  if (lane == 0) num1[0] = num1[len1 + 1] = num2[0] = num2[len2 + 1] = 4;
  wp.sync();
  for (int i = 1 + lane; i <= len1; i += n) {
    a1[i] = (uint8_t)(num1[i] * 5 + num1[i + 1]);
    ra[i] = (uint8_t)(num1[i] * 5 + num1[i - 1]);
  }
  for (int j = 1 + lane; j <= len2; j += n) {
    b[j] = (uint8_t)(num2[j] * 5 + num2[j + 1]);
    rb[j] = (uint8_t)(num2[j] * 5 + num2[j - 1]);
  }
  bool sym = (len1 % 2 == 0) && (len2 % 2 == 0);
  if (sym) {
    bool mine = true;
    for (int t = lane; t < len1 / 2; t += n) {
      const int x = num1[1 + t], y = num1[len1 - t];
      if ((x < 4 || y < 4) && x + y != 3) mine = false;
    }
    for (int t = lane; t < len2 / 2; t += n) {
      const int x = num2[1 + t], y = num2[len2 - t];
      if ((x < 4 || y < 4) && x + y != 3) mine = false;
    }
    sym = wp.all(mine);
  }
  ThalWork w;
  w.p = p; w.n1 = num1; w.n2 = num2; w.len1 = len1; w.len2 = len2; w.ds = tab; w.dh = tab + 1; w.stride = 2;
  w.rc = sym ? p->rc[0] : p->rc[1];
  // initMatrix and the LSH seed of every paired cell (both depend on the sequences alone)
  for (int i = 1; i <= len1; ++i)
    for (int j = 1 + lane; j <= len2; j += n) {
      const bool pr = thal_bp(num1[i], num2[j]) != 0;
      double s = pr ? kThalMinEntropy : -1.0, h = pr ? 0.0 : kThalInf;
      if (pr) {
        double ls = -1.0, lh = kThalInf;
        thal_left(w, i, j, ls, lh);   // (touches the table only when the bases do not pair)
        if (thal_fin(lh)) { s = ls; h = lh; }
      }
      w.S(i, j) = s;
      w.H(i, j) = h;
    }
  wp.sync();
  // fillMatrix: rows in order, the live cells of a row n columns at a time
  bool odd = false;
  for (int i = 2; i <= len1; ++i) {
    for (int jb = 2; jb <= len2; jb += n) {
      const int jm = jb + lane;
      const unsigned livemask = wp.ballot(jm <= len2 && thal_fin(w.H(i, jm)));
      const int cells = thal_popc(livemask);
      if (cells == 0) continue;
      const int gsize = n / cells, g = lane / gsize, gl = lane - g * gsize;
      const bool live = g < cells;
      const unsigned gmask = wp.group_mask(g * gsize, gsize, live);
      const int j = live ? jb + thal_nth_bit(livemask, g) : 0;
      double cs = -1.0, ch = kThalInf, rs = -1.0, rh = kThalInf, G2 = 0.0;
      double bestG = 1e300, bS = -1.0, bH = kThalInf;
      uint32_t bestO = kThalNone;
      if (live) {
        thal_right(w, i, j, rs, rh);
        thal_stack_value(w, i, j, rs, rh, cs, ch);
        G2 = thal_loop_energy(cs, ch, rs, rh);
        ThalRight rc;
        thal_right_consts(w, ra, rb, i, j, rc);
        for (int di = 1; di <= kThalMaxLoop + 1 && di < i; ++di) {
          const int ii = i - di;
          for (int dj = (di == 1 ? 2 : 1) + gl; dj <= kThalMaxLoop + 2 - di && dj < j; dj += gsize) {
            const int jj = j - dj;
            if (!thal_fin(w.H(ii, jj))) continue;
            double S, H;
            thal_loop_value_right(w, a1, b, rc, ii, jj, i, j, S, H);
            if (!thal_fin(H)) continue;                    // never stored by the reference
            if (S < kThalMinEntropyCutoff) odd = true;
            const double G1 = thal_loop_energy(S, H, rs, rh) + 0.0;   // (+ 0.0: one zero for the bitwise arg-min)
            const uint32_t order = ((uint32_t)(di + dj) << 6) | (uint32_t)(di - 1);
            if (bestO == kThalNone || G1 < bestG || (G1 == bestG && order < bestO)) { bestG = G1; bestO = order; bS = S; bH = H; }
          }
        }
      }
      wp.sync();
      wp.argmin(gmask, bestG, bestO, bS, bH);
      if (live && gl == 0) {
        const bool take = bestO != kThalNone && bestG < G2;
        w.S(i, j) = take ? bS : cs;
        w.H(i, j) = take ? bH : ch;
      }
      wp.sync();
    }
  }
  if (wp.any(odd)) return 2;
  // the most stable structure ending at the 3' end of the first sequence
  int bestI = len1, bestJ = 0;
  {
    double bestG = 1e300, dS_ = 0.0, dH_ = 0.0;
    uint32_t bestO = kThalNone;
    for (int j = 1 + lane; j <= len2; j += n) {
      double s, h;
      thal_right(w, len1, j, s, h);
      s = s + 0.000001;
      h = h + 0.000001;
      const double G1 = ((w.H(len1, j) + h + kThalInitH) - kThalTempK * (w.S(len1, j) + s + kThalInitS)) + 0.0;
      if (G1 < kThalInf && (bestO == kThalNone || G1 < bestG)) { bestG = G1; bestO = (uint32_t)j; }
    }
    wp.sync();
    wp.argmin(everyone, bestG, bestO, dS_, dH_);
    if (bestO != kThalNone && thal_fin(bestG)) bestJ = (int)bestO;
    else bestI = bestJ = 1;
  }
  double rs, rh;
  thal_right(w, bestI, bestJ, rs, rh);
  const double dH = w.H(bestI, bestJ) + rh + kThalInitH;
  const double dS = (w.S(bestI, bestJ) + rs + kThalInitS);
  if (!thal_fin(w.H(bestI, bestJ))) { *tm = 0.0; return 1; }
  // traceback: only the number of paired bases enters the temperature
  int paired = 2, i = bestI, j = bestJ;
  for (int guard = 0; guard < 4 * (len1 + len2); ++guard) {
    double s = -1.0, h = kThalInf;
    if (thal_bp(num1[i], num2[j])) thal_left(w, i, j, s, h);   // (a finite cell always pairs)
    const double cs = w.S(i, j), ch = w.H(i, j);
    if (cs == s && ch == h) break;
    if (i > 1 && j > 1) {
      const int k = thal_i4(num1[i - 1], num1[i], num2[j - 1], num2[j]);
      if (cs == p->stackS[k] + w.S(i - 1, j - 1) && ch == p->stackH[k] + w.H(i - 1, j - 1)) {
        --i; --j;
        paired += 2;
        continue;
      }
    }
    double g = 0.0, dS_ = 0.0, dH_ = 0.0;
    uint32_t bestO = kThalNone;
    ThalRight rc;
    thal_right_consts(w, ra, rb, i, j, rc);
    for (int di = 1; di <= kThalMaxLoop + 1 && di < i; ++di) {
      const int ii = i - di;
      for (int dj = (di == 1 ? 2 : 1) + lane; dj <= kThalMaxLoop + 2 - di && dj < j; dj += n) {
        const int jj = j - dj;
        if (!thal_fin(w.H(ii, jj))) continue;   // (cannot reproduce a finite cell)
        double S, H;
        thal_loop_value_right(w, a1, b, rc, ii, jj, i, j, S, H);
        const uint32_t order = ((uint32_t)(di + dj) << 6) | (uint32_t)(di - 1);
        if (cs == S && ch == H && order < bestO) bestO = order;
      }
    }
    wp.sync();
    wp.argmin(everyone, g, bestO, dS_, dH_);
    if (bestO == kThalNone) break;   // (the reference would spin here; never observed)
    const int d = (int)(bestO >> 6), di = 1 + (int)(bestO & 63);
    j = j - (d - di);
    i = i - di;
    paired += 2;
  }
  const int N = (paired / 2) - 1;
  *tm = ((dH) / (dS + (N * p->salt) + w.rc)) - kThalAbsZero;
  return 1;
}

}  // namespace dg
