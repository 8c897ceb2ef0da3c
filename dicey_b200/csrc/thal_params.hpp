// thal_params.hpp -- host-side loaders of the nearest-neighbour tables of dg_thal.cuh.
//
//   thal_params_from_dump   the table dump `oracle/_ref/dicey_ref thal ... params.tsv` writes (bit
//                           patterns of the arrays the reference holds after
//                           get_thermodynamic_values(), src/thal.h:2368-2393): test fixtures
//   thal_params_from_config the primer3_config directory a dicey installation ships (the `-i`
//                           option of `dicey search`, silica.h:216,300-320): product path
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "dg_thal.cuh"

namespace dg {

inline bool thal_params_from_dump(const char* path, ThalParams& tp, std::string& err) {
  std::ifstream f(path);
  if (!f) { err = std::string("cannot open ") + path; return false; }
  memset(&tp, 0, sizeof(tp));
  struct Slot { const char* name; double* dst; size_t n; };
  Slot slots[] = {{"stackEntropies", tp.stackS, 625}, {"stackEnthalpies", tp.stackH, 625}, {"stackint2Entropies", tp.int2S, 625},
                  {"stackint2Enthalpies", tp.int2H, 625}, {"dangleEntropies3", tp.dangS3, 125}, {"dangleEnthalpies3", tp.dangH3, 125},
                  {"dangleEntropies5", tp.dangS5, 125}, {"dangleEnthalpies5", tp.dangH5, 125}, {"interiorLoopEntropies", tp.intlS, 30},
                  {"bulgeLoopEntropies", tp.bulgeS, 30}, {"interiorLoopEnthalpies", tp.intlH, 30}, {"bulgeLoopEnthalpies", tp.bulgeH, 30},
                  {"tstackEntropies", tp.tstackS, 625}, {"tstackEnthalpies", tp.tstackH, 625}, {"tstack2Entropies", tp.tstack2S, 625},
                  {"tstack2Enthalpies", tp.tstack2H, 625}, {"atpS", tp.atpS, 25}, {"atpH", tp.atpH, 25}, {"saltCorrection", &tp.salt, 1},
                  {"RC_symmetric_asymmetric", tp.rc, 2}};
  size_t seen = 0;
  std::string line;
  while (std::getline(f, line)) {
    std::istringstream is(line);
    std::string name;
    size_t n = 0;
    is >> name >> n;
    for (auto& s : slots)
      if (name == s.name) {
        if (n != s.n) { err = "bad table size for " + name; return false; }
        for (size_t i = 0; i < n; ++i) {
          std::string hex;
          is >> hex;
          uint64_t u = strtoull(hex.c_str(), nullptr, 16);
          memcpy(s.dst + i, &u, 8);
        }
        ++seen;
      }
  }
  if (seen != sizeof(slots) / sizeof(slots[0])) { err = "incomplete table dump"; return false; }
  return true;
}

// One number per line, "inf" = _INFINITY (readDouble, thal.h:403-414).
class ThalFile {
 public:
  ThalFile(const std::string& dir, const char* name) : f_(fopen((dir + name).c_str(), "rt")) {}
  ~ThalFile() { if (f_) fclose(f_); }
  bool ok() const { return f_ != nullptr; }
  bool line(std::string& out) {
    out.clear();
    if (!f_) return false;
    char buf[1024];
    if (!fgets(buf, sizeof(buf), f_)) return false;
    out = buf;
    return true;
  }
  double number() {
    std::string l;
    if (!line(l)) { bad_ = true; return 0.0; }
    const char* p = l.c_str();
    while (isspace((unsigned char)*p)) ++p;
    if (!strncmp(p, "inf", 3)) return kThalInf;
    return strtod(p, nullptr);
  }
  // index, interior, bulge, hairpin (readLoop, thal.h:417-446)
  void loop_row(double& interior, double& bulge) {
    std::string l;
    if (!line(l)) { bad_ = true; return; }
    std::istringstream is(l);
    std::string idx, a, b;
    is >> idx >> a >> b;
    interior = a == "inf" ? kThalInf : strtod(a.c_str(), nullptr);
    bulge = b == "inf" ? kThalInf : strtod(b.c_str(), nullptr);
  }
  bool bad() const { return bad_ || !f_; }

 private:
  FILE* f_;
  bool bad_ = false;
};

// get_thermodynamic_values (thal.h:2368-2393) for the tables the duplex path reads, plus the two
// constants that depend on the run's conditions: saltCorrectS (thal.h:354-359) and R ln(c) (:2504-2508).
inline bool thal_params_from_config(const std::string& dir_in, double mv, double dv, double dntp, double dna_conc, ThalParams& tp,
                                    std::string& err) {
  std::string dir = dir_in;
  if (!dir.empty() && dir.back() != '/') dir += '/';
  memset(&tp, 0, sizeof(tp));
  auto fin = [](double x) { return x < kThalInf / 2; };
  auto four = [&](const char* sname, const char* hname, double* S, double* H, bool terminal) -> bool {
    ThalFile fs(dir, sname), fh(dir, hname);
    if (!fs.ok() || !fh.ok()) { err = std::string("cannot open ") + dir + sname + " / " + hname; return false; }
    for (int i = 0; i < 5; ++i)
      for (int ii = 0; ii < 5; ++ii)
        for (int j = 0; j < 5; ++j)
          for (int jj = 0; jj < 5; ++jj) {
            const int k = ((i * 5 + ii) * 5 + j) * 5 + jj;
            if (!terminal ? (i == 4 || j == 4 || ii == 4 || jj == 4) : (i == 4 || j == 4)) {
              S[k] = -1.0; H[k] = kThalInf;
            } else if (terminal && (ii == 4 || jj == 4)) {
              S[k] = 0.00000000001; H[k] = 0.0;
            } else {
              S[k] = fs.number(); H[k] = fh.number();
              if (!fin(S[k]) || !fin(H[k])) { S[k] = -1.0; H[k] = kThalInf; }
            }
          }
    if (fs.bad() || fh.bad()) { err = std::string("truncated parameter file ") + sname + " / " + hname; return false; }
    return true;
  };
  if (!four("stack.ds", "stack.dh", tp.stackS, tp.stackH, false)) return false;
  if (!four("stackmm.ds", "stackmm.dh", tp.int2S, tp.int2H, false)) return false;
  if (!four("tstack_tm_inf.ds", "tstack.dh", tp.tstackS, tp.tstackH, true)) return false;
  if (!four("tstack2.ds", "tstack2.dh", tp.tstack2S, tp.tstack2H, true)) return false;
  {
    ThalFile fs(dir, "dangle.ds"), fh(dir, "dangle.dh");
    if (!fs.ok() || !fh.ok()) { err = "cannot open " + dir + "dangle.ds / dangle.dh"; return false; }
    for (int pass = 0; pass < 2; ++pass) {
      double* S = pass ? tp.dangS5 : tp.dangS3;
      double* H = pass ? tp.dangH5 : tp.dangH3;
      for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j)
          for (int k = 0; k < 5; ++k) {
            const int at = pass ? (i * 5 + j) * 5 + k : (i * 5 + k) * 5 + j;   // [i][j][k] for 5', [i][k][j] for 3'
            if (i == 4 || j == 4 || k == 4) {
              S[at] = -1.0; H[at] = kThalInf;
            } else {
              S[at] = fs.number(); H[at] = fh.number();
              if (!fin(S[at]) || !fin(H[at])) { S[at] = -1.0; H[at] = kThalInf; }
            }
          }
    }
    if (fs.bad() || fh.bad()) { err = "truncated parameter file dangle.ds / dangle.dh"; return false; }
  }
  {
    ThalFile fs(dir, "loops.ds"), fh(dir, "loops.dh");
    if (!fs.ok() || !fh.ok()) { err = "cannot open " + dir + "loops.ds / loops.dh"; return false; }
    for (int k = 0; k < 30; ++k) {
      fs.loop_row(tp.intlS[k], tp.bulgeS[k]);
      fh.loop_row(tp.intlH[k], tp.bulgeH[k]);
    }
    if (fs.bad() || fh.bad()) { err = "truncated parameter file loops.ds / loops.dh"; return false; }
  }
  for (int i = 0; i < 25; ++i) { tp.atpS[i] = 0.00000000001; tp.atpH[i] = 0.0; }
  tp.atpS[0 * 5 + 3] = tp.atpS[3 * 5 + 0] = 6.9;      // AT_S
  tp.atpH[0 * 5 + 3] = tp.atpH[3 * 5 + 0] = 2200.0;   // AT_H
  if (dv <= 0) dntp = dv;
  tp.salt = 0.368 * ((log((mv + 120 * (sqrt(fmax(0.0, dv - dntp)))) / 1000)));
  tp.rc[0] = 1.9872 * log(dna_conc / 1000000000.0);
  tp.rc[1] = 1.9872 * log(dna_conc / 4000000000.0);
  return true;
}

}  // namespace dg
