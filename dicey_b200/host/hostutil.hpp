// hostutil.hpp -- the small pieces of Boost / htslib / nlohmann behaviour dicey's drivers rely on,
// restated on libstdc++ + zlib (neither Boost nor htslib can be built in this image):
//   Options      boost::program_options command_line_parser, default unix style
//                (hunter.h:185-212, index.h:37-58): -x v, -xv, --name v, --name=v, unique
//                long-option prefixes, sticky short flags, positional arguments
//   path_*       boost::filesystem::path::parent_path / stem / is_regular_file / file_size
//   json_*       nlohmann::json::dump() of flat objects: keys in std::map order, no blanks,
//                nlohmann's string escaping (json.hpp serializer::dump_escaped, ensure_ascii = false)
//   GzLines      boost::iostreams gzip_decompressor + std::getline (index.h:95-111), also plain text
//   gz_append    filtering_ostream(gzip_compressor, file_sink(app)) of hunter.h:165-170: every
//                call appends one gzip member
//   read_fai     getSeqLenName (util.h:183-206): faidx names and lengths (+1 is applied by the caller)
#pragma once
#include <sys/stat.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace dhost {

// ------------------------------------------------------------------------------------------
struct OptSpec {
  std::string long_name;
  char short_name;   // 0 = none
  bool takes_value;
};

class Options {
 public:
  explicit Options(std::vector<OptSpec> specs) : specs_(std::move(specs)) {}
  // throws std::runtime_error on unknown / ambiguous options or a missing value
  void parse(int argc, char** argv) {
    bool only_positional = false;
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      if (only_positional || a.size() < 2 || a[0] != '-') { positional.push_back(a); continue; }
      if (a == "--") { only_positional = true; continue; }
      if (a[1] == '-') {
        std::string name = a.substr(2), value;
        bool has_value = false;
        size_t eq = name.find('=');
        if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); has_value = true; }
        const OptSpec& s = find_long(name);
        if (s.takes_value) {
          if (!has_value) {
            if (i + 1 >= argc) throw std::runtime_error("the required argument for option '--" + s.long_name + "' is missing");
            value = argv[++i];
          }
          values[s.long_name] = value;
        } else {
          if (has_value) throw std::runtime_error("option '--" + s.long_name + "' does not take any arguments");
          values[s.long_name] = "";
        }
        continue;
      }
      // short options, possibly sticky: -nf, -d1, -d 1
      for (size_t k = 1; k < a.size(); ++k) {
        const OptSpec& s = find_short(a[k]);
        if (!s.takes_value) { values[s.long_name] = ""; continue; }
        std::string value = a.substr(k + 1);
        if (value.empty()) {
          if (i + 1 >= argc) throw std::runtime_error("the required argument for option '--" + s.long_name + "' is missing");
          value = argv[++i];
        }
        values[s.long_name] = value;
        break;
      }
    }
  }
  bool has(const std::string& n) const { return values.count(n) != 0; }
  std::string get(const std::string& n, const std::string& dflt = "") const {
    auto it = values.find(n);
    return it == values.end() ? dflt : it->second;
  }
  uint64_t get_u64(const std::string& n, uint64_t dflt) const {
    if (!has(n)) return dflt;
    const std::string v = get(n);
    size_t pos = 0;
    unsigned long long x = 0;
    try { x = std::stoull(v, &pos); } catch (...) { pos = 0; }
    if (pos != v.size() || v.empty() || v[0] == '-')
      throw std::runtime_error("the argument ('" + v + "') for option '--" + n + "' is invalid");
    return x;
  }
  double get_double(const std::string& n, double dflt) const {
    if (!has(n)) return dflt;
    const std::string v = get(n);
    size_t pos = 0;
    double x = 0;
    try { x = std::stod(v, &pos); } catch (...) { pos = 0; }
    if (pos != v.size() || v.empty()) throw std::runtime_error("the argument ('" + v + "') for option '--" + n + "' is invalid");
    return x;
  }
  std::vector<std::string> positional;
  std::map<std::string, std::string> values;

 private:
  const OptSpec& find_long(const std::string& name) const {
    const OptSpec* hit = nullptr;
    for (const auto& s : specs_) {
      if (s.long_name == name) return s;
      if (s.long_name.compare(0, name.size(), name) == 0) {
        if (hit) throw std::runtime_error("option '--" + name + "' is ambiguous");
        hit = &s;
      }
    }
    if (!hit) throw std::runtime_error("unrecognised option '--" + name + "'");
    return *hit;
  }
  const OptSpec& find_short(char c) const {
    for (const auto& s : specs_) if (s.short_name == c) return s;
    throw std::runtime_error(std::string("unrecognised option '-") + c + "'");
  }
  std::vector<OptSpec> specs_;
};

// ------------------------------------------------------------------------------------------
inline std::string path_parent(const std::string& p) {
  size_t s = p.find_last_of('/');
  if (s == std::string::npos) return "";
  if (s == 0) return "/";
  return p.substr(0, s);
}
inline std::string path_filename(const std::string& p) {
  size_t s = p.find_last_of('/');
  return s == std::string::npos ? p : p.substr(s + 1);
}
inline std::string path_stem(const std::string& p) {
  std::string f = path_filename(p);
  if (f == "." || f == "..") return f;
  size_t d = f.find_last_of('.');
  return d == std::string::npos ? f : f.substr(0, d);
}
inline std::string path_join(const std::string& a, const std::string& b) {
  if (a.empty()) return b;
  if (a.back() == '/') return a + b;
  return a + "/" + b;
}
inline bool is_regular_file(const std::string& p, uint64_t* size = nullptr) {
  struct stat st;
  if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return false;
  if (size) *size = (uint64_t)st.st_size;
  return true;
}
inline bool nonempty_regular_file(const std::string& p) {
  uint64_t sz = 0;
  return is_regular_file(p, &sz) && sz > 0;
}

// ------------------------------------------------------------------------------------------
inline std::string json_escape(const std::string& s) {
  std::string o;
  o.reserve(s.size() + 2);
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\b': o += "\\b"; break;
      case '\f': o += "\\f"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) {
          char buf[8];
          snprintf(buf, sizeof(buf), "\\u%04x", c);
          o += buf;
        } else {
          o += (char)c;
        }
    }
  }
  return o;
}

// A flat JSON object whose values are already-serialised fragments; dump() orders the keys as
// std::map<std::string, ...> does (nlohmann::json's default object type).
class JsonObject {
 public:
  void set_string(const std::string& k, const std::string& v) { kv_[k] = "\"" + json_escape(v) + "\""; }
  void set_uint(const std::string& k, uint64_t v) { kv_[k] = std::to_string(v); }
  void set_int(const std::string& k, int64_t v) { kv_[k] = std::to_string(v); }
  void set_bool(const std::string& k, bool v) { kv_[k] = v ? "true" : "false"; }
  void set_raw(const std::string& k, const std::string& serialised) { kv_[k] = serialised; }
  std::string dump() const {
    std::string o = "{";
    bool first = true;
    for (const auto& it : kv_) {
      if (!first) o += ",";
      first = false;
      o += "\"" + json_escape(it.first) + "\":" + it.second;
    }
    return o + "}";
  }

 private:
  std::map<std::string, std::string> kv_;
};

// ------------------------------------------------------------------------------------------
class GzLines {  // std::getline over a gzip or plain file
 public:
  explicit GzLines(const std::string& path) : f_(gzopen(path.c_str(), "rb")) {
    if (f_) gzbuffer(f_, 1 << 20);
  }
  ~GzLines() { if (f_) gzclose(f_); }
  bool ok() const { return f_ != nullptr; }
  bool getline(std::string& line) {
    line.clear();
    if (!f_) return false;
    char buf[1 << 16];
    bool any = false;
    while (gzgets(f_, buf, sizeof(buf))) {
      any = true;
      size_t n = strlen(buf);
      if (n && buf[n - 1] == '\n') { line.append(buf, n - 1); return true; }
      line.append(buf, n);
    }
    return any;
  }

 private:
  gzFile f_;
};

inline bool gz_append(const std::string& path, const std::string& data) {
  gzFile f = gzopen(path.c_str(), "ab");
  if (!f) return false;
  bool ok = true;
  for (size_t at = 0; ok && at < data.size();) {   // gzwrite takes an unsigned length and returns an int
    const size_t n = data.size() - at < (1u << 30) ? data.size() - at : (1u << 30);
    ok = gzwrite(f, data.data() + at, (unsigned)n) == (int)n;
    at += n;
  }
  return gzclose(f) == Z_OK && ok;
}

inline bool is_gz(const std::string& path) {   // util.h:21-31: the two gzip magic bytes
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  unsigned char m[2] = {0, 0};
  size_t n = fread(m, 1, 2, f);
  fclose(f);
  return n == 2 && m[0] == 0x1f && m[1] == 0x8b;
}

inline bool is_fasta(const std::string& path) {  // util.h:33-52: the first line starts with '>'
  GzLines in(path);
  std::string line;
  if (!in.getline(line)) return false;
  return !line.empty() && line[0] == '>';
}

// ------------------------------------------------------------------------------------------
// faidx names and lengths.  Uses <genome>.fai when present (what fai_load reads), otherwise scans
// the FASTA itself (what fai_build would compute) and writes the .fai when every record has
// uniform line lengths, as htslib does.
inline bool read_fai(const std::string& genome, std::vector<std::string>& names, std::vector<uint64_t>& lens) {
  names.clear();
  lens.clear();
  {
    std::ifstream fai(genome + ".fai");
    if (fai) {
      std::string line;
      while (std::getline(fai, line)) {
        if (line.empty()) continue;
        size_t t1 = line.find('\t');
        if (t1 == std::string::npos) return false;
        size_t t2 = line.find('\t', t1 + 1);
        names.push_back(line.substr(0, t1));
        lens.push_back(std::stoull(line.substr(t1 + 1, t2 == std::string::npos ? std::string::npos : t2 - t1 - 1)));
      }
      return !names.empty();
    }
  }
  GzLines in(genome);
  if (!in.ok()) return false;
  struct Rec { std::string name; uint64_t len = 0, offset = 0, linebases = 0, linewidth = 0; bool uniform = true, short_seen = false; };
  std::vector<Rec> recs;
  std::string line;
  uint64_t pos = 0;
  while (in.getline(line)) {
    uint64_t raw = line.size() + 1;  // the '\n' getline removed
    std::string body = line;
    if (!body.empty() && body.back() == '\r') body.pop_back();
    if (!body.empty() && body[0] == '>') {
      Rec r;
      size_t e = 1;
      while (e < body.size() && !isspace((unsigned char)body[e])) ++e;
      r.name = body.substr(1, e - 1);
      r.offset = pos + raw;
      recs.push_back(r);
    } else if (!recs.empty()) {
      Rec& r = recs.back();
      uint64_t bases = 0;
      for (unsigned char c : body) if (isgraph(c)) ++bases;
      if (bases) {
        if (r.short_seen) r.uniform = false;          // a short line may only be the last one
        if (r.linebases == 0) { r.linebases = bases; r.linewidth = raw; }
        else if (bases != r.linebases || raw != r.linewidth) {
          if (bases > r.linebases) r.uniform = false;
          r.short_seen = true;
        }
        r.len += bases;
      }
    }
    pos += raw;
  }
  if (recs.empty()) return false;
  bool uniform = true;
  for (const auto& r : recs) { names.push_back(r.name); lens.push_back(r.len); uniform = uniform && r.uniform; }
  if (uniform) {
    std::ofstream out(genome + ".fai");
    if (out) for (const auto& r : recs) out << r.name << '\t' << r.len << '\t' << r.offset << '\t' << r.linebases << '\t' << r.linewidth << '\n';
  }
  return true;
}

}  // namespace dhost
