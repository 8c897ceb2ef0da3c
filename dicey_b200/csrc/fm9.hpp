// fm9.hpp -- independent reader / writer of dicey's on-disk FM-index (`.fm9` + `.fm9_check`).
//
// The file is the raw SDSL serialization of
//   csa_wt<wt_huff<bit_vector, rank_support_v<1>, select_support_mcl<1>, select_support_mcl<0>,
//          byte_tree<>>, 32, 64, sa_order_sa_sampling<>, isa_sampling<>, byte_alphabet>
// (reference: src/index.h:79,122; member order src/xxsds/include/sdsl/csa_wt.hpp:362-373,
// wt_pc.hpp:611-624, rank_support_v.hpp:122-129, select_support_mcl.hpp:427-467,
// wt_helper.hpp:313-326, csa_sampling_strategy.hpp:70-93,669-706,
// csa_alphabet_strategy.hpp:233-244, int_vector.hpp:813-842,1812-1838).  Nothing of SDSL is
// included or linked; this is a byte-level parser of the layout written down in SURVEY.md 5.9.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

namespace dg {

struct Fm9Node {  // wt_helper.hpp:121-125 (_node<byte_tree>), 22 bytes on disk
  uint64_t bv_pos;
  uint64_t bv_pos_rank;  // leaves keep their symbol here
  uint16_t parent;
  uint16_t child[2];
};

// The large payloads stay where the kernel's page cache has them: the file is mapped read-only
// and the spans below point into the mapping (they may be unaligned: copy, do not dereference).
struct Fm9Span {
  const uint8_t* p = nullptr;
  uint64_t words = 0;   // 64-bit words
  uint64_t bytes() const { return words * 8; }
};

struct Fm9 {
  uint64_t n = 0;         // wt.m_size == csa.size() (text length incl. sentinel)
  uint64_t wt_sigma = 0;  // wt.m_sigma
  uint64_t bv_bits = 0;
  Fm9Span bv;                      // m_bv words
  Fm9Span rank_bb;                 // rank_support_v basic blocks
  std::vector<Fm9Node> nodes;      // byte_tree nodes (BFS order, root = 0)
  uint16_t c_to_leaf[256];
  uint64_t path[256];
  uint8_t sa_width = 0, isa_width = 0;
  uint64_t sa_count = 0, isa_count = 0;
  Fm9Span sa_words, isa_words;     // packed int_vector<0> payloads
  uint8_t char2comp[256];
  std::vector<uint8_t> comp2char;
  std::vector<uint64_t> C;  // sigma + 1 entries
  uint16_t sigma = 0;

  Fm9() = default;
  Fm9(const Fm9&) = delete;
  Fm9& operator=(const Fm9&) = delete;
  ~Fm9() { unmap(); }
  void unmap() {
    if (map_) munmap(map_, map_bytes_);
    map_ = nullptr;
    map_bytes_ = 0;
    bv = rank_bb = sa_words = isa_words = Fm9Span();
  }
  void* map_ = nullptr;
  size_t map_bytes_ = 0;
};

// The type name SDSL hashes into `.fm9_check`: util::demangle2 of the csa_wt<> type
// (io.hpp:789-793, util.hpp:309-316), hashed with std::hash<std::string>.
inline const char* fm9_type_name() {
  return "csa_wt<wt_pc<huff_shape, bit_vector, rank_support_v<1, 1>, select_support_mcl<1, 1>, "
         "select_support_mcl<0, 1>, byte_tree<false> >, 32u, 64u, sa_order_sa_sampling<0>, "
         "isa_sampling<0>, byte_alphabet>";
}
inline uint64_t fm9_type_hash() { return std::hash<std::string>()(fm9_type_name()); }

class Fm9Reader {
 public:
  explicit Fm9Reader(FILE* f) : f_(f) {}
  bool u64(uint64_t& v) { return fread(&v, 8, 1, f_) == 1; }
  bool u16(uint16_t& v) { return fread(&v, 2, 1, f_) == 1; }
  bool raw(void* p, size_t bytes) { return bytes == 0 || fread(p, 1, bytes, f_) == bytes; }
  bool skip(uint64_t bytes) { return fseeko(f_, (off_t)bytes, SEEK_CUR) == 0; }
  // int_vector<*>: header (width << 56 | bits), then ceil(bits/64) words.
  bool int_vector(std::vector<uint64_t>* words, uint64_t& bits, uint8_t& width) {
    uint64_t h;
    if (!u64(h)) return false;
    bits = h & ((1ULL << 56) - 1);
    width = (uint8_t)(h >> 56);
    uint64_t nw = (bits + 63) >> 6;
    if (words) {
      words->resize(nw);
      return raw(words->data(), nw * 8);
    }
    return skip(nw * 8);
  }

 private:
  FILE* f_;
};

inline uint64_t fm9_get_int(const std::vector<uint64_t>& w, uint64_t i, uint8_t width) {
  uint64_t bit = i * width, word = bit >> 6, off = bit & 63;
  uint64_t v = w[word] >> off;
  if (off + width > 64) v |= w[word + 1] << (64 - off);
  return width == 64 ? v : (v & ((1ULL << width) - 1));
}

// select_support_mcl<b>::load (select_support_mcl.hpp:470-499): parsed only to be skipped.
inline bool fm9_skip_select(Fm9Reader& rd) {
  uint64_t arg_cnt;
  if (!rd.u64(arg_cnt)) return false;
  if (!arg_cnt) return true;
  uint64_t sb = (arg_cnt + 4095) >> 12, bits;
  uint8_t w;
  if (!rd.int_vector(nullptr, bits, w)) return false;  // superblock
  std::vector<uint64_t> mol;
  uint64_t molbits;
  if (!rd.int_vector(&mol, molbits, w)) return false;  // mini_or_long
  for (uint64_t i = 0; i < sb; ++i)
    if (!rd.int_vector(nullptr, bits, w)) return false;  // long or mini block: same framing
  return true;
}

// Cursor over the mapped file.
class Fm9Cursor {
 public:
  Fm9Cursor(const uint8_t* p, uint64_t bytes) : p_(p), end_(p + bytes) {}
  bool raw(void* out, uint64_t bytes) {
    if ((uint64_t)(end_ - p_) < bytes) return false;
    memcpy(out, p_, bytes);
    p_ += bytes;
    return true;
  }
  bool u64(uint64_t& v) { return raw(&v, 8); }
  bool u16(uint16_t& v) { return raw(&v, 2); }
  // int_vector<*>: header (width << 56 | bits), then ceil(bits/64) words
  bool int_vector(Fm9Span* span, uint64_t& bits, uint8_t& width) {
    uint64_t h;
    if (!u64(h)) return false;
    bits = h & ((1ULL << 56) - 1);
    width = (uint8_t)(h >> 56);
    const uint64_t nw = (bits + 63) >> 6;
    if ((uint64_t)(end_ - p_) / 8 < nw) return false;
    if (span) { span->p = p_; span->words = nw; }
    p_ += nw * 8;
    return true;
  }
  bool int_vector_copy(std::vector<uint64_t>& words, uint64_t& bits, uint8_t& width) {
    Fm9Span sp;
    if (!int_vector(&sp, bits, width)) return false;
    words.resize(sp.words);
    if (sp.words) memcpy(words.data(), sp.p, sp.bytes());
    return true;
  }
  // select_support_mcl<b>::load (select_support_mcl.hpp:470-499): framing only
  bool skip_select() {
    uint64_t arg_cnt, bits;
    uint8_t w;
    if (!u64(arg_cnt)) return false;
    if (!arg_cnt) return true;
    const uint64_t sb = (arg_cnt + 4095) >> 12;
    if (!int_vector(nullptr, bits, w)) return false;   // superblock
    if (!int_vector(nullptr, bits, w)) return false;   // mini_or_long
    for (uint64_t i = 0; i < sb; ++i)
      if (!int_vector(nullptr, bits, w)) return false; // long or mini block: same framing
    return true;
  }
  bool at_end() const { return p_ == end_; }

 private:
  const uint8_t* p_;
  const uint8_t* end_;
};

// Returns 0 on success; -2 I/O, -3 format (codes of include/dicey_b200.h).
inline int fm9_parse(const std::string& path, Fm9& o, std::string& err, bool check_sidecar = true) {
  if (check_sidecar) {
    // load_from_checked_file (io.hpp:917-936): the sidecar must exist and hold the type hash.
    FILE* c = fopen((path + "_check").c_str(), "rb");
    uint64_t h = 0;
    if (!c || fread(&h, 8, 1, c) != 1) {
      if (c) fclose(c);
      err = "cannot read " + path + "_check";
      return -2;
    }
    fclose(c);
    if (h != fm9_type_hash()) {
      err = path + "_check does not hold the csa_wt<> type hash";
      return -3;
    }
  }
  o.unmap();
  {
    int fd = open(path.c_str(), O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
      if (fd >= 0) close(fd);
      err = "cannot open " + path;
      return -2;
    }
    if (st.st_size == 0) {
      close(fd);
      err = "malformed .fm9 (wt header): " + path;
      return -3;
    }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) {
      err = "cannot map " + path;
      return -2;
    }
    o.map_ = m;
    o.map_bytes_ = (size_t)st.st_size;
  }
  Fm9Cursor rd((const uint8_t*)o.map_, o.map_bytes_);
  auto fail = [&](const char* what) {
    err = std::string("malformed .fm9 (") + what + "): " + path;
    o.unmap();
    return -3;
  };
  uint8_t w;
  uint64_t bits;
  // 1. wavelet tree (wt_pc.hpp:627-636)
  if (!rd.u64(o.n) || !rd.u64(o.wt_sigma)) return fail("wt header");
  if (!rd.int_vector(&o.bv, o.bv_bits, w) || w != 1) return fail("m_bv");
  if (!rd.int_vector(&o.rank_bb, bits, w) || w != 64) return fail("rank_support_v");
  if (o.rank_bb.words != 2 * (((o.bv_bits + 63) >> 9) + 1) && !(o.bv_bits == 0 && o.rank_bb.words == 2))
    return fail("rank_support_v size");
  if (!rd.skip_select() || !rd.skip_select()) return fail("select_support_mcl");
  uint64_t nn;
  if (!rd.u64(nn) || nn > 511) return fail("byte_tree size");
  o.nodes.resize(nn);
  for (uint64_t i = 0; i < nn; ++i) {
    uint8_t rec[22];
    if (!rd.raw(rec, 22)) return fail("byte_tree node");
    memcpy(&o.nodes[i].bv_pos, rec, 8);
    memcpy(&o.nodes[i].bv_pos_rank, rec + 8, 8);
    memcpy(&o.nodes[i].parent, rec + 16, 2);
    memcpy(&o.nodes[i].child[0], rec + 18, 2);
    memcpy(&o.nodes[i].child[1], rec + 20, 2);
  }
  if (!rd.raw(o.c_to_leaf, sizeof(o.c_to_leaf)) || !rd.raw(o.path, sizeof(o.path))) return fail("byte_tree maps");
  // 2./3. SA and ISA samples (int_vector<0>)
  if (!rd.int_vector(&o.sa_words, bits, o.sa_width) || o.sa_width == 0) return fail("sa_sample");
  o.sa_count = bits / o.sa_width;
  if (!rd.int_vector(&o.isa_words, bits, o.isa_width) || o.isa_width == 0) return fail("isa_sample");
  o.isa_count = bits / o.isa_width;
  // 4. byte_alphabet
  std::vector<uint64_t> tmp;
  if (!rd.int_vector_copy(tmp, bits, w) || w != 8 || bits != 256 * 8) return fail("char2comp");
  memcpy(o.char2comp, tmp.data(), 256);
  if (!rd.int_vector_copy(tmp, bits, w) || w != 8) return fail("comp2char");
  o.comp2char.assign((uint8_t*)tmp.data(), (uint8_t*)tmp.data() + bits / 8);
  if (!rd.int_vector_copy(o.C, bits, w) || w != 64) return fail("C");
  if (!rd.u16(o.sigma)) return fail("sigma");
  if (o.C.size() != (size_t)o.sigma + 1 || o.comp2char.size() != o.sigma) return fail("alphabet sizes");
  if (o.sa_count != (o.n + 31) / 32 || o.isa_count != (o.n ? (o.n - 1) / 64 + 1 : 0)) return fail("sample counts");
  if (!rd.at_end()) return fail("trailing bytes");
  return 0;
}

}  // namespace dg
