// nbr_trunc.hpp -- the reference's neighbourhood generation replayed in its own order, for the
// queries whose neighbourhood may reach the cap -x (max_neighborhood).
//
// neighbors() (reference src/neighbors.h:86-92) runs a depth-first search over edit sequences and
// feeds every generated string to _insert (neighbors.h:29-45), which keeps the set an antichain
// under "is a substring of" ONLINE: inserting s erases every element that contains s, and s itself
// is dropped when an element is contained in it.  The search stops as soon as the set holds
// maxsize strings (neighbors.h:50, checked at the entry of every recursive call), so a truncated
// result is the state of that online process at the moment its size first reaches maxsize -- it
// depends on the visiting order, and it is NOT the set of substring-minimal strings the device
// path computes for the untruncated case (an element may survive only because the shorter string
// that would erase it is generated after the cut).
//
// This file restates that process with the same visiting order (deletion, no change, substitutions
// in alphabet order, insertions in alphabet order; neighbors.h:52-78) on top of a hashed substring
// index, so that one insertion costs O((2d+1)^2) look-ups instead of a scan of the whole set with
// two std::string::find per element.  Host code; plain C++; no CUDA.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

namespace dg {

class NeighborReplay {
 public:
  // Runs neighbors(query, {A,C,G,T}, dist, indel, maxsize, set) and leaves the resulting set in
  // strings() (unordered).  Returns true when the set reached maxsize (the condition of the warning
  // of hunter.h:342-345 for this strand).  peak() = the largest size the set ever had.
  bool run(const std::string& query, int dist, bool indel, uint32_t maxsize) {
    q0_ = query;
    dist_ = dist;
    indel_ = indel;
    maxsize_ = maxsize;
    minlen_ = (int)query.size() - dist;
    if (minlen_ < 1) minlen_ = 1;
    elems_.clear();
    alive_.clear();
    by_string_.clear();
    by_sub_.clear();
    size_ = 0;
    peak_ = 0;
    leaves_ = 0;
    std::string q(query);
    insert(q);                        // neighbors.h:89
    dfs(q, dist, 0);                  // neighbors.h:90
    return size_ >= maxsize_;
  }
  std::vector<std::string> strings() const {
    std::vector<std::string> out;
    out.reserve(size_);
    for (size_t i = 0; i < elems_.size(); ++i) if (alive_[i]) out.push_back(elems_[i]);
    return out;
  }
  uint32_t size() const { return size_; }
  uint32_t peak() const { return peak_; }
  uint64_t leaves() const { return leaves_; }

 private:
  std::string q0_;
  int dist_ = 0, minlen_ = 1;
  bool indel_ = true;
  uint32_t maxsize_ = 0, size_ = 0, peak_ = 0;
  uint64_t leaves_ = 0;
  std::vector<std::string> elems_;
  std::vector<char> alive_;
  std::unordered_map<std::string, uint32_t> by_string_;               // element -> id (alive or dead)
  std::unordered_map<std::string, std::vector<uint32_t>> by_sub_;    // substring (length >= minlen) -> elements holding it

  // _insert (neighbors.h:29-45)
  void insert(const std::string& s) {
    if (!indel_) {   // a plain std::set insert
      auto it = by_string_.find(s);
      if (it == by_string_.end()) add(s);
      return;
    }
    {
      auto it = by_string_.find(s);
      if (it != by_string_.end() && alive_[it->second]) return;   // erased and inserted again: no change
    }
    // every element that contains s goes ...
    auto sub = by_sub_.find(s);
    if (sub != by_sub_.end()) {
      for (uint32_t id : sub->second)
        if (alive_[id]) { alive_[id] = 0; --size_; }
      sub->second.clear();
    }
    // ... and s stays out when an element is contained in it (proper substrings; lengths >= minlen,
    // nothing shorter is ever generated)
    const int L = (int)s.size();
    for (int len = minlen_; len < L; ++len)
      for (int a = 0; a + len <= L; ++a) {
        auto it = by_string_.find(s.substr(a, len));
        if (it != by_string_.end() && alive_[it->second]) return;
      }
    add(s);
  }
  void add(const std::string& s) {
    uint32_t id;
    auto it = by_string_.find(s);
    if (it != by_string_.end()) {
      id = it->second;
      alive_[id] = 1;
    } else {
      id = (uint32_t)elems_.size();
      elems_.push_back(s);
      alive_.push_back(1);
      by_string_.emplace(s, id);
    }
    ++size_;
    if (size_ > peak_) peak_ = size_;
    if (indel_) {
      const int L = (int)s.size();
      for (int len = minlen_; len <= L; ++len)
        for (int a = 0; a + len <= L; ++a) by_sub_[s.substr(a, len)].push_back(id);
    }
  }

  // _neighbors (neighbors.h:47-83)
  void dfs(std::string& query, int dist, int pos) {
    if (size_ >= maxsize_) return;
    if (pos < (int)query.size()) {
      if (dist > 0 && indel_) {   // deletion
        std::string shorter = query.substr(0, pos) + query.substr(pos + 1);
        dfs(shorter, dist - 1, pos);
      }
      dfs(query, dist, pos + 1);   // no change
      if (dist > 0) {
        const char orig = query[pos];
        for (const char c : {'A', 'C', 'G', 'T'}) {
          if (c == orig) continue;
          query[pos] = c;
          dfs(query, dist - 1, pos + 1);
        }
        query[pos] = orig;
        if (indel_)
          for (const char c : {'A', 'C', 'G', 'T'}) {
            std::string longer = query.substr(0, pos) + std::string(1, c) + query.substr(pos);
            dfs(longer, dist - 1, pos + 1);
          }
      }
    } else {
      ++leaves_;
      if (dist < dist_) insert(query);   // only true neighbours
    }
  }
};

// The same replay for ACGT-only queries whose strings fit one 64-bit word (length + distance <= 29):
// strings are 2-bit codes (first base in the high bits of the used part) tagged with their length,
// the two hash maps are open-addressing tables of 64-bit keys.  Same visiting order, same online
// antichain, ~25 times faster than the std::string form (which stays for queries holding N or longer
// ones; tests/hostsim compares the two on every golden query).
class NeighborReplayPacked {
 public:
  static bool fits(const std::string& query, int dist) {
    if ((int)query.size() + dist > 29 || query.empty()) return false;
    for (char c : query) if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return false;
    return true;
  }
  bool run(const std::string& query, int dist, bool indel, uint32_t maxsize) {
    m_ = (int)query.size();
    dist_ = dist;
    indel_ = indel;
    maxsize_ = maxsize;
    minlen_ = m_ - dist > 1 ? m_ - dist : 1;
    size_ = peak_ = 0;
    elem_key_.clear();
    alive_.clear();
    // table sizes: every leaf may add one element and (2 d + 1)(2 d + 2) / 2 substring entries
    uint64_t leaves_ub = 1;
    for (int i = 0; i < dist; ++i) leaves_ub *= (uint64_t)(9 * (m_ + dist));
    uint64_t want = std::min<uint64_t>(leaves_ub + 16, (uint64_t)maxsize + 64) ;
    if (!indel) want = std::min<uint64_t>(want, (uint64_t)maxsize + 64);
    const int per = indel ? (2 * dist + 1) * (2 * dist + 2) / 2 + 1 : 1;
    resize_tables(want * 4, want * (uint64_t)per * 4);
    uint64_t code = 0;
    for (char c : query) code = (code << 2) | (uint64_t)(c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3);
    insert(code, m_);
    dfs(code, m_, dist, 0);
    return size_ >= maxsize_;
  }
  std::vector<std::string> strings() const {
    std::vector<std::string> out;
    out.reserve(size_);
    for (size_t i = 0; i < elem_key_.size(); ++i) {
      if (!alive_[i]) continue;
      const int L = (int)(elem_key_[i] >> 58);
      const uint64_t code = elem_key_[i] & ((1ULL << 58) - 1);
      std::string s((size_t)L, 'A');
      for (int j = 0; j < L; ++j) s[(size_t)j] = "ACGT"[(code >> (2 * (L - 1 - j))) & 3];
      out.push_back(s);
    }
    return out;
  }
  uint32_t size() const { return size_; }
  uint32_t peak() const { return peak_; }

 private:
  int m_ = 0, dist_ = 0, minlen_ = 1;
  bool indel_ = true;
  uint32_t maxsize_ = 0, size_ = 0, peak_ = 0;
  std::vector<uint64_t> elem_key_;       // id -> key
  std::vector<char> alive_;
  // open addressing, key 0 = empty (keys carry a length >= 1 in the top bits, so 0 never occurs)
  std::vector<uint64_t> ekeys_;          // element table: key -> id
  std::vector<uint32_t> eids_;
  std::vector<uint64_t> skeys_;          // substring table: (key, element id) pairs, several per key
  std::vector<uint32_t> sids_;
  uint64_t emask_ = 0, smask_ = 0, eused_ = 0, sused_ = 0;

  static uint64_t key_of(uint64_t code, int len) { return ((uint64_t)len << 58) | code; }
  static uint64_t hash(uint64_t k) { k ^= k >> 31; k *= 0x9E3779B97F4A7C15ULL; k ^= k >> 29; return k; }
  static uint64_t pow2_at_least(uint64_t n) { uint64_t p = 64; while (p < n) p <<= 1; return p; }
  void resize_tables(uint64_t ne, uint64_t ns) {
    ne = pow2_at_least(ne); ns = pow2_at_least(ns);
    ekeys_.assign(ne, 0); eids_.assign(ne, 0); emask_ = ne - 1; eused_ = 0;
    skeys_.assign(ns, 0); sids_.assign(ns, 0); smask_ = ns - 1; sused_ = 0;
  }
  void grow_elements() {
    std::vector<uint64_t> ok; std::vector<uint32_t> oi;
    ok.swap(ekeys_); oi.swap(eids_);
    const uint64_t ne = ok.size() * 2;
    ekeys_.assign(ne, 0); eids_.assign(ne, 0); emask_ = ne - 1;
    for (size_t i = 0; i < ok.size(); ++i)
      if (ok[i]) { uint64_t h = hash(ok[i]) & emask_; while (ekeys_[h]) h = (h + 1) & emask_; ekeys_[h] = ok[i]; eids_[h] = oi[i]; }
  }
  void grow_subs() {
    std::vector<uint64_t> ok; std::vector<uint32_t> oi;
    ok.swap(skeys_); oi.swap(sids_);
    const uint64_t ns = ok.size() * 2;
    skeys_.assign(ns, 0); sids_.assign(ns, 0); smask_ = ns - 1;
    for (size_t i = 0; i < ok.size(); ++i)
      if (ok[i]) { uint64_t h = hash(ok[i]) & smask_; while (skeys_[h]) h = (h + 1) & smask_; skeys_[h] = ok[i]; sids_[h] = oi[i]; }
  }
  int find_element(uint64_t key) const {   // id or -1
    uint64_t h = hash(key) & emask_;
    while (ekeys_[h]) {
      if (ekeys_[h] == key) return (int)eids_[h];
      h = (h + 1) & emask_;
    }
    return -1;
  }
  void insert(uint64_t code, int len) {
    const uint64_t key = key_of(code, len);
    const int known = find_element(key);
    if (!indel_) {
      if (known < 0) add(key, code, len, known);
      return;
    }
    if (known >= 0 && alive_[(size_t)known]) return;
    // every element that contains s goes
    {
      uint64_t h = hash(key) & smask_;
      while (skeys_[h]) {
        if (skeys_[h] == key) {
          const uint32_t id = sids_[h];
          if (alive_[id]) { alive_[id] = 0; --size_; }
        }
        h = (h + 1) & smask_;
      }
    }
    // s stays out when an element is contained in it
    for (int l2 = minlen_; l2 < len; ++l2) {
      const uint64_t mask = (1ULL << (2 * l2)) - 1;
      for (int a = 0; a + l2 <= len; ++a) {
        const uint64_t sub = (code >> (2 * (len - l2 - a))) & mask;
        const int id = find_element(key_of(sub, l2));
        if (id >= 0 && alive_[(size_t)id]) return;
      }
    }
    add(key, code, len, known);
  }
  void add(uint64_t key, uint64_t code, int len, int known) {
    uint32_t id;
    if (known >= 0) {
      id = (uint32_t)known;
      alive_[id] = 1;
    } else {
      id = (uint32_t)elem_key_.size();
      elem_key_.push_back(key);
      alive_.push_back(1);
      if ((eused_ + 1) * 2 > ekeys_.size()) grow_elements();
      uint64_t h = hash(key) & emask_;
      while (ekeys_[h]) h = (h + 1) & emask_;
      ekeys_[h] = key; eids_[h] = id; ++eused_;
      if (indel_) {   // its substrings (itself included), once per element
        for (int l2 = minlen_; l2 <= len; ++l2) {
          const uint64_t mask = (1ULL << (2 * l2)) - 1;
          for (int a = 0; a + l2 <= len; ++a) {
            const uint64_t sk = key_of((code >> (2 * (len - l2 - a))) & mask, l2);
            if ((sused_ + 1) * 2 > skeys_.size()) grow_subs();
            uint64_t h2 = hash(sk) & smask_;
            while (skeys_[h2]) h2 = (h2 + 1) & smask_;
            skeys_[h2] = sk; sids_[h2] = id; ++sused_;
          }
        }
      }
    }
    ++size_;
    if (size_ > peak_) peak_ = size_;
  }
  // _neighbors (neighbors.h:47-83) on (code, len): position pos counts from the left
  void dfs(uint64_t code, int len, int dist, int pos) {
    if (size_ >= maxsize_) return;
    if (pos < len) {
      const int sh = 2 * (len - 1 - pos);
      const uint64_t low = code & ((1ULL << sh) - 1), high = (code >> sh) >> 2;
      const uint64_t orig = (code >> sh) & 3;
      if (dist > 0 && indel_) dfs((high << sh) | low, len - 1, dist - 1, pos);   // deletion
      dfs(code, len, dist, pos + 1);                                             // no change
      if (dist > 0) {
        for (uint64_t c = 0; c < 4; ++c)
          if (c != orig) dfs((high << (sh + 2)) | (c << sh) | low, len, dist - 1, pos + 1);
        if (indel_)
          for (uint64_t c = 0; c < 4; ++c)
            dfs((high << (sh + 4)) | (c << (sh + 2)) | (orig << sh) | low, len + 1, dist - 1, pos + 1);
      }
    } else if (dist < dist_) {
      insert(code, len);
    }
  }
};

}  // namespace dg
