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

}  // namespace dg
