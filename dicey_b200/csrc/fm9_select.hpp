// fm9_select.hpp -- the two select_support_mcl sections of a `.fm9`, so that dg_index_write_fm9
// emits a file that is byte-identical to what `dicey index` (SDSL construct + store_to_checked_file)
// writes, not merely loadable.
//
// What is restated (host code; the bit vector is the wavelet tree's m_bv, already on the host for
// writing): the block structure select_support_mcl builds over a bit_vector
//   * geometry            select_support_mcl.hpp:404-418 (initData: logn, logn^4)
//   * vectors < 100000 b  select_support_mcl.hpp:190-243 (init_slow)
//   * longer vectors      select_support_mcl.hpp:246-341 (init_fast), including its quirks, which the
//                         bytes depend on: a block is closed when its 4033rd argument has been
//                         seen, "last position in the block" then scans up to 64 arguments further
//                         (i.e. it is the first argument of the next block when there is one), the
//                         final partial block is always stored long with the width of the whole
//                         vector and leaves its superblock entry 0, and for zeros the padding bits
//                         of the last word are counted while sampling
//   * serialization       select_support_mcl.hpp:427-467, int_vector header int_vector.hpp:831-842
// Nothing of SDSL is included; count / locate / extract never read these sections, the reference
// only needs them to be present to reproduce its own file.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace dg {

class SelectMclWriter {
 public:
  // bv: ceil(bit_size / 64) words, padding bits zero.  ones = true: select_support_mcl<1>, false: <0>.
  SelectMclWriter(const uint64_t* bv, uint64_t bit_size, bool ones) : bv_(bv), n_(bit_size), ones_(ones) { build(); }

  // The section as it appears in the file.
  const std::vector<uint8_t>& bytes() const { return out_; }

 private:
  static const uint64_t kSuper = 4096;

  static uint32_t hi(uint64_t x) { return x ? 63u - (uint32_t)__builtin_clzll(x) : 0u; }
  static uint32_t sel(uint64_t w, uint32_t i) {  // position of the i-th (1-based) set bit
    for (uint32_t k = 1; k < i; ++k) w &= w - 1;
    return (uint32_t)__builtin_ctzll(w);
  }
  bool arg(uint64_t i) const { return (((bv_[i >> 6] >> (i & 63)) & 1ULL) != 0) == ones_; }
  uint64_t word(uint64_t wi) const { return ones_ ? bv_[wi] : ~bv_[wi]; }

  void put_u64(std::vector<uint8_t>& o, uint64_t v) {
    size_t at = o.size();
    o.resize(at + 8);
    memcpy(o.data() + at, &v, 8);
  }
  // int_vector<0>(count, 0, width) with the given leading values, the rest zero
  void put_int_vector(std::vector<uint8_t>& o, const uint64_t* vals, uint64_t nvals, uint64_t count, uint8_t width) {
    const uint64_t bits = count * width, nw = (bits + 63) >> 6;
    put_u64(o, ((uint64_t)width << 56) | bits);
    size_t at = o.size();
    o.resize(at + nw * 8, 0);
    std::vector<uint64_t>& tmp = scratch_;
    tmp.assign(nw, 0);
    for (uint64_t i = 0; i < nvals; ++i) {
      const uint64_t b = i * width, wd = b >> 6, off = b & 63;
      const uint64_t v = width == 64 ? vals[i] : (vals[i] & ((1ULL << width) - 1ULL));
      tmp[wd] |= v << off;
      if (off + width > 64) tmp[wd + 1] |= v >> (64 - off);
    }
    if (nw) memcpy(o.data() + at, tmp.data(), nw * 8);
  }

  void close_block_long(uint64_t first, uint64_t last_pos, uint8_t width) {
    // every argument position from `first` on, at most 4096 of them, not beyond last_pos
    uint64_t vals[kSuper];
    uint64_t k = 0;
    for (uint64_t j = first; k < kSuper && j <= last_pos && j < n_; ++j)
      if (arg(j)) vals[k++] = j;
    put_int_vector(blocks_, vals, k, kSuper, width);
    is_mini_.push_back(0);
    any_long_ = true;
  }
  void close_block_mini(const uint64_t* samples, uint64_t nsamples, uint64_t pos_diff) {
    uint64_t vals[64];
    for (uint64_t j = 0; j < nsamples; ++j) vals[j] = samples[j] - samples[0];
    put_int_vector(blocks_, vals, nsamples, 64, (uint8_t)(hi(pos_diff) + 1));
    is_mini_.push_back(1);
  }

  void build_slow() {
    uint64_t pos[kSuper];
    uint64_t cnt = 0;
    for (uint64_t i = 0; i < n_; ++i) {
      if (!arg(i)) continue;
      pos[cnt % kSuper] = i;
      ++cnt;
      if (cnt % kSuper == 0 || cnt == arg_cnt_) {
        const uint64_t last = (cnt - 1) % kSuper;
        super_.push_back(pos[0]);
        const uint64_t pos_diff = pos[last] - pos[0];
        if (pos_diff > logn4_) {
          put_int_vector(blocks_, pos, last + 1, kSuper, (uint8_t)(hi(pos[last]) + 1));
          is_mini_.push_back(0);
          any_long_ = true;
        } else {
          uint64_t samples[64], ns = 0;
          for (uint64_t j = 0; j <= last; j += 64) samples[ns++] = pos[j];
          close_block_mini(samples, ns, pos_diff);
        }
      }
    }
  }

  void build_fast() {
    uint64_t samples[64];          // positions of arguments 1, 65, 129, ... of the open block
    uint64_t last_k64 = 1, last_k64_sum = 1, cnt_old = 0, cnt_new = 0;
    const uint64_t nwords = (n_ + 63) >> 6;
    for (uint64_t wi = 0; wi < nwords; ++wi) {
      const uint64_t w = word(wi);
      cnt_new += (uint64_t)__builtin_popcountll(w);
      if (cnt_new >= last_k64_sum) {
        samples[(last_k64 - 1) >> 6] = wi * 64 + sel(w, (uint32_t)(last_k64_sum - cnt_old));
        last_k64 += 64;
        last_k64_sum += 64;
        if (last_k64 == kSuper + 1 && super_.size() >= nsuper_) {
          last_k64 = 1;   // (a block completed by padding bits only: SDSL writes past its arrays here)
        } else if (last_k64 == kSuper + 1) {
          super_.push_back(samples[0]);
          uint64_t last_pos = samples[63];
          for (uint64_t ii = samples[63] + 1, j = kSuper - 64; ii < n_ && j < kSuper; ++ii)
            if (arg(ii)) { last_pos = ii; ++j; }
          const uint64_t pos_diff = last_pos - samples[0];
          if (pos_diff > logn4_) close_block_long(samples[0], last_pos, (uint8_t)(hi(last_pos) + 1));
          else close_block_mini(samples, 64, pos_diff);
          last_k64 = 1;
        }
      }
      cnt_old = cnt_new;
    }
    if (last_k64 > 1 && is_mini_.size() < nsuper_) {
      // the final, partial block: always long, width of the whole vector, superblock entry left 0
      super_.push_back(0);
      close_block_long(samples[0], n_ ? n_ - 1 : 0, (uint8_t)(hi(n_ - 1) + 1));
    }
  }

  void build() {
    uint64_t ones = 0;
    const uint64_t nwords = (n_ + 63) >> 6;
    for (uint64_t i = 0; i < nwords; ++i) ones += (uint64_t)__builtin_popcountll(bv_[i]);
    arg_cnt_ = ones_ ? ones : n_ - ones;
    const uint64_t logn = hi(((n_ + 63) >> 6) << 6) + 1;
    logn4_ = logn * logn * logn * logn;
    nsuper_ = (arg_cnt_ + kSuper - 1) / kSuper;
    put_u64(out_, arg_cnt_);
    if (!arg_cnt_) return;
    if (n_ < 100000) build_slow(); else build_fast();
    while (super_.size() < nsuper_) super_.push_back(0);
    put_int_vector(out_, super_.data(), nsuper_, nsuper_, (uint8_t)logn);
    if (any_long_) {
      std::vector<uint64_t> bits((nsuper_ + 63) >> 6, 0);
      for (uint64_t i = 0; i < is_mini_.size() && i < nsuper_; ++i) if (is_mini_[i]) bits[i >> 6] |= 1ULL << (i & 63);
      put_u64(out_, (1ULL << 56) | nsuper_);
      size_t at = out_.size();
      out_.resize(at + bits.size() * 8);
      memcpy(out_.data() + at, bits.data(), bits.size() * 8);
    } else {
      put_u64(out_, 1ULL << 56);   // empty bit_vector
    }
    out_.insert(out_.end(), blocks_.begin(), blocks_.end());
  }

  const uint64_t* bv_;
  uint64_t n_;
  bool ones_;
  uint64_t arg_cnt_ = 0, logn4_ = 0, nsuper_ = 0;
  bool any_long_ = false;
  std::vector<uint64_t> super_;
  std::vector<uint8_t> is_mini_;
  std::vector<uint8_t> blocks_, out_;
  std::vector<uint64_t> scratch_;
};

}  // namespace dg
