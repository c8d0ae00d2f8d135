/* keymap.h — host-side key -> trajectory-entry table of the C-ABI library (no CUDA in here; unit-tested on the CPU by
 * tests/cpp/test_keymap.cpp). */
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

/* key -> trajectory entry.  GTSAM keys are (chr << 56 | index): a flat open-addressing table (Fibonacci hashing, linear
 * probing) resolves the 150 k lookups of a 50 k-closure batch in well under a millisecond; std::unordered_map needed 4-5 ms
 * (a node allocation per insert, a pointer chase per find).  Interface: the subset of std::unordered_map the library uses. */
class KeyMap {
 public:
  struct Slot { uint64_t first; int32_t second; };
  class iterator {
   public:
    iterator(const Slot* p, const Slot* e) : p_(p), e_(e) { skip(); }
    const Slot& operator*() const { return *p_; }
    const Slot* operator->() const { return p_; }
    iterator& operator++() { ++p_; skip(); return *this; }
    bool operator==(const iterator& o) const { return p_ == o.p_; }
    bool operator!=(const iterator& o) const { return p_ != o.p_; }
   private:
    void skip() { while (p_ != e_ && p_->first == EMPTY) ++p_; }
    const Slot* p_;
    const Slot* e_;
  };
  iterator begin() const { return iterator(slots_.data(), slots_.data() + slots_.size()); }
  iterator end() const { return iterator(slots_.data() + slots_.size(), slots_.data() + slots_.size()); }
  size_t size() const { return n_; }
  void clear() { slots_.clear(); n_ = 0; shift_ = 64; }
  void reserve(size_t n) { if (n * 2 > slots_.size()) rehash(n * 2); }
  iterator find(uint64_t k) const {
    const Slot* s = lookup(k);
    return s ? iterator(s, slots_.data() + slots_.size()) : end();
  }
  size_t count(uint64_t k) const { return lookup(k) ? 1 : 0; }
  int32_t& operator[](uint64_t k) {
    if ((n_ + 1) * 2 > slots_.size()) rehash(std::max<size_t>(64, slots_.size() * 2));
    const size_t mask = slots_.size() - 1;
    size_t i = hash(k);
    /* the all-ones key doubles as the empty marker; it is not a valid gtsam::Symbol (chr 0xff, index 2^56-1) */
    while (slots_[i].first != EMPTY && slots_[i].first != k) i = (i + 1) & mask;
    if (slots_[i].first == EMPTY) { slots_[i].first = k; slots_[i].second = 0; ++n_; }
    return slots_[i].second;
  }

 private:
  static constexpr uint64_t EMPTY = ~0ull;
  size_t hash(uint64_t k) const { return (size_t)((k * 0x9E3779B97F4A7C15ull) >> shift_); }
  const Slot* lookup(uint64_t k) const {
    if (slots_.empty()) return nullptr;
    const size_t mask = slots_.size() - 1;
    size_t i = hash(k);
    while (slots_[i].first != EMPTY) {
      if (slots_[i].first == k) return &slots_[i];
      i = (i + 1) & mask;
    }
    return nullptr;
  }
  void rehash(size_t want) {
    size_t cap = 64;
    int bits = 6;
    while (cap < want) { cap <<= 1; ++bits; }
    std::vector<Slot> old;
    old.swap(slots_);
    slots_.assign(cap, Slot{EMPTY, 0});
    shift_ = 64 - bits;
    n_ = 0;
    for (const Slot& s : old)
      if (s.first != EMPTY) (*this)[s.first] = s.second;
  }
  std::vector<Slot> slots_;
  size_t n_ = 0;
  int shift_ = 64;
};

