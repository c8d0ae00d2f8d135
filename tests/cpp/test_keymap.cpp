// Unit test of the library's flat key table (kimera-rpgo_b200/csrc/keymap.h) against std::unordered_map: CPU only.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <unordered_map>

#include "keymap.h"

#define REQUIRE(c)                                                          \
  do {                                                                      \
    if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
  } while (0)

int main() {
  std::mt19937_64 rng(12345);
  KeyMap m;
  std::unordered_map<uint64_t, int32_t> ref;
  REQUIRE(m.size() == 0 && m.find(7) == m.end() && m.count(7) == 0 && m.begin() == m.end());
  // gtsam::Symbol-shaped keys: a handful of prefixes, mostly consecutive indices, some random ones
  const unsigned char chrs[] = {'a', 'b', 'c', 'l', 0};
  for (int round = 0; round < 3; ++round) {
    if (round == 1) m.reserve(m.size() + 200000);  // both growth paths: explicit reserve and doubling
    for (int i = 0; i < 150000; ++i) {
      const uint64_t c = chrs[rng() % 5];
      const uint64_t idx = (rng() % 4 == 0) ? (rng() & 0x00ffffffffffffffull) : (uint64_t)(i + round * 150000);
      const uint64_t k = (c << 56) | idx;
      const int32_t v = (int32_t)(rng() & 0x7fffffff);
      if (rng() % 3 == 0) {
        auto it = m.find(k);
        auto jt = ref.find(k);
        REQUIRE((it == m.end()) == (jt == ref.end()));
        if (jt != ref.end()) REQUIRE(it->first == k && it->second == jt->second);
        REQUIRE(m.count(k) == ref.count(k));
      } else {
        m[k] = v;
        ref[k] = v;
      }
    }
    REQUIRE(m.size() == ref.size());
  }
  // key 0 (chr 0, index 0) is an ordinary key
  m[0] = 42;
  ref[0] = 42;
  REQUIRE(m.find(0) != m.end() && m.find(0)->second == 42);
  // operator[] on a missing key value-initialises, like std::unordered_map
  REQUIRE(m[0x6100000000abcdefull] == 0);
  ref[0x6100000000abcdefull] = 0;
  // iteration visits every entry exactly once
  size_t seen = 0;
  for (const auto& kv : m) {
    auto jt = ref.find(kv.first);
    REQUIRE(jt != ref.end() && jt->second == kv.second);
    ++seen;
  }
  REQUIRE(seen == ref.size());
  for (const auto& kv : ref) REQUIRE(m.find(kv.first) != m.end() && m.find(kv.first)->second == kv.second);
  m.clear();
  REQUIRE(m.size() == 0 && m.find(0) == m.end() && m.begin() == m.end());
  m[5] = 1;
  REQUIRE(m.size() == 1 && m.find(5)->second == 1);
  std::printf("KEYMAP PASS (%zu keys)\n", ref.size());
  return 0;
}
