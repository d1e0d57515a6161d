// host_widen_check.cpp — the in-place u32 -> u64 widening of downloaded index arrays (csrc/host_widen.hpp):
// every length around the wave / tail thresholds, several pool sizes, several jobs queued at once.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../formoniq_b200/csrc/host_widen.hpp"

static uint32_t value_of(size_t i, size_t salt) { return uint32_t((i * 2654435761u) ^ (salt * 40503u) ^ 0x9E3779B9u); }

int main() {
  const size_t sizes[] = {0, 1, 2, 3, 7, 1000, 65535, 65536, 65537, 65538, 131071, 131073, 262145, 1000003, 4194304 + 5};
  size_t checked = 0;
  for (int threads : {1, 2, 3, 7, 16}) {
    fq::HostWidener pool(threads);
    std::vector<std::vector<uint64_t>> bufs;
    for (size_t n : sizes) bufs.emplace_back(n ? n : 1, 0xDEADBEEFDEADBEEFull);
    for (size_t k = 0; k < bufs.size(); ++k) {
      const size_t n = sizes[k];
      uint32_t* src = reinterpret_cast<uint32_t*>(bufs[k].data()) + n;
      for (size_t i = 0; i < n; ++i) src[i] = value_of(i, k);
      if (n) pool.enqueue(bufs[k].data(), n);  // all jobs queued back to back
    }
    pool.wait_idle();
    for (size_t k = 0; k < bufs.size(); ++k)
      for (size_t i = 0; i < sizes[k]; ++i) {
        if (bufs[k][i] != uint64_t(value_of(i, k))) {
          std::printf("FAILED threads=%d n=%zu i=%zu got %llx\n", threads, sizes[k], i, (unsigned long long)bufs[k][i]);
          return 1;
        }
        ++checked;
      }
  }
  // the serial definition alone
  for (size_t n : sizes) {
    std::vector<uint64_t> b(n ? n : 1);
    uint32_t* src = reinterpret_cast<uint32_t*>(b.data()) + n;
    for (size_t i = 0; i < n; ++i) src[i] = value_of(i, 99);
    fq::HostWidener::widen_serial(b.data(), n);
    for (size_t i = 0; i < n; ++i)
      if (b[i] != uint64_t(value_of(i, 99))) {
        std::printf("FAILED serial n=%zu i=%zu\n", n, i);
        return 1;
      }
  }
  std::printf("OK %zu\n", checked);
  return 0;
}
