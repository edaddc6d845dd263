// host_pack.cu — host side of the batch staging: copy many small host arrays back to back into one (pinned)
// arena.  Replaces torch.cat over ~7,000 tiny CPU tensors per 128-scene batch (the reference's dict-of-lists
// collate format, data.py:555-561), which cost ~15 ms of host time per batch; the reference itself issues one
// cudaMemcpyAsync per tensor (utils.py:74-85).  No device code here.
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"

extern "C" int lgcn_pack_host(const void* const* h_srcs, const int64_t* h_nbytes, int64_t n, void* h_dst,
                              int64_t dst_bytes, int n_threads) {
  LGCN_CHECK_ARG(n >= 0 && (n == 0 || (h_srcs && h_nbytes && h_dst)), "pack_host: NULL argument");
  std::vector<int64_t> off((size_t)n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    LGCN_CHECK_ARG(h_nbytes[i] >= 0, "pack_host: negative size at %lld", (long long)i);
    off[i + 1] = off[i] + h_nbytes[i];
  }
  LGCN_CHECK_ARG(off[n] <= dst_bytes, "pack_host: %lld bytes do not fit the %lld-byte arena", (long long)off[n],
                 (long long)dst_bytes);
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i)
      if (h_nbytes[i]) memcpy((char*)h_dst + off[i], h_srcs[i], (size_t)h_nbytes[i]);
  };
  if (n_threads <= 1 || off[n] < (1 << 20)) {
    work(0, n);
    return 0;
  }
  // split by bytes, not by count
  std::vector<std::thread> pool;
  int64_t begin = 0;
  for (int t = 0; t < n_threads; ++t) {
    const int64_t target = off[n] * (t + 1) / n_threads;
    int64_t end = begin;
    while (end < n && off[end + 1] <= target) ++end;
    if (t == n_threads - 1) end = n;
    if (end > begin) pool.emplace_back(work, begin, end);
    begin = end;
  }
  for (auto& th : pool) th.join();
  return 0;
}
