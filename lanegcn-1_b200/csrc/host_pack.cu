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

// ------------------------------------------------------------------ packed scenes -> staging arenas
// A PACKED SCENE is one contiguous host blob made once per sample from the preprocessed-pickle schema
// (preprocess_data.py:78-95, data.py:564-575 — lanegcn.pack_scene):
//   int64 header[8 + n_kv]: magic 'LGCNSCN1', n_nodes, n_actors, n_scales, idx_bytes, n_index, 0, 0, seg_len[n_kv]
//                           (n_kv = 2 * (2 n_scales + 2): entries of pre0.u, pre0.v, suc0.u, ..., right.v)
//   float  arena: ctrs[2N] feats[2N] turn[2N] control[N] intersect[N] actor_feats[60A] actor_ctrs[2A] rot[4] orig[2]
//   index  arena: the n_kv segments back to back (idx_bytes per entry), padded to 8 bytes
// lgcn_stage_scenes assembles a batch of them into the four staging buffers of the one-call forward in the bucket's
// CAPACITY layout (so each buffer goes to the device with ONE copy): what collate_fn + utils.gpu + the offset / cat
// loops of graph_gather do per tensor (data.py:555-561, utils.py:74-96, lanegcn.py:171-209), as B x 37 memcpys in C.
static const int64_t kSceneMagic = 0x314e43534e43474cLL;   // "LGCNSCN1"

extern "C" int lgcn_stage_scenes(const void* const* h_blobs, int n_scenes, int64_t cap_nodes, int64_t cap_actors,
                                 int64_t cap_index, int cap_scenes, int n_scales, int idx_bytes, float* h_fl, void* h_idx,
                                 int64_t* h_t64, int32_t* h_t32, int n_threads) {
  LGCN_CHECK_ARG(h_blobs && h_fl && h_idx && h_t64 && h_t32, "stage_scenes: NULL argument");
  LGCN_CHECK_ARG(n_scenes >= 1 && n_scenes <= cap_scenes, "stage_scenes: %d scenes, capacity %d", n_scenes, cap_scenes);
  const int n_kv = 2 * (2 * n_scales + 2);
  const int64_t B = n_scenes, Bc = cap_scenes;
  std::vector<int64_t> noff((size_t)Bc + 1), aoff((size_t)Bc + 1);
  std::vector<int64_t> seg_start((size_t)n_kv * Bc + 1, 0);
  noff[0] = aoff[0] = 0;
  for (int64_t b = 0; b < B; ++b) {
    const int64_t* h = (const int64_t*)h_blobs[b];
    LGCN_CHECK_ARG(h && h[0] == kSceneMagic, "stage_scenes: scene %lld is not a packed scene", (long long)b);
    LGCN_CHECK_ARG(h[3] == n_scales && h[4] == idx_bytes, "stage_scenes: scene %lld has another index width / scale count",
                   (long long)b);
    noff[b + 1] = noff[b] + h[1];
    aoff[b + 1] = aoff[b] + h[2];
  }
  const int64_t N = noff[B], A = aoff[B];
  LGCN_CHECK_ARG(N <= cap_nodes && A <= cap_actors, "stage_scenes: %lld nodes / %lld actors exceed the capacities",
                 (long long)N, (long long)A);
  for (int64_t b = B + 1; b <= Bc; ++b) {
    noff[b] = N;
    aoff[b] = A;
  }
  // segment table: (kv-major, scene-slot-minor); slots of absent scenes are empty
  int64_t run = 0;
  for (int kv = 0; kv < n_kv; ++kv)
    for (int64_t b = 0; b < Bc; ++b) {
      seg_start[(size_t)kv * Bc + b] = run;
      if (b < B) run += ((const int64_t*)h_blobs[b])[8 + kv];
    }
  seg_start[(size_t)n_kv * Bc] = run;
  LGCN_CHECK_ARG(run <= cap_index, "stage_scenes: %lld index entries exceed the capacity %lld", (long long)run, (long long)cap_index);
  const int64_t n_seg = (int64_t)n_kv * Bc;
  memcpy(h_t64, seg_start.data(), (size_t)(n_seg + 1) * 8);
  for (int kv = 0; kv < n_kv; ++kv)
    for (int64_t b = 0; b < Bc; ++b) h_t64[n_seg + 1 + (int64_t)kv * Bc + b] = noff[b];
  for (int64_t b = 0; b <= Bc; ++b) {
    h_t32[b] = (int32_t)noff[b];
    h_t32[Bc + 1 + b] = (int32_t)aoff[b];
  }
  h_t32[2 * Bc + 2] = (int32_t)N;
  h_t32[2 * Bc + 3] = (int32_t)A;
  h_t32[2 * Bc + 4] = h_t32[2 * Bc + 5] = 0;
  // float regions at capacity offsets: ctrs | feats | turn | control | intersect | actor feats | actor ctrs | rot | orig
  const int64_t per_node[5] = {2, 2, 2, 1, 1};
  int64_t reg_off[9];
  reg_off[0] = 0;
  for (int r = 0; r < 5; ++r) reg_off[r + 1] = reg_off[r] + per_node[r] * cap_nodes;
  reg_off[6] = reg_off[5] + 60 * cap_actors;
  reg_off[7] = reg_off[6] + 2 * cap_actors;
  reg_off[8] = reg_off[7] + 4 * Bc;
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t b = lo; b < hi; ++b) {
      const int64_t* h = (const int64_t*)h_blobs[b];
      const int64_t n = h[1], a = h[2];
      const float* src = (const float*)(h + 8 + n_kv);
      for (int r = 0; r < 5; ++r) {
        memcpy(h_fl + reg_off[r] + per_node[r] * noff[b], src, (size_t)(per_node[r] * n) * 4);
        src += per_node[r] * n;
      }
      memcpy(h_fl + reg_off[5] + 60 * aoff[b], src, (size_t)(60 * a) * 4);
      src += 60 * a;
      memcpy(h_fl + reg_off[6] + 2 * aoff[b], src, (size_t)(2 * a) * 4);
      src += 2 * a;
      memcpy(h_fl + reg_off[7] + 4 * b, src, 16);
      src += 4;
      memcpy(h_fl + reg_off[8] + 2 * b, src, 8);
      src += 2;
      const char* isrc = (const char*)src;
      for (int kv = 0; kv < n_kv; ++kv) {
        const int64_t len = h[8 + kv];
        memcpy((char*)h_idx + seg_start[(size_t)kv * Bc + b] * idx_bytes, isrc, (size_t)(len * idx_bytes));
        isrc += len * idx_bytes;
      }
    }
  };
  // a thread costs ~30 us to start: one per ~2 MB of payload
  const int64_t payload = (8 * N + 62 * A + 6 * B) * 4 + run * idx_bytes;
  int nt = (int)(payload >> 21);
  nt = nt < n_threads ? nt : n_threads;
  nt = nt < B ? nt : (int)B;
  if (nt <= 1) {
    work(0, B);
    return 0;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) pool.emplace_back(work, B * t / nt, B * (t + 1) / nt);
  for (auto& th : pool) th.join();
  return 0;
}
