// pred_net.cu — PredNet + AttDest + score sort + world transform (lanegcn.py:575-631, 713-737, 145-150) as ONE kernel.
//
// Per actor: 6 regression heads (LinearRes 128 -> Linear 128x60, + the actor centre), destination attention on the 6
// end points (Linear(2,128)+ReLU -> Linear+GN+ReLU -> Linear(256,128)+GN+ReLU on cat(dist, actor)), the score head
// (LinearRes -> Linear 128x1), a descending sort of the 6 scores with the matching permutation of the trajectories,
// and the rotation / translation into world coordinates of the scene the actor belongs to.  The reference spells
// this as ~60 small cuBLAS / ATen launches plus per-scene python loops (lanegcn.py:609-612, 626-630, 147-150).
//
// Mapping: a CTA owns kGA actors; warp m (0..5) owns MODE m of all of them (its own head weights, then the rows
// (actor, m) of the shared AttDest / cls layers), so the six pipelines need no block-wide barrier until the sort.
// A lane owns 4 output channels x kGA rows: weights are read once per warp as coalesced 512-byte rows of the
// pre-transposed [in][out] matrices (L2 / L1), the input rows are warp-broadcast shared loads; every row's 128
// channels live in one warp, so GroupNorm(1) is two shuffles reductions.  Plain fp32 FMA.
#include <atomic>

#include "common.cuh"

namespace {

constexpr int kModes = 6;
constexpr int kThreads = 2 * 32 * kModes;   // two warps per mode, each owning half of the CTA's actors
constexpr int kPred = 60;                   // 2 * num_preds
constexpr int C = LGCN_C;

// pack layout (floats); every matrix transposed to [in][out]
constexpr int64_t kHead = 2 * (C * C + 2 * C) + C * 64 + 64;          // W1t g1 b1 W2t g2 b2 W3t[128][64] bias3[64]
constexpr int64_t kOffDest = kModes * kHead;                           // Wd0t[2][128] bd0 | Wd2t g b | Wagt_t[256][128] g b
constexpr int64_t kDestFloats = 3 * C + (C * C + 2 * C) + (2 * C * C + 2 * C);
constexpr int64_t kOffCls = kOffDest + kDestFloats;                    // Wc1t g b | Wc2t g b | wc3[128] | bc3 (4)
constexpr int64_t kClsFloats = 2 * (C * C + 2 * C) + C + 4;
constexpr int64_t kPackFloats = kOffCls + kClsFloats;

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// acc[r][0..3] += sum_k in[r][k] * Wt[k][co0..co0+3]   for RT rows of `in` (shared, row stride ld)
template <int RT, int KIN>
__device__ __forceinline__ void rows_fma(const float* __restrict__ in, int ld, const float* __restrict__ Wt, int nout,
                                         int co0, float (&acc)[RT][4]) {
  if (co0 >= nout) return;
  // weight rows come straight from L2 (each warp streams its own matrices): the unroll factor is the number of 4-row
  // batches in flight, and with few rows per warp the loop is a chain of L2 round trips (94 us for 320 actors at 2)
  constexpr int kUnroll = RT <= 4 ? 8 : 4;
#pragma unroll kUnroll
  for (int k = 0; k < KIN; k += 4) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wt + (int64_t)k * nout + co0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wt + (int64_t)(k + 1) * nout + co0));
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(Wt + (int64_t)(k + 2) * nout + co0));
    const float4 w3 = __ldg(reinterpret_cast<const float4*>(Wt + (int64_t)(k + 3) * nout + co0));
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float4 v = lds4(in + r * ld + k);
      acc[r][0] = fmaf(v.w, w3.x, fmaf(v.z, w2.x, fmaf(v.y, w1.x, fmaf(v.x, w0.x, acc[r][0]))));
      acc[r][1] = fmaf(v.w, w3.y, fmaf(v.z, w2.y, fmaf(v.y, w1.y, fmaf(v.x, w0.y, acc[r][1]))));
      acc[r][2] = fmaf(v.w, w3.z, fmaf(v.z, w2.z, fmaf(v.y, w1.z, fmaf(v.x, w0.z, acc[r][2]))));
      acc[r][3] = fmaf(v.w, w3.w, fmaf(v.z, w2.w, fmaf(v.y, w1.w, fmaf(v.x, w0.w, acc[r][3]))));
    }
  }
}

template <int RT>
__device__ __forceinline__ void zero(float (&acc)[RT][4]) {
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
}

// acc <- [relu]( GN(acc) * gamma + beta [+ res] ) in registers, and written to `out` (shared [RT][128]) unless NULL;
// a warp holds whole rows (32 lanes x 4 channels)
template <int RT>
__device__ __forceinline__ void gn_rows(float (&acc)[RT][4], const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const float* __restrict__ res, bool relu, float* __restrict__ out, int lane) {
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane), b = __ldg(reinterpret_cast<const float4*>(beta) + lane);
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    float4 y = warp_gn128(make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]), g, b);
    if (res) {
      const float4 x = lds4(res + r * C + lane * 4);
      y.x += x.x; y.y += x.y; y.z += x.z; y.w += x.w;
    }
    if (relu) y = relu4(y);
    acc[r][0] = y.x; acc[r][1] = y.y; acc[r][2] = y.z; acc[r][3] = y.w;
    if (out) *reinterpret_cast<float4*>(out + r * C + lane * 4) = y;
  }
  __syncwarp();
}

template <int RT>
struct Smem {   // floats
  static constexpr int GA = 2 * RT;                           // actors per CTA
  static constexpr int offA = 0;                              // actors [GA][128]
  static constexpr int offH1 = offA + GA * C;                 // [modes][GA][128]
  static constexpr int offH2 = offH1 + kModes * GA * C;
  static constexpr int offReg = offH2 + kModes * GA * C;      // [GA][modes][64]
  static constexpr int offScore = offReg + GA * kModes * 64;  // [GA][8]
  static constexpr int offCtr = offScore + GA * 8;            // [GA][2]
  static constexpr int floats = offCtr + GA * 2;
};

template <int RT>
__global__ void __launch_bounds__(kThreads, 1)
k_pred_net(const float* __restrict__ actors, const float* __restrict__ ctrs, const int32_t* __restrict__ actor_off, int n_scenes,
           const float* __restrict__ rot, const float* __restrict__ orig, const float* __restrict__ pack,
           float* __restrict__ cls_out /* [A][6] */, float* __restrict__ reg_out /* [A][6][30][2] */, int64_t a_cap,
           const int32_t* __restrict__ a_dev) {
  using S = Smem<RT>;
  constexpr int GA = S::GA;
  extern __shared__ __align__(16) float sm[];
  const int64_t A = lgcn_devn(a_dev, a_cap);
  const int64_t a0 = (int64_t)blockIdx.x * GA;
  if (a0 >= A) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, co0 = lane * 4;
  const int m = warp % kModes, row0 = (warp / kModes) * RT;   // this warp: mode m, actors row0 .. row0 + RT - 1 of the CTA
  float *sA = sm + S::offA + row0 * C, *H1 = sm + S::offH1 + (m * GA + row0) * C, *H2 = sm + S::offH2 + (m * GA + row0) * C,
        *sReg = sm + S::offReg, *sScore = sm + S::offScore, *sCtr = sm + S::offCtr;
  for (int i = threadIdx.x; i < GA * 32; i += kThreads) {
    const int r = i >> 5, c = (i & 31) * 4;
    reinterpret_cast<float4*>(sm + S::offA)[i] = a0 + r < A ? __ldg(reinterpret_cast<const float4*>(actors + (a0 + r) * C + c))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < GA * 2; i += kThreads) sCtr[i] = a0 + (i >> 1) < A ? ctrs[(a0 + (i >> 1)) * 2 + (i & 1)] : 0.f;
  __syncthreads();

  float acc[RT][4];
  // ---- regression head m: LinearRes (layers.py:193-238) then Linear(128, 60) + bias, + actor centre   lanegcn.py:602-612
  {
    const float* h = pack + m * kHead;
    const float *W1 = h, *g1 = W1 + C * C, *b1 = g1 + C, *W2 = b1 + C, *g2 = W2 + C * C, *b2 = g2 + C, *W3 = b2 + C,
                *bias3 = W3 + C * 64;
    zero<RT>(acc);
    rows_fma<RT, C>(sA, C, W1, C, co0, acc);
    gn_rows<RT>(acc, g1, b1, nullptr, true, H1, lane);
    zero<RT>(acc);
    rows_fma<RT, C>(H1, C, W2, C, co0, acc);
    gn_rows<RT>(acc, g2, b2, sA, true, H2, lane);
    zero<RT>(acc);
    rows_fma<RT, C>(H2, C, W3, 64, co0, acc);
    if (co0 < 64) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias3 + co0));
#pragma unroll
      for (int r = 0; r < RT; ++r) {   // columns alternate x, y: reg[..., t, 0] += ctr.x, reg[..., t, 1] += ctr.y
        const float cx = sCtr[2 * (row0 + r)], cy = sCtr[2 * (row0 + r) + 1];
        *reinterpret_cast<float4*>(sReg + ((row0 + r) * kModes + m) * 64 + co0) =
            make_float4(acc[r][0] + bb.x + cx, acc[r][1] + bb.y + cy, acc[r][2] + bb.z + cx, acc[r][3] + bb.w + cy);
      }
    }
    __syncwarp();
  }
  // ---- AttDest on (actor, mode m) rows (lanegcn.py:726-737): dist = relu(L2(ctr - dest) + b) -> Linear+GN+ReLU;
  //      feats = relu(GN(Linear_256(cat(dist, actor))))
  {
    const float* d = pack + kOffDest;
    const float *Wd0 = d, *bd0 = Wd0 + 2 * C, *Wd2 = bd0 + C, *gd = Wd2 + C * C, *bd = gd + C, *Wa = bd + C,
                *ga = Wa + 2 * C * C, *ba = ga + C;
    const float4 wx = __ldg(reinterpret_cast<const float4*>(Wd0 + co0)), wy = __ldg(reinterpret_cast<const float4*>(Wd0 + C + co0)),
                 b0 = __ldg(reinterpret_cast<const float4*>(bd0 + co0));
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float* pr = sReg + ((row0 + r) * kModes + m) * 64;
      const float dx = sCtr[2 * (row0 + r)] - pr[kPred - 2], dy = sCtr[2 * (row0 + r) + 1] - pr[kPred - 1];
      // same association as addmm(bias, x, W^T): (x0*w0 + x1*w1) + b
      *reinterpret_cast<float4*>(H1 + r * C + co0) =
          relu4(make_float4(fmaf(dy, wy.x, dx * wx.x) + b0.x, fmaf(dy, wy.y, dx * wx.y) + b0.y, fmaf(dy, wy.z, dx * wx.z) + b0.z,
                            fmaf(dy, wy.w, dx * wx.w) + b0.w));
    }
    __syncwarp();
    zero<RT>(acc);
    rows_fma<RT, C>(H1, C, Wd2, C, co0, acc);
    gn_rows<RT>(acc, gd, bd, nullptr, true, H2, lane);
    zero<RT>(acc);
    rows_fma<RT, C>(H2, C, Wa, C, co0, acc);             // columns 0..127 of the 256-wide input: dist
    rows_fma<RT, C>(sA, C, Wa + C * C, C, co0, acc);     // columns 128..255: the actor feature
    gn_rows<RT>(acc, ga, ba, nullptr, true, H1, lane);   // feats
  }
  // ---- score head: LinearRes then Linear(128, 1) + bias                                          lanegcn.py:619
  {
    const float* c = pack + kOffCls;
    const float *W1 = c, *g1 = W1 + C * C, *b1 = g1 + C, *W2 = b1 + C, *g2 = W2 + C * C, *b2 = g2 + C, *w3 = b2 + C, *b3 = w3 + C;
    zero<RT>(acc);
    rows_fma<RT, C>(H1, C, W1, C, co0, acc);
    gn_rows<RT>(acc, g1, b1, nullptr, true, H2, lane);
    zero<RT>(acc);
    rows_fma<RT, C>(H2, C, W2, C, co0, acc);
    gn_rows<RT>(acc, g2, b2, H1, true, nullptr, lane);   // stays in registers: only the score is needed
    const float4 w = __ldg(reinterpret_cast<const float4*>(w3 + co0));
    const float bias = __ldg(b3);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float s = warp_sum(fmaf(acc[r][3], w.w, fmaf(acc[r][2], w.z, fmaf(acc[r][1], w.y, acc[r][0] * w.x))));
      if (lane == 0) sScore[(row0 + r) * 8 + m] = s + bias;
    }
  }
  __syncthreads();
  // ---- descending sort of the 6 scores per actor, trajectories permuted alike (lanegcn.py:621-631), then the world
  //      transform reg . rot[scene] + orig[scene] (lanegcn.py:145-150)
  for (int r = warp; r < GA; r += 2 * kModes) {
    const int64_t a = a0 + r;
    if (a >= A) continue;
    float s[kModes];
    int rank[kModes];
#pragma unroll
    for (int i = 0; i < kModes; ++i) s[i] = sScore[r * 8 + i];
#pragma unroll
    for (int i = 0; i < kModes; ++i) {   // rank = number of scores that come before s[i] (ties: lower index first)
      int k = 0;
#pragma unroll
      for (int j = 0; j < kModes; ++j) k += (s[j] > s[i]) || (s[j] == s[i] && j < i);
      rank[i] = k;
    }
    float4 R = make_float4(1.f, 0.f, 0.f, 1.f);   // r00 r01 r10 r11
    float2 o = make_float2(0.f, 0.f);
    if (rot) {   // rot == NULL: scene coordinates (PredNet.forward on its own, lanegcn.py:602-631)
      const int b = scene_of(actor_off, n_scenes, (int32_t)a);
      R = make_float4(__ldg(rot + 4 * b), __ldg(rot + 4 * b + 1), __ldg(rot + 4 * b + 2), __ldg(rot + 4 * b + 3));
      o = make_float2(__ldg(orig + 2 * b), __ldg(orig + 2 * b + 1));
    }
#pragma unroll
    for (int i = 0; i < kModes; ++i) {
      if (lane == 0) cls_out[a * kModes + rank[i]] = s[i];
      if (lane < 30) {
        const float x = sReg[(r * kModes + i) * 64 + 2 * lane], y = sReg[(r * kModes + i) * 64 + 2 * lane + 1];
        reinterpret_cast<float2*>(reg_out)[(a * kModes + rank[i]) * 30 + lane] =
            rot ? make_float2(__fadd_rn(__fadd_rn(__fmul_rn(x, R.x), __fmul_rn(y, R.z)), o.x),
                              __fadd_rn(__fadd_rn(__fmul_rn(x, R.y), __fmul_rn(y, R.w)), o.y))
                : make_float2(x, y);
      }
    }
  }
}

// [out][in] (nn.Linear) -> [in][out_pad] (zero padded columns)
__global__ void k_transpose_linear(const float* __restrict__ w, float* __restrict__ dst, int n_out, int n_in, int out_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_in * out_pad) return;
  const int o = i % out_pad, k = i / out_pad;
  dst[i] = o < n_out ? w[(int64_t)o * n_in + k] : 0.f;
}

std::atomic<int> g_attr[64];

int tr(const float* w, float* dst, int n_out, int n_in, int out_pad, cudaStream_t st) {
  k_transpose_linear<<<lgcn_cdiv((int64_t)n_in * out_pad, 256), 256, 0, st>>>(w, dst, n_out, n_in, out_pad);
  LGCN_LAUNCH_OK();
  return 0;
}
int cp(const float* src, float* dst, int n, cudaStream_t st) {
  LGCN_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace

extern "C" int64_t lgcn_pred_net_wpack_floats(void) { return kPackFloats; }

// h_params: device pointers in state_dict order of PredNet (lanegcn.py:580-600):
//   per mode m (6x): pred.m.0.linear1.weight, pred.m.0.linear2.weight, pred.m.0.norm1.weight, .bias, pred.m.0.norm2.weight,
//                    .bias, pred.m.1.weight [60,128], pred.m.1.bias [60]                                  (8 each)
//   att_dest: dist.0.weight [128,2], dist.0.bias, dist.2.linear.weight, dist.2.norm.weight, .bias, agt.linear.weight
//             [128,256], agt.norm.weight, .bias                                                           (8)
//   cls: cls.0.linear1.weight, cls.0.linear2.weight, cls.0.norm1.weight, .bias, cls.0.norm2.weight, .bias, cls.1.weight
//        [1,128], cls.1.bias [1]                                                                          (8)
extern "C" int lgcn_pred_net_pack(const float* const* h_params, float* wpack, void* stream) {
  LGCN_CHECK_ARG(h_params && wpack, "pred_net_pack: NULL argument");
  for (int i = 0; i < 8 * kModes + 16; ++i) LGCN_CHECK_ARG(h_params[i], "pred_net_pack: NULL parameter %d", i);
  cudaStream_t st = (cudaStream_t)stream;
  LGCN_CUDA_OK(cudaMemsetAsync(wpack, 0, kPackFloats * 4, st));
  for (int m = 0; m < kModes; ++m) {
    const float* const* p = h_params + 8 * m;
    float* h = wpack + m * kHead;
    float *W1 = h, *g1 = W1 + C * C, *b1 = g1 + C, *W2 = b1 + C, *g2 = W2 + C * C, *b2 = g2 + C, *W3 = b2 + C, *bias3 = W3 + C * 64;
    if (tr(p[0], W1, C, C, C, st) || tr(p[1], W2, C, C, C, st) || cp(p[2], g1, C, st) || cp(p[3], b1, C, st) ||
        cp(p[4], g2, C, st) || cp(p[5], b2, C, st) || tr(p[6], W3, kPred, C, 64, st) || cp(p[7], bias3, kPred, st))
      return -2;
  }
  {
    const float* const* p = h_params + 8 * kModes;
    float* d = wpack + kOffDest;
    float *Wd0 = d, *bd0 = Wd0 + 2 * C, *Wd2 = bd0 + C, *gd = Wd2 + C * C, *bd = gd + C, *Wa = bd + C, *ga = Wa + 2 * C * C,
          *ba = ga + C;
    if (tr(p[0], Wd0, C, 2, C, st) || cp(p[1], bd0, C, st) || tr(p[2], Wd2, C, C, C, st) || cp(p[3], gd, C, st) ||
        cp(p[4], bd, C, st) || tr(p[5], Wa, C, 2 * C, C, st) || cp(p[6], ga, C, st) || cp(p[7], ba, C, st))
      return -2;
  }
  {
    const float* const* p = h_params + 8 * kModes + 8;
    float* c = wpack + kOffCls;
    float *W1 = c, *g1 = W1 + C * C, *b1 = g1 + C, *W2 = b1 + C, *g2 = W2 + C * C, *b2 = g2 + C, *w3 = b2 + C, *b3 = w3 + C;
    if (tr(p[0], W1, C, C, C, st) || tr(p[1], W2, C, C, C, st) || cp(p[2], g1, C, st) || cp(p[3], b1, C, st) ||
        cp(p[4], g2, C, st) || cp(p[5], b2, C, st) || cp(p[6], w3, C, st) || cp(p[7], b3, 1, st))
      return -2;
  }
  return 0;
}

extern "C" int lgcn_pred_net(const float* actors, const float* actor_ctrs, const int32_t* actor_off, int n_scenes,
                             const float* rot, const float* orig, const float* wpack, float* cls, float* reg,
                             int64_t n_actors, const int32_t* n_actors_dev, void* stream) {
  LGCN_CHECK_ARG(n_actors >= 0 && n_scenes >= 1, "pred_net: sizes");
  if (n_actors == 0) return 0;
  LGCN_CHECK_ARG(actors && actor_ctrs && wpack && cls && reg, "pred_net: NULL argument");
  LGCN_CHECK_ARG(!rot || (orig && actor_off), "pred_net: rot without orig / actor_off");
  int dev = 0, sms = 148;
  LGCN_CUDA_OK(cudaGetDevice(&dev));
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  if (dev >= 0 && dev < 64 && g_attr[dev].exchange(1) == 0) {
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_pred_net<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::floats * 4));
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_pred_net<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<4>::floats * 4));
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_pred_net<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<6>::floats * 4));
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_pred_net<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<8>::floats * 4));
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_pred_net<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<10>::floats * 4));
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_pred_net<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<12>::floats * 4));
  }
  // actors per CTA (2 RT): the CTAs run one per SM in waves and a CTA takes about 15 us + 15 us x RT (measured: 45 / 78 /
  // 166 us at RT = 2 / 4 / 10), so pick the group size with the smallest waves x (RT + 1)
  // (2,560 actors on 148 SMs: 20 per CTA = 128 CTAs = one wave; 320 actors: 4 per CTA = 80 CTAs)
  int best = 2;
  int64_t best_cost = -1;
  for (int rt : {2, 4, 6, 8, 10, 12}) {
    const int64_t ctas = (n_actors + 2 * rt - 1) / (2 * rt), cost = ((ctas + sms - 1) / sms) * (rt + 1);
    if (best_cost < 0 || cost < best_cost) {
      best = rt;
      best_cost = cost;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = lgcn_cdiv(n_actors, 2 * best);
#define LGCN_PRED(RT_)                                                                                                  \
  k_pred_net<RT_><<<grid, kThreads, Smem<RT_>::floats * 4, st>>>(actors, actor_ctrs, actor_off, n_scenes, rot, orig, wpack, \
                                                                 cls, reg, n_actors, n_actors_dev)
  switch (best) {
    case 2: LGCN_PRED(2); break;
    case 4: LGCN_PRED(4); break;
    case 6: LGCN_PRED(6); break;
    case 8: LGCN_PRED(8); break;
    case 10: LGCN_PRED(10); break;
    default: LGCN_PRED(12); break;
  }
#undef LGCN_PRED
  LGCN_LAUNCH_OK();
  return 0;
}
