// actor_net.cu — ActorNet (lanegcn.py:212-263; layers.py:40-62 Conv1d, :142-190 Res1d) as ONE kernel.
//
// The reference runs ~20 Conv1d + GroupNorm(1) + ReLU layers over [A, C, L] (A actors, L = 20/10/5 steps,
// C = 3/32/64/128): on a GPU that is ~110 tiny cuDNN / ATen launches whose activations bounce through HBM.  Here a CTA
// takes kActors actors through the WHOLE network with every activation resident in shared memory (channels-last:
// [actor][step + 2 zero pad rows][channel]); only the 3x20 input and the 128-float result touch HBM.  Weights stream
// from L2 through a shared-memory stage (re-packed once per weight version as [tap][c_in][c_out], so a warp reads 512
// contiguous bytes).  Arithmetic is plain fp32 FMA (the reference's convs are true fp32: cudnn.allow_tf32 = False):
// 8.6 MFLOP per actor, compute-bound on the FP32 pipe; GroupNorm(1) normalises over all (C, L) of an actor, two-pass.
//
// Thread mapping of a conv (a [rows = actors x L_out] x [K = taps x C_in] x [C_out] product): a lane owns CPL output
// channels (4, 2 or 1: distinct, coalesced weight reads; chosen per layer so that no thread is left without outputs) x RT
// consecutive steps of one actor (input rows are warp-broadcast 128-bit shared loads): CPL x RT x 4 FMAs per 4 + RT
// shared loads.
#include <atomic>

#include "common.cuh"

namespace {

constexpr int kActors = 2;      // actors per CTA; 112 KB of shared memory per CTA -> two CTAs per SM, so one CTA's
constexpr int kThreads = 256;   // barriers / GroupNorm passes overlap the other's FMA phases; 16 warps per SM (with 128
                                // threads the schedulers held two warps each and issued 0.42 instructions per cycle)
constexpr int kWChunk = 8;      // input channels per staged weight chunk (x 3 taps x 128 outputs x 4 B = 12 KB), two stages

struct ConvW {
  const float* w;       // [K][CinP][Cout]
  const float* gamma;   // [Cout]
  const float* beta;    // [Cout]
};

// float offsets of the 20 layers inside the pack (lgcn_actor_net_wpack_floats documents the order)
struct Spec { int cin, cout, k; };
__host__ __device__ constexpr int cin_padded(int c) { return c < 4 ? 4 : c; }
__host__ __device__ constexpr Spec spec(int i) {
  constexpr Spec t[20] = {
      {3, 32, 3}, {32, 32, 3}, {3, 32, 1}, {32, 32, 3}, {32, 32, 3},            // groups.0
      {32, 64, 3}, {64, 64, 3}, {32, 64, 1}, {64, 64, 3}, {64, 64, 3},          // groups.1
      {64, 128, 3}, {128, 128, 3}, {64, 128, 1}, {128, 128, 3}, {128, 128, 3},  // groups.2
      {32, 128, 3}, {64, 128, 3}, {128, 128, 3},                                // lateral.0/1/2
      {128, 128, 3}, {128, 128, 3}};                                            // output
  return t[i];
}
__host__ __device__ constexpr int64_t spec_floats(int i) {
  return (int64_t)spec(i).k * cin_padded(spec(i).cin) * spec(i).cout + 2 * spec(i).cout;
}
__host__ __device__ constexpr int64_t spec_offset(int i) {
  int64_t o = 0;
  for (int j = 0; j < i; ++j) o += spec_floats(j);
  return o;
}
template <int I>
__device__ __forceinline__ ConvW layer(const float* pack) {
  constexpr int64_t off = spec_offset(I), nw = (int64_t)spec(I).k * cin_padded(spec(I).cin) * spec(I).cout;
  constexpr int cout = spec(I).cout;
  ConvW c;
  c.w = pack + off;
  c.gamma = c.w + nw;
  c.beta = c.gamma + cout;
  return c;
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Raw convolution: in [G][LIN + 2][CIN] -> out rows 1..LOUT of [G][LOUT + 2][COUT]   (LIN = LOUT * STRIDE; K = 3: pad 1,
// K = 1: pad 0).  All threads of the CTA call it; ends with a __syncthreads().
template <int CIN, int COUT, int K, int STRIDE, int LOUT, int RT, int CPL = 4>
__device__ __forceinline__ void conv(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ Wt,
                                     float* __restrict__ wst) {
  constexpr int LIN = LOUT * STRIDE;
  constexpr int LPG = COUT / CPL;            // lanes per row group (a lane owns CPL output channels)
  constexpr int GROUPS = kThreads / LPG;     // row groups in the CTA
  constexpr int TILES_PER_ACTOR = LOUT / RT;
  constexpr int N_TILES = kActors * TILES_PER_ACTOR;
  constexpr int NR = K == 3 ? (RT - 1) * STRIDE + 3 : RT;   // input rows a tile touches
  constexpr int CH = CIN < kWChunk ? CIN : kWChunk;
  static_assert(LOUT % RT == 0 && COUT % CPL == 0 && kThreads % LPG == 0 && CIN % 4 == 0 && CIN % CH == 0, "shape");
  static_assert(CPL == 1 || CPL == 2 || CPL == 4, "channels per lane");
  const int co0 = (threadIdx.x % LPG) * CPL;
  const int group = threadIdx.x / LPG;

  for (int tile0 = 0; tile0 < N_TILES; tile0 += GROUPS) {
    const int tile = tile0 + group;
    const bool live = tile < N_TILES;
    const int g = live ? tile / TILES_PER_ACTOR : 0, l0 = live ? (tile % TILES_PER_ACTOR) * RT : 0;
    // first input row (padded coordinates) of the tile: K = 3 -> l0 * S + kk, K = 1 -> l0 * S + 1
    const float* xin = in + ((int64_t)g * (LIN + 2) + l0 * STRIDE + (K == 3 ? 0 : 1)) * CIN;
    float acc[RT][CPL];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc[r][c] = 0.f;
    // weight chunks stream through a two-stage cp.async ring: chunk c+1 is in flight while chunk c is consumed
    auto prefetch = [&](int c0, int stage) {   // wst[stage][kk][ci][co] <- Wt[kk][c0 + ci][co]
      float* dst = wst + stage * (3 * kWChunk * 128);
      for (int i = threadIdx.x; i < K * CH * COUT / 4; i += kThreads) {
        const int co4 = i % (COUT / 4), ci = (i / (COUT / 4)) % CH, kk = i / (COUT / 4 * CH);
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + 4 * i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Wt + ((int64_t)kk * CIN + c0 + ci) * COUT + 4 * co4)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0, 0);
    int stage = 0;
    for (int c0 = 0; c0 < CIN; c0 += CH, stage ^= 1) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();   // chunk c0 has landed for every thread; everybody is done with the other stage (and with `in`'s producer)
      if (c0 + CH < CIN) prefetch(c0 + CH, stage ^ 1);
      const float* wcur = wst + stage * (3 * kWChunk * 128);
      if (live) {
#pragma unroll 1
        for (int ci = 0; ci < CH; ci += 4) {
          float4 x[NR];
#pragma unroll
          for (int r = 0; r < NR; ++r) x[r] = lds4(xin + (K == 3 ? r : r * STRIDE) * CIN + c0 + ci);
#pragma unroll
          for (int kk = 0; kk < K; ++kk) {
            const float* wp = wcur + ((int64_t)kk * CH + ci) * COUT + co0;
            float w[4][CPL];   // [input channel ci + q][this lane's output channel]
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if constexpr (CPL == 4) {
                const float4 t = lds4(wp + q * COUT);
                w[q][0] = t.x; w[q][1] = t.y; w[q][2] = t.z; w[q][3] = t.w;
              } else if constexpr (CPL == 2) {
                const float2 t = *reinterpret_cast<const float2*>(wp + q * COUT);
                w[q][0] = t.x; w[q][1] = t.y;
              } else {
                w[q][0] = wp[q * COUT];
              }
            }
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              const float4 v = x[K == 3 ? r * STRIDE + kk : r];
#pragma unroll
              for (int c = 0; c < CPL; ++c)
                acc[r][c] = fmaf(v.w, w[3][c], fmaf(v.z, w[2][c], fmaf(v.y, w[1][c], fmaf(v.x, w[0][c], acc[r][c]))));
            }
          }
        }
      }
    }
    __syncthreads();   // (multi-round tiles only) the last chunk's stage is free before the next round's first prefetch
    if (live) {
      float* o = out + ((int64_t)g * (LOUT + 2) + l0 + 1) * COUT + co0;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        if constexpr (CPL == 4)
          *reinterpret_cast<float4*>(o + (int64_t)r * COUT) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        else if constexpr (CPL == 2)
          *reinterpret_cast<float2*>(o + (int64_t)r * COUT) = make_float2(acc[r][0], acc[r][1]);
        else
          o[(int64_t)r * COUT] = acc[r][0];
      }
    }
  }
  __syncthreads();
}

enum { GN_RELU = 1, GN_ADD_RES = 2, GN_ACCUM = 4 };
// GroupNorm(1 group over all C x L of an actor; layers.py:52 / :158, eps 1e-5, biased variance) of the raw rows
// 1..L of `src`, then  v = norm * gamma + beta  [+ res]  [relu]  ->  dst (= v, or += v with GN_ACCUM).  dst may be src.
// Also zeroes dst's two pad rows.  Ends with a __syncthreads().
template <int C, int L>
__device__ __forceinline__ void gn_apply(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ res, int flags,
                                         float* __restrict__ stats) {
  constexpr int N = C * L, STR = (L + 2) * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < kActors) {
    const float* p = src + (int64_t)warp * STR + C;
    float s = 0.f;
    for (int i = lane * 4; i < N; i += 128) {
      const float4 v = lds4(p + i);
      s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(s) * (1.0f / N);
    float q = 0.f;
    for (int i = lane * 4; i < N; i += 128) {
      const float4 v = lds4(p + i);
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float var = warp_sum(q) * (1.0f / N);
    if (lane == 0) {
      stats[2 * warp] = mean;
      stats[2 * warp + 1] = 1.0f / sqrtf(var + LGCN_GN_EPS);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kActors * N / 4; i += kThreads) {
    const int g = i / (N / 4), e = (i % (N / 4)) * 4, c = e % C;
    const float mean = stats[2 * g], rstd = stats[2 * g + 1];
    const int64_t at = (int64_t)g * STR + C + e;
    const float4 v = lds4(src + at), ga = __ldg(reinterpret_cast<const float4*>(gamma + c)),
                 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y = make_float4((v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y,
                           (v.z - mean) * rstd * ga.z + be.z, (v.w - mean) * rstd * ga.w + be.w);
    if (flags & GN_ADD_RES) {
      const float4 r = lds4(res + at);
      y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    if (flags & GN_RELU) y = relu4(y);
    if (flags & GN_ACCUM) {
      const float4 d = lds4(dst + at);
      y.x += d.x; y.y += d.y; y.z += d.z; y.w += d.w;
    }
    *reinterpret_cast<float4*>(dst + at) = y;
  }
  for (int i = threadIdx.x; i < kActors * 2 * C; i += kThreads) {   // pad rows 0 and L + 1
    const int g = i / (2 * C), r = (i / C) & 1, c = i % C;
    dst[(int64_t)g * STR + (r ? (L + 1) * C : 0) + c] = 0.f;
  }
  __syncthreads();
}

// y = relu(a + b) on rows 1..L (the tail of a Res1d with a down-sampled shortcut), in place into a
template <int C, int L>
__device__ __forceinline__ void add_relu(float* __restrict__ a, const float* __restrict__ b) {
  constexpr int N = C * L, STR = (L + 2) * C;
  for (int i = threadIdx.x; i < kActors * N / 4; i += kThreads) {
    const int64_t at = (int64_t)(i / (N / 4)) * STR + C + (i % (N / 4)) * 4;
    const float4 x = lds4(a + at), y = lds4(b + at);
    *reinterpret_cast<float4*>(a + at) = relu4(make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w));
  }
  __syncthreads();
}

// F.interpolate(scale_factor=2, mode="linear", align_corners=False) along the steps (lanegcn.py:259):
// out[2i] = .25 x[i-1] + .75 x[i], out[2i+1] = .75 x[i] + .25 x[i+1], edges clamped.  src [G][L+2][C] -> dst [G][2L+2][C]
template <int C, int L>
__device__ __forceinline__ void upsample2(const float* __restrict__ src, float* __restrict__ dst) {
  constexpr int SS = (L + 2) * C, DS = (2 * L + 2) * C;
  for (int i = threadIdx.x; i < kActors * 2 * L * C / 4; i += kThreads) {
    const int c = (i % (C / 4)) * 4, j = (i / (C / 4)) % (2 * L), g = i / (C / 4 * 2 * L);
    const int k = j >> 1;
    const int nb = (j & 1) ? (k + 1 < L ? k + 1 : L - 1) : (k > 0 ? k - 1 : 0);
    const float4 x = lds4(src + (int64_t)g * SS + (k + 1) * C + c), y = lds4(src + (int64_t)g * SS + (nb + 1) * C + c);
    // torch: w0 * x[i0] + w1 * x[i1] with (w0, w1) = (0.75, 0.25) on the near / far sample
    *reinterpret_cast<float4*>(dst + (int64_t)g * DS + (j + 1) * C + c) =
        make_float4(0.75f * x.x + 0.25f * y.x, 0.75f * x.y + 0.25f * y.y, 0.75f * x.z + 0.25f * y.z, 0.75f * x.w + 0.25f * y.w);
  }
  for (int i = threadIdx.x; i < kActors * 2 * C; i += kThreads) {
    const int g = i / (2 * C), r = (i / C) & 1, c = i % C;
    dst[(int64_t)g * DS + (r ? (2 * L + 1) * C : 0) + c] = 0.f;
  }
  __syncthreads();
}

// shared memory (floats)
constexpr int kBig = kActors * 22 * 128;                          // a [G][20 + 2][128] buffer
constexpr int kOffX0 = 0;                                         // input [G][22][4]
constexpr int kOffP = kOffX0 + kActors * 22 * 4;
constexpr int kOffQ = kOffP + kBig;
constexpr int kOffR = kOffQ + kBig;
constexpr int kOffF0 = kOffR + kBig;                              // [G][22][32]
constexpr int kOffF1 = kOffF0 + kActors * 22 * 32;                // [G][12][64]
constexpr int kOffF2 = kOffF1 + kActors * 12 * 64;                // [G][7][128]
constexpr int kOffW = kOffF2 + kActors * 7 * 128;                 // weight stages 2 x [3][kWChunk][128]
constexpr int kOffStat = kOffW + 2 * 3 * kWChunk * 128;
constexpr int kSmemFloats = kOffStat + 2 * kActors + 8;
constexpr int kSmemBytes = kSmemFloats * 4;

// Res1d (layers.py:142-190): x -> relu(GN(conv2(relu(GN(conv1(x))))) + shortcut); h, y: scratch; result in `dst`
template <int CIN, int COUT, int STRIDE, int LOUT, int RT, int I1, int I2, int ID, int CPL = 4>
__device__ __forceinline__ void res1d(const float* x, float* h, float* y, float* d, float* dst, const float* pack, float* wst,
                                      float* stats) {
  const ConvW c1 = layer<I1>(pack), c2 = layer<I2>(pack);
  conv<CIN, COUT, 3, STRIDE, LOUT, RT, CPL>(x, h, c1.w, wst);
  gn_apply<COUT, LOUT>(h, h, c1.gamma, c1.beta, nullptr, GN_RELU, stats);
  conv<COUT, COUT, 3, 1, LOUT, RT, CPL>(h, y, c2.w, wst);
  if constexpr (ID >= 0) {
    const ConvW cd = layer<ID>(pack);
    gn_apply<COUT, LOUT>(y, y, c2.gamma, c2.beta, nullptr, 0, stats);
    conv<CIN, COUT, 1, STRIDE, LOUT, RT, CPL>(x, d, cd.w, wst);
    gn_apply<COUT, LOUT>(d, d, cd.gamma, cd.beta, nullptr, 0, stats);
    add_relu<COUT, LOUT>(y, d);   // the blocks with a shortcut conv leave their result in y (dst == y)
  } else {
    gn_apply<COUT, LOUT>(y, dst, c2.gamma, c2.beta, x, GN_ADD_RES | GN_RELU, stats);
  }
}

// FPN_ONLY: stop after the feature pyramid and write its [A][20][128] rows to `out`; the output Res1d (46 % of the
// network's FLOPs: two 128 -> 128 k=3 convs over all 20 steps) then runs on the tensor core (lgcn_actor_net_tc below).
template <bool FPN_ONLY>
__global__ void __launch_bounds__(kThreads, 2)
k_actor_net(const float* __restrict__ feats /* [A][20][3] */, const float* __restrict__ pack,
            float* __restrict__ out /* [A][128], or [A][20][128] with FPN_ONLY */, int64_t a_cap, const int32_t* __restrict__ a_dev) {
  extern __shared__ __align__(16) float sm[];
  const int64_t A = lgcn_devn(a_dev, a_cap);
  const int64_t a0 = (int64_t)blockIdx.x * kActors;
  if (a0 >= A) return;
  float *X0 = sm + kOffX0, *P = sm + kOffP, *Q = sm + kOffQ, *R = sm + kOffR, *F0 = sm + kOffF0, *F1 = sm + kOffF1,
        *F2 = sm + kOffF2, *wst = sm + kOffW, *stats = sm + kOffStat;
  // input: [A][20][3] (dx, dy, valid per step; lanegcn.py:155-168 transposes it to channels-first for Conv1d — the
  // channels-last layout used here is the untransposed one) -> [G][22][4], zero pads and zero 4th channel
  for (int i = threadIdx.x; i < kActors * 22; i += kThreads) {
    const int g = i / 22, r = i % 22;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 1 && r <= 20 && a0 + g < A) {
      const float* p = feats + ((a0 + g) * 20 + (r - 1)) * 3;
      v = make_float4(p[0], p[1], p[2], 0.f);
    }
    reinterpret_cast<float4*>(X0)[i] = v;
  }
  __syncthreads();
  // Rows per tile and channels per lane (the last two shape arguments) are chosen so that every thread of the CTA owns
  // outputs in every layer: kActors x L x C outputs / kThreads = 5 per thread in groups 0-2 (10 rows x 4 channels would
  // leave most threads without a tile there)
  // groups.0 (L = 20, 32 channels)
  res1d<4, 32, 1, 20, 5, 0, 1, 2, 1>(X0, P, Q, R, Q, pack, wst, stats);      // -> Q
  res1d<32, 32, 1, 20, 5, 3, 4, -1, 1>(Q, P, R, nullptr, F0, pack, wst, stats);   // -> F0
  // groups.1 (L = 10, 64 channels)
  res1d<32, 64, 2, 10, 5, 5, 6, 7, 1>(F0, P, Q, R, Q, pack, wst, stats);
  res1d<64, 64, 1, 10, 5, 8, 9, -1, 1>(Q, P, R, nullptr, F1, pack, wst, stats);
  // groups.2 (L = 5, 128 channels)
  res1d<64, 128, 2, 5, 5, 10, 11, 12, 1>(F1, P, Q, R, Q, pack, wst, stats);
  res1d<128, 128, 1, 5, 5, 13, 14, -1, 1>(Q, P, R, nullptr, F2, pack, wst, stats);
  // FPN: out = lateral2(f2); out = up(out) + lateral1(f1); out = up(out) + lateral0(f0)            lanegcn.py:256-261
  {
    const ConvW l2 = layer<17>(pack), l1 = layer<16>(pack), l0 = layer<15>(pack);
    conv<128, 128, 3, 1, 5, 5, 1>(F2, P, l2.w, wst);
    gn_apply<128, 5>(P, P, l2.gamma, l2.beta, nullptr, 0, stats);
    upsample2<128, 5>(P, Q);
    conv<64, 128, 3, 1, 10, 5, 2>(F1, R, l1.w, wst);
    gn_apply<128, 10>(R, Q, l1.gamma, l1.beta, nullptr, GN_ACCUM, stats);
    upsample2<128, 10>(Q, P);
    conv<32, 128, 3, 1, 20, 10, 2>(F0, R, l0.w, wst);
    gn_apply<128, 20>(R, P, l0.gamma, l0.beta, nullptr, GN_ACCUM, stats);
  }
  if constexpr (FPN_ONLY) {
    for (int i = threadIdx.x; i < kActors * 20 * 32; i += kThreads) {
      const int g = i / (20 * 32), l = (i / 32) % 20, c = (i & 31) * 4;
      if (a0 + g < A)
        *reinterpret_cast<float4*>(out + ((a0 + g) * 20 + l) * LGCN_C + c) = lds4(P + ((int64_t)g * 22 + l + 1) * 128 + c);
    }
    return;
  }
  // output Res1d, last step only                                                                  lanegcn.py:262
  res1d<128, 128, 1, 20, 10, 18, 19, -1, 2>(P, Q, R, nullptr, R, pack, wst, stats);
  for (int i = threadIdx.x; i < kActors * 32; i += kThreads) {
    const int g = i >> 5, c = (i & 31) * 4;
    if (a0 + g < A)
      *reinterpret_cast<float4*>(out + (a0 + g) * LGCN_C + c) = lds4(R + (int64_t)g * 22 * 128 + 20 * 128 + c);
  }
}

// conv.weight [Cout][Cin][K] (torch) -> [K][CinP][Cout]; one thread per output element
__global__ void k_pack_conv(const float* __restrict__ w, float* __restrict__ dst, int cin, int cinp, int cout, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * cinp * cout) return;
  const int co = i % cout, ci = (i / cout) % cinp, kk = i / (cout * cinp);
  dst[i] = ci < cin ? w[((int64_t)co * cin + ci) * k + kk] : 0.f;
}

// ---- output Res1d on the tensor core: a k=3 conv over rows (actor, step) is a Linear with three K=128 sources, the
// row itself and its two neighbours (row gathers; -1 = the zero pad at an actor's first / last step)
// conv.weight [128][128][3] -> [co][slot][ci] with slot 0 = centre tap, 1 = previous step, 2 = next step
__global__ void k_pack_conv3(const float* __restrict__ w, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 3 * 128) return;
  const int ci = i & 127, slot = (i >> 7) % 3, co = i / 384;
  const int tap = slot == 0 ? 1 : slot == 1 ? 0 : 2;
  dst[i] = w[((int64_t)co * 128 + ci) * 3 + tap];
}
__global__ void k_actor_conv_idx(int32_t* __restrict__ prev, int32_t* __restrict__ next, int64_t rows) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int l = (int)(r % 20);
  prev[r] = l == 0 ? -1 : (int32_t)r - 1;
  next[r] = l == 19 ? -1 : (int32_t)r + 1;
}
// GroupNorm(1 group over the [128 x 20] outputs of an actor, two-pass like gn_apply) of the raw conv rows h [A*20][128].
// LAST = false: h <- relu(norm * gamma + beta) in place (bn1 of the output Res1d, layers.py:176-178);
// LAST = true : out[a] = relu(norm(h[a, 19]) * gamma + beta + res[a, 19])   (bn2 + shortcut + ReLU, last step only:
//               lanegcn.py:262).  One CTA of 256 threads per actor.
template <bool LAST>
__global__ void __launch_bounds__(256)
k_actor_gn(float* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ beta,
           const float* __restrict__ res, float* __restrict__ out, int64_t a_cap, const int32_t* __restrict__ a_dev) {
  __shared__ float red[8];
  const int64_t a = blockIdx.x;
  if (a >= lgcn_devn(a_dev, a_cap)) return;
  float4* row = reinterpret_cast<float4*>(h + a * 20 * LGCN_C);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 v[3];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int i = threadIdx.x + 256 * j;
    v[j] = i < 640 ? row[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float mean = tot * (1.0f / 2560.0f);
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (threadIdx.x + 256 * j < 640) {
      const float a0 = v[j].x - mean, a1 = v[j].y - mean, a2 = v[j].z - mean, a3 = v[j].w - mean;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) qt += red[w];
  const float rstd = 1.0f / sqrtf(qt * (1.0f / 2560.0f) + LGCN_GN_EPS);
  if (!LAST) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = threadIdx.x + 256 * j;
      if (i < 640) {
        const int c = (i & 31) * 4;
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
        row[i] = relu4(make_float4((v[j].x - mean) * rstd * ga.x + be.x, (v[j].y - mean) * rstd * ga.y + be.y,
                                   (v[j].z - mean) * rstd * ga.z + be.z, (v[j].w - mean) * rstd * ga.w + be.w));
      }
    }
  } else if (threadIdx.x < 32) {
    const int c = threadIdx.x * 4;
    const float4 x = row[19 * 32 + threadIdx.x], r = *reinterpret_cast<const float4*>(res + (a * 20 + 19) * LGCN_C + c);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
    *reinterpret_cast<float4*>(out + a * LGCN_C + c) =
        relu4(make_float4((x.x - mean) * rstd * ga.x + be.x + r.x, (x.y - mean) * rstd * ga.y + be.y + r.y,
                          (x.z - mean) * rstd * ga.z + be.z + r.z, (x.w - mean) * rstd * ga.w + be.w + r.w));
  }
}

// float offsets of the tensor-core images behind the 20 fp32 layers of the pack
constexpr int64_t kOffW3 = spec_offset(20);                        // 2 x [128][3][128] (output.conv1, output.conv2)
constexpr int64_t kOffHi = kOffW3 + 2 * 128 * 384;                 // tf32 hi images: 6 blocks [128][128]
constexpr int64_t kOffLo = kOffHi + 6 * 128 * 128;                 // lo images
constexpr int64_t kPackFloats = kOffLo + 6 * 128 * 128;

std::atomic<int> g_attr[64];

}  // namespace

extern "C" int64_t lgcn_actor_net_wpack_floats(void) { return kPackFloats; }

extern "C" int lgcn_actor_net_pack(const float* const* h_conv_w, const float* const* h_gamma, const float* const* h_beta,
                                   float* wpack, void* stream) {
  LGCN_CHECK_ARG(h_conv_w && h_gamma && h_beta && wpack, "actor_net_pack: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < 20; ++i) {
    const Spec s = spec(i);
    const int cp = cin_padded(s.cin), n = s.k * cp * s.cout;
    LGCN_CHECK_ARG(h_conv_w[i] && h_gamma[i] && h_beta[i], "actor_net_pack: NULL layer %d", i);
    float* dst = wpack + spec_offset(i);
    k_pack_conv<<<lgcn_cdiv(n, 256), 256, 0, st>>>(h_conv_w[i], dst, s.cin, cp, s.cout, s.k);
    LGCN_LAUNCH_OK();
    LGCN_CUDA_OK(cudaMemcpyAsync(dst + n, h_gamma[i], s.cout * 4, cudaMemcpyDeviceToDevice, st));
    LGCN_CUDA_OK(cudaMemcpyAsync(dst + n + s.cout, h_beta[i], s.cout * 4, cudaMemcpyDeviceToDevice, st));
  }
#if LGCN_HAVE_TC
  // the two convs of the output Res1d once more as three-source Linear weights + their tf32 hi / lo images
  LgcnSplitList sl;
  sl.n_blocks = 6;
  for (int j = 0; j < 2; ++j) {
    float* w3 = wpack + kOffW3 + (int64_t)j * 128 * 384;
    k_pack_conv3<<<lgcn_cdiv(128 * 384, 256), 256, 0, st>>>(h_conv_w[18 + j], w3);
    LGCN_LAUNCH_OK();
    for (int sidx = 0; sidx < 3; ++sidx) {
      sl.p[3 * j + sidx] = w3 + sidx * 128;
      sl.ldw[3 * j + sidx] = 384;
    }
  }
  if (int rc = lgcn_split_blocks_many(sl, wpack + kOffHi, wpack + kOffLo, st)) return rc;
#endif
  return 0;
}

static int launch_actor_net(const float* feats, const float* wpack, float* out, int64_t a_cap, const int32_t* a_dev,
                            bool fpn_only, cudaStream_t st) {
  if (a_cap <= 0) return 0;
  int dev = 0;
  LGCN_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && g_attr[dev].exchange(1) == 0) {
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_actor_net<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_actor_net<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  if (fpn_only) k_actor_net<true><<<lgcn_cdiv(a_cap, kActors), kThreads, kSmemBytes, st>>>(feats, wpack, out, a_cap, a_dev);
  else k_actor_net<false><<<lgcn_cdiv(a_cap, kActors), kThreads, kSmemBytes, st>>>(feats, wpack, out, a_cap, a_dev);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_actor_net(const float* feats, const float* wpack, float* out, int64_t n_actors,
                              const int32_t* n_actors_dev, void* stream) {
  LGCN_CHECK_ARG(n_actors >= 0, "actor_net: n_actors %lld", (long long)n_actors);
  LGCN_CHECK_ARG(n_actors == 0 || (feats && wpack && out), "actor_net: NULL argument");
  return launch_actor_net(feats, wpack, out, n_actors, n_actors_dev, false, (cudaStream_t)stream);
}

// ---- the same network with the output Res1d on the tensor core (3xTF32 like every Linear of the path):
//   k_actor_net<FPN_ONLY> -> fpn [A*20,128]  ->  conv1 = 3-source Linear  ->  GN + ReLU per actor  ->  conv2  ->
//   GN + shortcut + ReLU at the last step -> out [A,128]
// workspace: fpn | h1 | h2 [A*20,128] fp32 each | prev, next int32 [A*20]
LinearArgs lgcn_lin1(const float* x, const int32_t* idx, const float* W, const float* gamma, const float* beta,
                     const float* res, int flags, float* out, int64_t m, const int32_t* m_dev);

extern "C" int64_t lgcn_actor_net_tc_workspace_bytes(int64_t n_actors) {
  const int64_t rows = n_actors * 20;
  return 3 * lgcn_align_up(rows * LGCN_C * 4, 1024) + 2 * lgcn_align_up(rows * 4, 1024) + 1024;
}

extern "C" int lgcn_actor_net_tc(const float* feats, const float* wpack, float* out, int64_t n_actors,
                                 const int32_t* n_actors_dev, void* workspace, void* stream) {
#if LGCN_HAVE_TC
  LGCN_CHECK_ARG(n_actors >= 0 && n_actors * 20 < ((int64_t)1 << 31), "actor_net_tc: n_actors %lld", (long long)n_actors);
  LGCN_CHECK_ARG(n_actors == 0 || (feats && wpack && out && workspace), "actor_net_tc: NULL argument");
  LGCN_CHECK_ARG(lgcn_get_gemm_engine() == 1, "actor_net_tc needs the tcgen05 engine (lgcn_actor_net is the fp32 kernel)");
  if (n_actors == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = n_actors * 20, fb = lgcn_align_up(rows * LGCN_C * 4, 1024), ib = lgcn_align_up(rows * 4, 1024);
  float* buf[3] = {(float*)workspace, (float*)((char*)workspace + fb), (float*)((char*)workspace + 2 * fb)};   // fpn, h1, h2
  int32_t* prev = (int32_t*)((char*)workspace + 3 * fb);
  int32_t* next = (int32_t*)((char*)workspace + 3 * fb + ib);
  k_actor_conv_idx<<<lgcn_cdiv(rows, 256), 256, 0, st>>>(prev, next, rows);
  LGCN_LAUNCH_OK();
  if (int rc = launch_actor_net(feats, wpack, buf[0], n_actors, n_actors_dev, true, st)) return rc;
  for (int j = 0; j < 2; ++j) {   // output.conv1 / bn1 / ReLU, then output.conv2 / bn2 + shortcut + ReLU at the last step
    const Spec sp = spec(18 + j);
    const float* gamma = wpack + spec_offset(18 + j) + (int64_t)sp.k * cin_padded(sp.cin) * sp.cout;
    const float* beta = gamma + sp.cout;
    // rows past the live actors (n_actors is a capacity when n_actors_dev is given) are computed too: row-independent
    LinearArgs la = lgcn_lin1(buf[j], nullptr, wpack + kOffW3 + (int64_t)j * 128 * 384, nullptr, nullptr, nullptr, 0,
                              buf[j + 1], rows, nullptr);
    la.n_src = 3;
    la.a[1] = buf[j]; la.idx[1] = prev;
    la.a[2] = buf[j]; la.idx[2] = next;
    la.w_hi = wpack + kOffHi + (int64_t)j * 3 * 128 * 128;
    la.w_lo = wpack + kOffLo + (int64_t)j * 3 * 128 * 128;
    if (int rc = lgcn_launch_linear(la, st)) return rc;
    if (j == 0) k_actor_gn<false><<<(unsigned)n_actors, 256, 0, st>>>(buf[1], gamma, beta, nullptr, nullptr, n_actors, n_actors_dev);
    else k_actor_gn<true><<<(unsigned)n_actors, 256, 0, st>>>(buf[2], gamma, beta, buf[0], out, n_actors, n_actors_dev);
    LGCN_LAUNCH_OK();
  }
  return 0;
#else
  (void)feats; (void)wpack; (void)out; (void)n_actors; (void)n_actors_dev; (void)workspace; (void)stream;
  LGCN_CHECK_ARG(false, "actor_net_tc: built without the tcgen05 engine");
  return -1;
#endif
}
