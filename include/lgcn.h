/*
 * lgcn.h — C ABI of liblgcn_b200.so: the LaneGCN forward graph path on B200 (sm_100a).
 *
 * The reference (leepaul009/LaneGCN-1) has no plugin/FFI surface: its boundary is the Python module API of
 * lanegcn.py.  Each entry point below replaces the body of one reference function / loop; the reference-side
 * binding is a ctypes stub, the in-repo one is lanegcn-1_b200/_C.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless its name starts with h_.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises unless
 *     the comment says so.  The library never frees or retains caller memory and does not allocate: outputs and
 *     workspaces are caller-provided (sizes from the *_workspace_bytes helpers); the single exception is the scratch
 *     ring of the convenience entry lgcn_linear128 (see lgcn_linear128_ws).  Library state is per device (cached
 *     function attributes, SM count, that ring); entry points are safe to call from one host thread per GPU.
 *   - return 0 on success, <0 on error; lgcn_last_error() gives the thread-local message.
 *   - feature matrices are row-major fp32 with 128 channels (config n_map = n_actor = 128).
 *   - nn.Linear weights are [out, in] row-major (y = x * W^T), as in the reference state_dict.
 */
#ifndef LGCN_H_
#define LGCN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGCN_C 128            /* channels */
#define LGCN_MAX_KEYS 16      /* edge sets per LaneConv block (reference: 14) */

/* epilogue flags of lgcn_linear128 (order of application: GN -> RELU1 -> +RES -> RELU2) */
#define LGCN_EPI_GN    1      /* GroupNorm(1 group) over the 128 outputs: layers.py:72 */
#define LGCN_EPI_RELU1 2      /* ReLU right after the norm: layers.py:84-85 (act=True) */
#define LGCN_EPI_RES   4      /* += res[m,:] : lanegcn.py:360, 707 */
#define LGCN_EPI_RELU2 8      /* final ReLU: lanegcn.py:361, 708 */

int lgcn_version(void);
const char* lgcn_last_error(void);
/* which GEMM engine the library will use on this process: 0 = fp32 SIMT, 1 = tcgen05 3xTF32.
 * set_gemm_engine returns the previous value. */
int lgcn_get_gemm_engine(void);
int lgcn_set_gemm_engine(int engine);

/* number of CUDA kernels this library has launched in this process (all entry points) */
int64_t lgcn_launch_count(void);
/* (profiling / ablation switches live in lgcn_debug.h: they are not part of the product interface) */

/* ------------------------------------------------------------------ host staging
 * Copies n host arrays (h_srcs[i], h_nbytes[i] bytes) back to back into the host arena h_dst (normally pinned
 * memory, dst_bytes capacity) with n_threads copy threads.  Host-only; replaces the per-tensor cudaMemcpyAsync of
 * utils.gpu (utils.py:74-85) together with ONE cudaMemcpyAsync of the arena by the caller.                      */
int lgcn_pack_host(const void* const* h_srcs, const int64_t* h_nbytes, int64_t n, void* h_dst, int64_t dst_bytes,
                   int n_threads);

/* Batch assembly from PACKED SCENES: one contiguous host blob per sample, made once from the preprocessed-pickle schema
 * (preprocess_data.py:78-95; lanegcn.pack_scene):
 *   int64 header[8 + n_kv] = {magic "LGCNSCN1", n_nodes, n_actors, n_scales, idx_bytes, n_index, 0, 0, seg_len[n_kv]}
 *                            with n_kv = 2 (2 n_scales + 2): entries of pre0.u, pre0.v, suc0.u, ..., left.u, .., right.v
 *   float arena  ctrs[2N] feats[2N] turn[2N] control[N] intersect[N] actor_feats[60A] actor_ctrs[2A] rot[4] orig[2]
 *   index arena  the n_kv segments back to back (idx_bytes per entry)
 * Writes the four staging buffers of lgcn_forward in the CAPACITY layout of LgcnForwardArgs (host memory, normally
 * pinned; each then needs ONE H2D copy):
 *   h_fl   float regions at offsets cumsum(2Nc, 2Nc, 2Nc, Nc, Nc, 60Ac, 2Ac, 4Bc, 2Bc)   (Nc/Ac/Bc = capacities)
 *   h_idx  local_idx;   h_t64 = segs (2 n_kv Bc + 1);   h_t32 = node_off[Bc+1] | actor_off[Bc+1] | dims[4]
 * Replaces collate_fn + utils.gpu + utils.to_long + the offset/cat loops (data.py:555-561, utils.py:74-96,
 * lanegcn.py:171-209) for host inputs.  Host-only, n_threads copy threads.                                       */
int lgcn_stage_scenes(const void* const* h_blobs, int n_scenes, int64_t cap_nodes, int64_t cap_actors, int64_t cap_index,
                      int cap_scenes, int n_scales, int idx_bytes, float* h_fl, void* h_idx, int64_t* h_t64,
                      int32_t* h_t32, int n_threads);

/* ------------------------------------------------------------------ graph batching
 * replaces utils.to_long (utils.py:88-96) + the offset/cat loops of graph_gather (lanegcn.py:191-208).
 * `local` holds S segments of scene-local indices (idx_bytes = 2, 4 or 8: int16/int32/int64) laid out
 * back to back in output order; seg_start[S+1] are element offsets, seg_add[S] the node offset of the
 * scene each segment belongs to.  out[i] = (int64) local[i] + seg_add[seg(i)].                           */
int lgcn_offset_indices(const void* local, int idx_bytes, const int64_t* seg_start, const int64_t* seg_add,
                        int n_segments, int64_t total, int64_t* out, void* stream);

/* node-side batching helper: meta[n] = (turn[n,0], turn[n,1], control[n], intersect[n])
 * (the cat at lanegcn.py:387-394, done once per batch).                                                  */
int lgcn_pack_meta(const float* turn, const float* control, const float* intersect, float* meta, int64_t n,
                   void* stream);

/* actor_gather (lanegcn.py:155-168): out[a, c, t] = feats[a, t, c] for the concatenated actor histories
 * ([n_actors, n_steps = 20, n_channels = 3] -> [n_actors, 3, 20], the channels-first layout ActorNet reads).      */
int lgcn_actor_gather(const float* feats, float* out, int64_t n_actors, int n_steps, int n_channels, void* stream);

/* Destination-sorted merged CSR of the K edge sets of one LaneConv block (accumulation order of
 * lanegcn.py:333-354: key order as given, then edge-list order — a STABLE sort by destination u).
 *   h_u[k], h_v[k] : device pointers to int64 u_k (destination) / v_k (source), h_len[k] their lengths
 *   rowptr int32[n_nodes+1];  col int32[E] with col = v*(K+1) + (k+1)  (block index into Y[n, (K+1)*128])
 * workspace: lgcn_csr_workspace_bytes(n_nodes, E).  Indices out of [0,n_nodes) are an error reported
 * through *err_flag (device int32, set non-zero) — checked by the caller when it next synchronises.       */
int64_t lgcn_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges);
int lgcn_csr_build(const int64_t* const* h_u, const int64_t* const* h_v, const int64_t* h_len, int n_keys,
                   int64_t n_nodes, int32_t* rowptr, int32_t* col, void* workspace, int32_t* err_flag,
                   void* stream);

/* Stable destination-sorted CSR of ONE scatter  out[dst[e]] += rows[src[e]]  (what CPU index_add_ does on an unsorted
 * destination index, in edge order): rowptr int32[n_dst+1], col int32[n_edges] = src in CSR order.  dst in [0,n_dst),
 * src in [0,n_src); violations set *err_flag.  Workspace: lgcn_csr_workspace_bytes(n_dst, n_edges).  Used by the
 * LaneRCNN clients (LanePooling scatters on wi, lanercnn.py:506; LaneInput on a2m.v, :327-331).               */
int lgcn_scatter_csr_build(const int64_t* dst, const int64_t* src, int64_t n_edges, int64_t n_dst, int64_t n_src,
                           int32_t* rowptr, int32_t* col, void* workspace, int32_t* err_flag, void* stream);

/* ------------------------------------------------------------------ graph construction before the path
 * Scale-0 pre (is_suc = 0) / suc (is_suc = 1) node edges from the lane topology, in the reference's order
 * (data.py:272-295): per lane the in-lane chain, then one boundary link per lane pair.  lane_idcs int64[n_nodes]
 * (node -> lane, ascending, lanes 0..n_lanes-1), pairs int64[n_pairs,2] (lane, neighbour lane) SORTED by lane
 * (data.py:302-317 appends them lane by lane; *err_flag is set otherwise).  u, v: int64[n_nodes - n_lanes + n_pairs].   */
int64_t lgcn_scale0_workspace_bytes(int64_t n_lanes);
int lgcn_scale0_edges(const int64_t* lane_idcs, int64_t n_nodes, int64_t n_lanes, const int64_t* pairs, int64_t n_pairs,
                      int is_suc, int64_t* u, int64_t* v, void* workspace, int32_t* err_flag, void* stream);
/* Left / right node edges of preprocess() (preprocess_data.py:287-392, cross_angle = None): for every node the nearest
 * node among the lanes reachable as side neighbour or a predecessor / successor of it (side_pairs = left_pairs or
 * right_pairs; lane-level int64[.,2]), kept if closer than cross_dist and heading within pi/4.  u (ascending), v:
 * int64[>= n_nodes]; *h_count (host) receives the edge count — SYNCHRONISES (dataset-build-time code).               */
int64_t lgcn_side_edges_workspace_bytes(int64_t n_nodes, int64_t n_lanes);
int lgcn_side_edges(const float* ctrs, const float* feats, const int64_t* lane_idcs, int64_t n_nodes, int64_t n_lanes,
                    const int64_t* side_pairs, int64_t n_side, const int64_t* pre_pairs, int64_t n_pre,
                    const int64_t* suc_pairs, int64_t n_suc, float cross_dist, int64_t* u, int64_t* v, void* workspace,
                    int64_t* h_count, void* stream);

/* ------------------------------------------------------------------ multi-scale dilation (data.py:520-534)
 * Boolean CSR squaring with scipy's column order (reverse first discovery), bit-exact.  int32 CSR on the device.
 *   lgcn_dilate_csr0   : scale-0 edge list (int64 u = row, v = col; duplicates merged, columns ascending) -> CSR;
 *                        rowptr int32[n+1], col int32[>= n_edges]; *h_nnz = number of stored entries.
 *   lgcn_dilate_bound  : upper bound of nnz(A*A) (also prepares the per-row scratch offsets inside `workspace`).
 *   lgcn_dilate_square : A*A -> rowptr_out/col_out (+ optional COO int64 u_out/v_out, i.e. the `u`,`v` the reference
 *                        returns for that scale); needs cap >= the bound, on the SAME workspace as the bound call.
 * All three SYNCHRONISE the stream to return their host counts (dataset-build-time code, not the forward path).
 * workspace: lgcn_dilate_workspace_bytes(n_nodes, cap) with cap >= max(n_edges, bound).                          */
int64_t lgcn_dilate_workspace_bytes(int64_t n_nodes, int64_t cap);
int lgcn_dilate_csr0(const int64_t* u, const int64_t* v, int64_t n_edges, int64_t n_nodes, int32_t* rowptr,
                     int32_t* col, void* workspace, int64_t* h_nnz, void* stream);
int lgcn_dilate_bound(const int32_t* rowptr, const int32_t* col, int64_t n_nodes, void* workspace, int64_t* h_bound,
                      void* stream);
int lgcn_dilate_square(const int32_t* rowptr, const int32_t* col, int64_t n_nodes, int64_t cap, int32_t* rowptr_out,
                       int32_t* col_out, int64_t* u_out, int64_t* v_out, void* workspace, int64_t* h_nnz, void* stream);

/* ------------------------------------------------------------------ dense pieces
 * out[m, ob*128 + c] (row stride ldo) for ob in [0, n_out_blocks):
 *     acc = sum_s  A_s[ idx_s ? idx_s[m] : m , 0:128 ] . W[ob*128 + c, s*128 : (s+1)*128]
 *         + sum_j  xs[m, j] * W[ob*128 + c, n_src*128 + j]          (j < ks <= 4, optional)
 *     then the epilogue flags (only valid when n_out_blocks == 1).
 * W is [n_out_blocks*128, n_src*128 + ks] row-major.  n_src in 1..3.  idx_s are int32 row gathers.
 * Replaces every bias-free nn.Linear / layers.Linear on the path:
 *   fuse.{ctr,preS,sucS,left,right} (lanegcn.py:332-354, as ONE call with n_out_blocks = 15),
 *   fuse.ctr2 (:359-361), MapNet input.2/seg.2 (:325-327), A2M.meta (:395, ks = 4),
 *   Att.dist.2 / query / ctx.0 (3 sources, no materialised cat) / ctx.1 / agt / linear (:693-708).       */
int lgcn_linear128(const float* a0, const int32_t* idx0, const float* a1, const int32_t* idx1,
                   const float* a2, const int32_t* idx2, int n_src, const float* xs, int ks, const float* W,
                   int n_out_blocks, const float* gamma, const float* beta, const float* res, int flags,
                   float* out, int64_t ldo, int64_t m, void* stream);

/* The same with a caller-provided scratch buffer of lgcn_linear128_workspace_bytes() bytes for the tf32 images of W
 * (the tcgen05 engine splits W before the GEMM).  lgcn_linear128 itself is the convenience form: without a workspace it
 * uses a small scratch ring that the library allocates once per device on first use — the ONE place the library owns
 * device memory; every fused entry point (lgcn_forward, lgcn_att_forward, lgcn_laneconv_stack*) takes its scratch from
 * the caller.  The workspace must stay untouched until the call's kernels have run (stream order is enough).       */
int64_t lgcn_linear128_workspace_bytes(void);
int lgcn_linear128_ws(const float* a0, const int32_t* idx0, const float* a1, const int32_t* idx1,
                      const float* a2, const int32_t* idx2, int n_src, const float* xs, int ks, const float* W,
                      int n_out_blocks, const float* gamma, const float* beta, const float* res, int flags,
                      float* out, int64_t ldo, int64_t m, void* workspace, void* stream);

/* h[m,:] = relu(W1[128,2] . x[m] + b1), x[m] = p[ip ? ip[m] : m] - (q ? q[iq ? iq[m] : m] : 0)
 * The nn.Linear(2,128)+ReLU heads of MapNet.input/seg (lanegcn.py:277-286) and Att.dist (:644-648, with
 * x = agt_ctrs[hi] - ctx_ctrs[wi], :693).                                                                */
int lgcn_mlp2_in(const float* p, const int32_t* ip, const float* q, const int32_t* iq, const float* W1,
                 const float* b1, float* h, int64_t m, void* stream);

/* Same for 4-float rows: h[m,:] = relu(W1[128,4] . (p[ip[m]] - q[iq[m]]) + b1) — LanePooling.relpose on pose
 * differences (lanercnn.py:443-446, 494-495).                                                              */
int lgcn_mlp4_in(const float* p, const int32_t* ip, const float* q, const int32_t* iq, const float* W1,
                 const float* b1, float* h, int64_t m, void* stream);

/* LaneConv gather-reduce with fused GroupNorm(1)+ReLU (lanegcn.py:333-357 after the wide projection):
 *   t = Y[n, 0:128];  for e in rowptr[n]..rowptr[n+1]: t += Y_blocks[col[e]];  out[n] = relu(GN(t))
 * Y is [n_nodes, n_blocks*128]; Y_blocks[b] = Y + b*128 floats.  Fixed summation order => deterministic. */
int lgcn_laneconv_gather_gn_relu(const float* Y, int n_blocks, const int32_t* rowptr, const int32_t* col,
                                 const float* gamma, const float* beta, float* out, int64_t n_nodes,
                                 void* stream);

/* General form: out[r] = relu(GN(base[r*base_ld ..+128] + sum_e blocks[col[e]*128 ..+128])) over rowptr[r]..rowptr[r+1],
 * in CSR order.  Replaces index_add_ on an UNSORTED destination index (LanePooling scatters on wi, lanercnn.py:506;
 * LaneInput on a2m.v, :327-331) once the caller has a stable destination-sorted CSR (lgcn_csr_build).        */
int lgcn_gather_rows_gn_relu(const float* base, int64_t base_ld, const float* blocks, const int32_t* rowptr,
                             const int32_t* col, const float* gamma, const float* beta, float* out, int64_t n_rows,
                             void* stream);

/* Att scatter (lanegcn.py:702-705): out[r] = relu(GN(a[r] + sum_{p in rowptr[r]..rowptr[r+1]} c[p])).
 * Pairs are destination-sorted already (hi ascending), so the segments are contiguous rows of c.         */
int lgcn_segsum_gn_relu(const float* a, const float* c, const int32_t* rowptr, const float* gamma,
                        const float* beta, float* out, int64_t n_rows, void* stream);

/* ------------------------------------------------------------------ Att pair list (lanegcn.py:672-689)
 * Scenes b = 0..B-1 own agent rows agt_off[b]..agt_off[b+1] and context rows ctx_off[b]..ctx_off[b+1]
 * (int32[B+1], device).  A pair (i,j) of the same scene exists iff sqrt(dx*dx + dy*dy) <= th, evaluated in
 * fp32 with separately rounded sub / mul / add / sqrt (no FMA contraction) like the torch ops it replaces.
 * Step 1 (count): per-agent counts -> rowptr int32[n_agt+1] (exclusive scan; rowptr[n_agt] = P) and the
 *                 reference's empty-scene offset quirk resolved into per-scene output
 *                 offsets.  *h_total (host int64, may be NULL) receives P — passing it SYNCHRONISES.
 * Step 2 (fill):  hi/wi in row-major order.  hi32/wi32/hi64/wi64 may each be NULL.
 * keep_quirk != 0 reproduces the reference (scenes after an empty scene are shifted down).
 * rowptr is indexed by OUTPUT destination row (the value of hi), so it feeds lgcn_segsum_gn_relu as is. */
int64_t lgcn_pairs_workspace_bytes(int64_t n_agt, int n_scenes);
int lgcn_pairs_count(const float* agt_ctrs, const float* ctx_ctrs, const int32_t* agt_off,
                     const int32_t* ctx_off, int n_scenes, int64_t n_agt, float th, int keep_quirk,
                     int32_t* rowptr, void* workspace, int64_t* h_total, void* stream);
int lgcn_pairs_fill(const float* agt_ctrs, const float* ctx_ctrs, const int32_t* agt_off,
                    const int32_t* ctx_off, int n_scenes, int64_t n_agt, float th, const void* workspace,
                    int32_t* hi32, int32_t* wi32, int64_t* hi64, int64_t* wi64, void* stream);

/* ------------------------------------------------------------------ fused sequences (one call per module)
 * LaneConv stack: the 4-block loop of MapNet.forward (lanegcn.py:331-362) == M2M.forward (:448-479).
 * wpack: per block, contiguous fp32
 *     Wcat[(K+1)*128,128] (rows: ctr, then the K keys in accumulation order) | Wctr2[128,128]
 *     | norm.weight[128] | norm.bias[128] | ctr2.norm.weight[128] | ctr2.norm.bias[128]
 * feat is updated in place ([n_nodes,128]).  workspace: lgcn_laneconv_workspace_bytes(n_nodes, K).       */
int64_t lgcn_laneconv_wpack_floats(int n_keys);
int64_t lgcn_laneconv_workspace_bytes(int64_t n_nodes, int n_keys);
int lgcn_laneconv_stack(float* feat, const int32_t* rowptr, const int32_t* col, int n_keys, int n_blocks,
                        const float* wpack, int64_t n_nodes, void* workspace, void* stream);

/* The same stack, aggregate-first: each block is ONE tcgen05 kernel that gathers the neighbour rows of feat straight
 * into the GEMM operand (sum_k W_k . sum_{e in key k} feat[src(e)]), so the [n_nodes, (K+1)*128] projection is never
 * written to memory; ctr2 + GroupNorm + residual + ReLU run in the same kernel.  The plan is static per graph
 * (built from the merged CSR of lgcn_csr_build; n_edges = its entry count) and shared by every block and by
 * MapNet and M2M.  Same wpack, same result up to fp32 summation order.  tcgen05 engine only.
 * plan: lgcn_laneconv_plan_bytes; workspace: lgcn_laneconv_planned_workspace_bytes; n_blocks <= 8.           */
int64_t lgcn_laneconv_plan_bytes(int64_t n_nodes, int64_t n_edges, int n_keys);
int lgcn_laneconv_plan_build(const int32_t* rowptr, const int32_t* col, int n_keys, int64_t n_nodes,
                             int64_t n_edges, void* plan, void* stream);
int64_t lgcn_laneconv_planned_workspace_bytes(int64_t n_nodes, int64_t n_edges, int n_keys);
int lgcn_laneconv_stack_planned(float* feat, void* plan, int64_t n_edges, int n_keys, int n_blocks,
                                const float* wpack, int64_t n_nodes, void* workspace, void* stream);

/* One Att layer (lanegcn.py:662-710) on a prebuilt pair list.  wpack, contiguous fp32:
 *   dist.0.weight[128,2] | dist.0.bias[128] | dist.2.linear.weight[128,128] | dist.2.norm.{weight,bias}
 *   | query.linear.weight[128,128] | query.norm.{w,b} | ctx.0.linear.weight[128,384] | ctx.0.norm.{w,b}
 *   | ctx.1.weight[128,128] | agt.weight[128,128] | norm.{w,b} | linear.linear.weight[128,128]
 *   | linear.norm.{w,b}
 * agts_in/agts_out [n_agt,128] (may alias), ctx [n_ctx,128]; n_pairs may be 0 only together with
 * n_ctx == 0 (the reference's early-out path :664-670, which skips self.norm).                          */
int64_t lgcn_att_wpack_floats(void);
int64_t lgcn_att_workspace_bytes(int64_t n_agt, int64_t n_pairs);
int lgcn_att_forward(const float* agts_in, float* agts_out, const float* ctx, const float* agt_ctrs,
                     const float* ctx_ctrs, const int32_t* hi, const int32_t* wi, const int32_t* rowptr,
                     int64_t n_agt, int64_t n_ctx, int64_t n_pairs, const float* wpack, void* workspace,
                     void* stream);

/* ------------------------------------------------------------------ the whole forward graph path in ONE call
 * graph_gather (utils.to_long + lanegcn.py:171-209) -> CSR + gather plan -> the three Att pair lists (:672-689) ->
 * MapNet (:311-363) -> A2M (:385-407) -> M2M (:445-480) -> M2A (:502-513) -> A2A (:534-545), i.e. lanegcn.py:134-141
 * of Net.forward, enqueued on `stream` with NO host synchronisation and no data-dependent host decision: every row
 * count that depends on the batch (nodes, actors, pairs) is read from DEVICE memory, the buffers are sized by the
 * capacities below, and the call is therefore capturable in a CUDA graph that is replayed for any batch that fits the
 * capacities.  tcgen05 engine only.
 *
 *   dims        device int32[4]: {n_nodes, n_actors, 0, 0} of this batch (<= cap_nodes / cap_actors)
 *   node_off / actor_off   device int32[cap_scenes + 1]: row offsets per scene, PADDED with the totals
 *   local_idx   scene-local edge indices (idx_bytes = 2/4/8), segments back to back in output order
 *               (key: pre0,suc0,...,left,right; then u|v; then scene slot 0..cap_scenes-1);
 *   segs        device int64[2 S + 1], S = 2 * (2 n_scales + 2) * cap_scenes: seg_start[S + 1] | seg_add[S]
 *               (lgcn_offset_indices); slots of absent scenes are empty segments
 *   nodes       [cap_nodes,128] in/out: MapNet writes it, A2M / M2M update it
 *   actors      [cap_actors,128] in/out: ActorNet's output on entry, M2A / A2A update it
 *   stages      bit mask of LGCN_STAGE_* (a tap after any stage = run the mask in pieces)
 *   status      device int32[8], zeroed by the GRAPH stage: [0] flag bits LGCN_ST_*, [1..3] exact pair counts of
 *               A2M / M2A / A2A, [4] != 0 if an edge index was out of range.  The caller reads it when it next
 *               synchronises: an overflow means the pair capacity was too small (results invalid: rerun larger).
 *   w           the module weights (fp32, state_dict layouts named below) + `prepared`: their tf32 hi / lo images
 *               written by lgcn_forward_prepare (once per weight version; lgcn_forward_prepared_bytes).           */
#define LGCN_STAGE_GRAPH 1
#define LGCN_STAGE_MAPNET 2
#define LGCN_STAGE_A2M 4
#define LGCN_STAGE_M2M 8
#define LGCN_STAGE_M2A 16
#define LGCN_STAGE_A2A 32
#define LGCN_STAGE_ALL 63
#define LGCN_ST_OVERFLOW_A2M 1
#define LGCN_ST_OVERFLOW_M2A 2
#define LGCN_ST_OVERFLOW_A2A 4
#define LGCN_ST_EMPTY_A2M 8
#define LGCN_ST_EMPTY_M2A 16
#define LGCN_ST_EMPTY_A2A 32

typedef struct LgcnForwardWeights {
  const float* map_input;   /* input.0.weight[128,2] | input.0.bias[128] | input.2.linear.weight[128,128] | input.2.norm.{weight,bias} */
  const float* map_seg;     /* same layout for MapNet.seg                                                              */
  const float* map_fuse;    /* 4 blocks of the LaneConv wpack (lgcn_laneconv_wpack_floats)                             */
  const float* a2m_meta;    /* meta.linear.weight[128,132] | meta.norm.weight[128] | meta.norm.bias[128]               */
  const float* att[6];      /* Att wpack (lgcn_att_wpack_floats): a2m.att.0, a2m.att.1, m2a.att.0/1, a2a.att.0/1       */
  const float* m2m_fuse;    /* 4 blocks of the LaneConv wpack                                                          */
  void* prepared;           /* lgcn_forward_prepared_bytes(n_scales) bytes, written by lgcn_forward_prepare            */
} LgcnForwardWeights;

typedef struct LgcnForwardArgs {
  int64_t cap_nodes, cap_actors, cap_index;   /* cap_index: capacity of local_idx in entries (2 per edge)          */
  int64_t cap_pairs[3];                       /* A2M, M2A, A2A                                                     */
  int32_t cap_scenes, n_scales, idx_bytes, keep_pair_quirk;
  float dist_th[3];                           /* actor2map_dist, map2actor_dist, actor2actor_dist                  */
  int32_t stages;
  const int32_t* dims;
  const int32_t* node_off;
  const int32_t* actor_off;
  const float* node_ctrs;                     /* [n_nodes,2]                                                       */
  const float* node_feats;                    /* [n_nodes,2]                                                       */
  const float* turn;                          /* [n_nodes,2]                                                       */
  const float* control;                       /* [n_nodes]                                                         */
  const float* intersect;                     /* [n_nodes]                                                         */
  const float* actor_ctrs;                    /* [n_actors,2]                                                      */
  const void* local_idx;
  const int64_t* segs;
  float* nodes;
  float* actors;
  int32_t* status;
  LgcnForwardWeights w;
  void* workspace;                            /* lgcn_forward_workspace_bytes(&args)                               */
  void* aux_streams[2];                       /* optional (NULL = none): two more streams of the same device.  Kernels
                                                 that do not depend on each other are forked onto them and joined back
                                                 inside the call — the pair lists beside the CSR / plan build, the
                                                 query and agt Linears of an Att layer beside its pair-side chain — so
                                                 they become parallel branches of a captured graph.  The library keeps
                                                 three fork / join events per device for this (created on first use). */
} LgcnForwardArgs;

/* ------------------------------------------------------------------ ActorNet (lanegcn.py:212-263) as ONE kernel
 * feats [n_actors, 20, 3] (the concatenated per-actor histories as the dataset stores them: step-major, i.e. BEFORE the
 * transpose of actor_gather) -> out [n_actors, 128] = output Res1d at the last step.  fp32 FMA arithmetic.
 * wpack (lgcn_actor_net_wpack_floats floats) is written by lgcn_actor_net_pack from the 20 conv layers in this order:
 *   groups.g.0.{conv1,conv2,downsample.0}, groups.g.1.{conv1,conv2}  for g = 0,1,2;  lateral.0, lateral.1, lateral.2;
 *   output.{conv1,conv2}
 * with h_conv_w[i] the Conv1d weight [C_out, C_in, k] and h_gamma[i] / h_beta[i] the GroupNorm that follows it
 * (HOST arrays of 20 DEVICE pointers).  n_actors_dev (may be NULL): live actor count in device memory.               */
int64_t lgcn_actor_net_wpack_floats(void);
int lgcn_actor_net_pack(const float* const* h_conv_w, const float* const* h_gamma, const float* const* h_beta,
                        float* wpack, void* stream);
int lgcn_actor_net(const float* feats, const float* wpack, float* out, int64_t n_actors, const int32_t* n_actors_dev,
                   void* stream);
/* The same network with its output Res1d (46 % of the FLOPs: two 128 -> 128, k = 3 convs over all 20 steps) on the tensor
 * core: the fp32 kernel stops after the feature pyramid, each conv is a Linear with three K = 128 sources (a row and its
 * two neighbours in time, zero rows at an actor's ends; 3xTF32 like every Linear of the path), GroupNorm over an actor's
 * [128 x 20] outputs in a small kernel.  Same wpack; tcgen05 engine only.  workspace: lgcn_actor_net_tc_workspace_bytes. */
int64_t lgcn_actor_net_tc_workspace_bytes(int64_t n_actors);
int lgcn_actor_net_tc(const float* feats, const float* wpack, float* out, int64_t n_actors, const int32_t* n_actors_dev,
                      void* workspace, void* stream);

/* ------------------------------------------------------------------ PredNet + AttDest + sort + world transform, ONE kernel
 * (lanegcn.py:575-631, 713-737, 145-150).  actors [n,128], actor_ctrs [n,2] -> cls [n,6] (descending), reg [n,6,30,2]
 * (trajectories in the order of their scores; in world coordinates reg . rot[b] + orig[b] of the actor's scene b when
 * rot != NULL, else in scene coordinates as PredNet.forward returns them).  actor_off int32[n_scenes + 1].
 * wpack (lgcn_pred_net_wpack_floats floats) is written by lgcn_pred_net_pack from 64 DEVICE pointers (HOST array):
 *   for m in 0..5: pred.m.0.linear1.weight, pred.m.0.linear2.weight, pred.m.0.norm1.{weight,bias},
 *                  pred.m.0.norm2.{weight,bias}, pred.m.1.weight [60,128], pred.m.1.bias
 *   att_dest.dist.0.{weight [128,2],bias}, att_dest.dist.2.linear.weight, att_dest.dist.2.norm.{weight,bias},
 *   att_dest.agt.linear.weight [128,256], att_dest.agt.norm.{weight,bias}
 *   cls.0.linear1.weight, cls.0.linear2.weight, cls.0.norm1.{weight,bias}, cls.0.norm2.{weight,bias}, cls.1.weight, cls.1.bias */
int64_t lgcn_pred_net_wpack_floats(void);
int lgcn_pred_net_pack(const float* const* h_params, float* wpack, void* stream);
int lgcn_pred_net(const float* actors, const float* actor_ctrs, const int32_t* actor_off, int n_scenes, const float* rot,
                  const float* orig, const float* wpack, float* cls, float* reg, int64_t n_actors,
                  const int32_t* n_actors_dev, void* stream);

/* World transform of the predictions (lanegcn.py:145-150): reg[a, :, :, :] <- reg[a] . rot[b] + orig[b] with b the
 * scene of actor a (actor_off int32[n_scenes + 1], padded with the total), in place; reg holds points_per_actor
 * (= num_mods x num_preds) float2 per actor.  n_actors_dev (may be NULL): live actor count in device memory.      */
int lgcn_world_transform(float* reg, const int32_t* actor_off, int n_scenes, const float* rot, const float* orig,
                         int64_t n_actors, const int32_t* n_actors_dev, int points_per_actor, void* stream);

int64_t lgcn_forward_prepared_bytes(int n_scales);
int lgcn_forward_prepare(const LgcnForwardWeights* w, int n_scales, void* stream);
int64_t lgcn_forward_workspace_bytes(const LgcnForwardArgs* args);
int lgcn_forward(const LgcnForwardArgs* args, void* stream);
/* Device pointers into the workspace for parity tests of the index side (valid after the GRAPH stage): which = 0 CSR
 * rowptr int32[n+1], 1 CSR col int32[E], 2 batched int64 u|v arena (e64), 3/4/5 hi int32[P] of A2M / M2A / A2A,
 * 6/7/8 wi int32[P], 9/10/11 destination rowptr of the three lists.  Returns NULL for an unknown selector.       */
void* lgcn_forward_buffer(const LgcnForwardArgs* args, int which);

#ifdef __cplusplus
}
#endif
#endif /* LGCN_H_ */
