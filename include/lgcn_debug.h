/*
 * lgcn_debug.h — profiling and ablation switches of liblgcn_b200.so.  NOT part of the product interface (include/lgcn.h):
 * lgcn_debug_flags can make results WRONG on purpose (it switches parts of the tcgen05 kernels off to time the rest);
 * tools/ and bench.py's per-kernel timing pass are the only users.  Process-global, not thread-safe.
 */
#ifndef LGCN_DEBUG_H_
#define LGCN_DEBUG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Profiling ablations of the tcgen05 kernels (results become WRONG unless noted; never set in production; returns
 * the previous value): 1 = epilogue skips staging + TMA stores, 2 = epilogue skips the cross-accumulator reads
 * (gemm_tc*.cu), 4 = MMA issuer skips the tcgen05.mma instructions, 8 = producers skip global loads (and, in
 * gemm_tc*.cu, the conversion), 16 = route the multi-block projection through the generic kernel instead of the
 * A-in-TMEM one (results stay right); aggregate-first kernel only: 32 = no A conversion / tcgen05.st,
 * 64 = no accumulator flushes, 128 = no weight TMA, 256 = write the clock timeline (needs a -DLGCN_TIMELINE build,
 * lgcn_debug_timeline), 512 = flush every 5 keys instead of 3 (results stay within tolerance on the goldens but
 * not on the stress case of tools/precision_fused.py); 32768 = one-block Linears on the first-generation k_linear_tc
 * instead of the linear mode of the aggregate-first kernel (results stay right); 16384 = launch the tcgen05 kernels without
 * the programmatic-dependent-launch attribute (results stay right); 65536 = Att: ctx.1 as its own launch instead of chained
 * inside ctx.0's kernel (results stay right); 131072 = the K=2 heads (MapNet input / seg, Att dist) as their own launches
 * instead of computed inside the following Linear's kernel (results stay right). */
int lgcn_debug_flags(int flags);
/* Profiling aid: device buffer [1024][8] of int64 clock stamps filled by CTA 0 of the aggregate-first LaneConv kernel
 * while debug flag 256 is set (tools/timeline_fused.py). */
int lgcn_debug_timeline(long long* device_buffer);
/* Per-kernel timing for the benchmark: while enabled, the library brackets its LaneConv / Att launches with CUDA
 * events on the launching stream.  on = 1: eager launches only (stream captures are left alone); on = 2: launches
 * INSIDE a stream capture only, as external event-record nodes (cudaEventRecordExternal) of the captured graph, so
 * every replay re-records them and the kernels are timed in the product's own sequence (branch overlap, priorities).
 * lgcn_prof_collect SYNCHRONISES on those events, returns summed milliseconds and launch counts per kind (0 wide
 * projection GEMM, 1 LaneConv gather, 2 ctr2 linear, 3 whole Att layer, 4 aggregate-first LaneConv block incl. its
 * multi-source pre-pass, 5 the block's main kernel alone (k_laneconv_v2); ARRAYS OF 8, the rest reserved) and forgets
 * the events; lgcn_prof_peek does the same without forgetting them (read after each replay of a mode-2 graph).
 * lgcn_prof_enable returns the previous state. */
int lgcn_prof_enable(int on);
int lgcn_prof_collect(double* h_ms_by_kind, int64_t* h_launches_by_kind);
int lgcn_prof_peek(double* h_ms_by_kind, int64_t* h_launches_by_kind);

#ifdef __cplusplus
}
#endif
#endif /* LGCN_DEBUG_H_ */
