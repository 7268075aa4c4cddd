/* hannoy_b200_dev.h — development entry points of libhannoy_b200.so: tuning knobs, phase timers, event trace.
 * NOT part of the drop-in boundary (include/hannoy_b200.h): nothing here has a reference counterpart and none of it
 * changes results.  Used by tools/ and by the dev builds of hannoy_b200/build.py. */
#ifndef HANNOY_B200_DEV_H
#define HANNOY_B200_DEV_H
#include "hannoy_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Engine tuning knob (no reference counterpart; never changes results).  Takes effect for indexes finalized / workspaces
 * created afterwards; the environment variable HB_<KEY> is the default.  Keys (default):
 *   ring_bytes (8192 for rows of at most 2 KB, else 12288)  shared-memory bytes of rows in flight per query warp
 *   ring_short (1 for rows of at most 2 KB)   the f32 ring kernel's instantiation compiled for 4 CTAs/SM instead of 3
 *   gather4 (1)            rows of at most 1 KB gathered four per TMA instruction
 *   ring_min_row (0)       shorter rows are gathered with plain loads instead of the ring
 *   bin_wide (1)           binary kernel: the 128-register instantiation when shared memory limits the SM to 4 CTAs anyway
 *   vis_atomic (0 binary / 1 f32)  visited set by atomicOr with return (1) or by read + conditional reduction (0)
 *   defer (1)              layer 0: decide the next pop before the pending heap update is applied
 *   team (1)               idle warps gather rows for the walks of their CTA
 *   blocks_per_sm (64)     cap on resident CTAs per SM
 *   fixed_adjacency (1)    fixed-stride copy of layer 0 (one or two 128-byte lines per item)
 *   touched_cap (16384)    visited-set entries cleared one by one before the whole bitset is cleared instead
 *   exact_tc (1), exact_tc_min_pairs (1 << 22)   exact k-NN through the tensor-core shortlist
 *   build_inflight_div (64), build_sync (0), build_link_blocks (0)   graph builder: batch size rule, debugging aids */
hb_status hb_tune(const char* key, int value);

/* Development aid: cycles per phase of the search kernel summed over all queries since the last call
 * (stage, upper layers, adjacency wait, visited filter, row gather+distance, heap update, tail, total, helper jobs posted,
 * helper results collected, accept step, pop decision; 16 slots).
 * All zero unless the library was built with -DHB_PHASES. */
void hb_debug_phases(uint64_t* out16);
/* Development aid: event trace (clock64 << 8 | event id) of one query warp since the last call; returns the
 * number of records copied.  Always 0 unless the library was built with -DHB_TRACE. */
uint32_t hb_debug_trace(uint64_t* out, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif
