/* hannoy_b200_dev.h — development entry points of libhannoy_b200.so: tuning knobs, phase timers, event trace.
 * NOT part of the drop-in boundary (include/hannoy_b200.h): nothing here has a reference counterpart and none of it
 * changes results.  Used by tools/ and by the dev builds of hannoy_b200/build.py. */
#ifndef HANNOY_B200_DEV_H
#define HANNOY_B200_DEV_H
#include "hannoy_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Engine tuning knob (no reference counterpart; never changes results): e.g. "ring_bytes" (shared-memory
 * bytes of rows in flight per query warp), "blocks_per_sm", "fixed_adjacency", "touched_cap".  Takes
 * effect for indexes finalized / workspaces created afterwards; the environment variable HB_<KEY> is the
 * default. */
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
