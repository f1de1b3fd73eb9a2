/* Shared between mapping_shim.cpp (chunk-level path) and sw_shims.cpp (the per-call symbols and the *_stats
 * functions print_statistics reads, gmapper.c:741-760). */
#ifndef SHRIMP_SHIM_STATE_H
#define SHRIMP_SHIM_STATE_H

#include <stdint.h>

#include "shrimp_b200.h"

namespace shrimp_shim {

/* what one OpenMP thread's chunk calls did, in the units the reference's own counters use */
struct ThreadStats {
  uint64_t vector_calls = 0, vector_cells = 0, vector_bypassed = 0;   /* sw-vector.c:507-509, f1-wrapper.h:110 */
  uint64_t full_calls = 0, full_cells = 0;                            /* sw-full-ls.c:237, :662 */
  uint64_t post_columns = 0;
  uint64_t batches = 0, mispredicted = 0, reads = 0, records = 0;
  double t_prep = 0, t_device = 0, t_first_device = 0, t_build = 0, t_output = 0;   /* host seconds of this thread, SHRIMP_B200_VERBOSE */
  void add(const shrimp_map_stats &s) {
    vector_calls += s.vector_calls;
    vector_cells += s.vector_cells;
    vector_bypassed += s.vector_bypassed;
    full_calls += s.full_calls;
    full_cells += s.full_cells;
    post_columns += s.post_sw_columns;
  }
};
extern thread_local ThreadStats tstats;

/* the calling thread's context of the chunk path (created on first use) */
shrimp_gpu_ctx *thread_ctx();
/* set by mapping_shim.cpp: the calling thread's chunk-path context if it exists (NULL otherwise), so that the
 * *_stats functions of sw_shims.cpp can read its device stage times; sw_shims.cpp also links on its own */
extern shrimp_gpu_ctx *(*chunk_ctx_hook)();
/* set by mapping_shim.cpp: creates the calling thread's chunk-path context.  gmapper.c calls sw_full_{ls,cs}_setup
 * from every -N thread after load_genome and before the mapping clock starts (gmapper.c:2907-2965), which is where
 * the reference allocates its per-thread DP state; the device state (genome upload, projection build) is set up at the
 * same point, so that "Read Mapping Time" (gmapper.c:3015-3021) covers mapping only, as it does in the reference. */
extern void (*chunk_init_hook)();

}  // namespace shrimp_shim

#endif
