/* How far mapping_shim.cpp may look ahead in gmapper.c's re_buffer[] from the entry it was called with.
 *
 * The shim is not told the bounds of re_buffer[].  What it knows: a thread's re_buffer is ONE allocation of chunk_size
 * entries for the whole run, zeroed before every fill (gmapper.c:325-332), and handle_read is called for its entries in
 * ascending order, skipping the ones the loop dropped -- each of which is counted in total_reads_dropped /
 * total_pairs_dropped (gmapper.c:510-527).  So, with D the drops counted since this thread's previous call (by any
 * thread: more only makes the bound smaller), `re - D * step` is at or below the first entry of the buffer whenever
 * `re` is the first surviving entry of a chunk, and the buffer's end is at least `re - D * step + chunk_size` there:
 * the MINIMUM of that expression over all calls is a safe end (a call in the middle of a chunk gives a larger value
 * and is ignored by the minimum).  One past the highest entry ever seen is a lower bound of the end as well, and it
 * becomes exact once a full chunk has gone by.
 *
 * (Until round 2 the bound was taken from the current call alone: a batch that had been cut short by another thread's
 * drops was followed by a look-ahead from the middle of the chunk that ran past the buffer.)
 * tests/test_shim_lookahead.py simulates the loop of gmapper.c against this header. */
#ifndef SHRIMP_SHIM_LOOKAHEAD_H
#define SHRIMP_SHIM_LOOKAHEAD_H

namespace shrimp_shim {

template <class Entry>
struct LookaheadBound {
  const Entry *end_min = nullptr, *end_low = nullptr;
  /* entries [re, re + limit) may be read; `drops`: dropped reads / pairs counted since this thread's previous call;
   * `step`: entries per unit (2 in paired mode).  Never less than `step`: the entries of the call itself exist. */
  long long limit(const Entry *re, long long drops, int step, long long chunk_size) {
    if (drops < 0) drops = 0;
    const Entry *cand = re + chunk_size - drops * step;
    if (!end_min || cand < end_min) end_min = cand;
    if (!end_low || re + step > end_low) end_low = re + step;
    const Entry *end = end_min > end_low ? end_min : end_low;
    const long long lim = end - re;
    return lim < step ? step : lim;
  }
};

}  // namespace shrimp_shim

#endif
