"""The look-ahead bound of the mapping shim (integration/shim_lookahead.h) against a simulation of gmapper.c's loop
(gmapper.c:325-607): several threads, each with ONE buffer of chunk_size entries that it refills chunk after chunk; the
loop drops some entries (one shared counter of drops) and calls handle_read for the others in ascending order; a batch
mapped by one call covers the entries up to the bound, so the next call that computes a bound may come from the middle
of a chunk.  Whatever the interleaving of the threads: the bound never reaches past the thread's buffer, and without
drops it reaches the end of the chunk."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "shim_lookahead.h"
struct Entry { char pad[256]; };
using shrimp_shim::LookaheadBound;
static unsigned long long s = 88172645463325252ull;
static unsigned rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (unsigned)(s >> 33); }
struct Thread {
  std::vector<Entry> buf;
  LookaheadBound<Entry> bound;
  long long snapshot = 0;        // the drop counter at this thread's previous call (t_drop_snapshot)
  long long covered_until = -1;  // entries below this index belong to the batch in flight
  int load = 0, next = 0;        // entries of the current chunk, next entry of the loop
  std::vector<char> dropped;
};
int main(int argc, char **argv) {
  const int trials = argc > 1 ? atoi(argv[1]) : 2000;
  long long overruns = 0, full_without_drops = 0, calls = 0, cut_short = 0;
  for (int trial = 0; trial < trials; trial++) {
    const int n_threads = 1 + rnd() % 4, step = (trial & 1) ? 2 : 1;
    const int chunk = step * (2 + rnd() % 40);
    const int drop_pct = (trial % 3 == 0) ? 0 : rnd() % 40;
    long long total_drops = 0;
    std::vector<Thread> T(n_threads);
    for (auto &t : T) t.buf.resize(chunk);
    for (int iter = 0; iter < 4000; iter++) {
      Thread &t = T[rnd() % n_threads];       // any interleaving of the threads' steps
      if (t.next >= t.load) {                  // fill the buffer: a full chunk, or a last short one
        t.load = (rnd() % 8 == 0) ? step * (int)(1 + rnd() % (chunk / step)) : chunk;
        t.next = 0;
        t.covered_until = -1;
        t.dropped.assign(t.load, 0);
        for (int i = 0; i < t.load; i += step)
          if ((int)(rnd() % 100) < drop_pct)
            for (int k = 0; k < step; k++) t.dropped[i + k] = 1;
        continue;
      }
      const int i = t.next;
      t.next += step;
      if (t.dropped[i]) {                      // gmapper.c:510-527
        total_drops++;
        continue;
      }
      // handle_read(&buf[i]): a new look-ahead only when the entry is not covered by the batch in flight
      if (i < t.covered_until) { t.snapshot = total_drops; continue; }
      const Entry *re = &t.buf[i];
      const long long lim = t.bound.limit(re, total_drops - t.snapshot, step, chunk);
      calls++;
      if (re + lim > t.buf.data() + chunk) overruns++;
      if (lim < step) overruns++;
      if (drop_pct == 0 && i == 0 && t.load == chunk && lim == chunk) full_without_drops++;
      if (i + lim < t.load) cut_short++;
      t.covered_until = i + lim;
      t.snapshot = total_drops;
    }
  }
  printf("%lld calls, %lld overruns, %lld full chunks without drops, %lld batches cut short\n", calls, overruns,
         full_without_drops, cut_short);
  return overruns != 0 || full_without_drops == 0 || cut_short == 0;
}
'''


def test_lookahead_bound_never_leaves_the_buffer(tmp_path):
    src = tmp_path / "sim.cpp"
    src.write_text(SRC)
    exe = str(tmp_path / "sim")
    subprocess.run(["g++", "-O2", "-std=gnu++17", "-I", os.path.join(ROOT, "integration"), str(src), "-o", exe], check=True)
    r = subprocess.run([exe, "3000"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert " 0 overruns" in r.stdout, r.stdout
