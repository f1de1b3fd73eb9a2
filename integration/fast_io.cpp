/*
 * fast_io.cpp -- SURVEY section 8 row f2: the two host-side string paths that bound the drop-in once the mapping
 * itself runs on the GPU (the serial FASTA/FASTQ reader inside gmapper.c's
 * `omp critical (fill_reads_buffer)`, gmapper.c:338, and the SAM record formatter).
 *
 *     bool fasta_get_next_read_with_range(fasta_t, read_entry *)                        common/fasta.c:316
 *     void hit_output(read_entry *, read_hit *, read_hit *, bool, int *, int, bool)     gmapper/output.c:227
 *     uint32_t *fasta_sequence_to_bitfield(fasta_t, char *)                             common/fasta.c:610
 *     uint32_t *reverse_complement_read_cs(uint32_t *, int8_t, int8_t, uint32_t, bool)  common/util.c:601
 * (the last two are the per-read packing of gmapper.c:475-488: one out-of-line call per base in the reference, 0.5 us
 * per read -- as much as the device spends on the whole read)
 *
 * Both carry the reference's C++ linkage and signatures.  integration/Makefile links them over the reference's own
 * definitions, which `objcopy --weaken-symbol` turns weak in a COPY of the reference's unchanged fasta.o / output.o
 * (all references in output.o itself go through the symbol, R_X86_64_PLT32, so read_output / readpair_output --
 * output.c:955, :1070, compiled unchanged -- call this formatter).  The reference's hit_output stays reachable as
 * `shrimp_ref_hit_output` (objcopy --add-symbol at its address): the output forms this file does not restate (the
 * old SHRiMP format of -E-less runs, --bfast, --extra-sam-fields) and anything unusual go there.
 *
 * Reader: the reference copies every byte three times (gzread block -> line buffer -> parse buffer -> malloc'd
 * string, util.c:949-1040, fasta.c:344-500), one byte at a time; here lines are found with memchr in one large block
 * read straight from the same gzFile (plain or gzip input alike) and copied once.  Same strings, same trailing
 * 17 zero bytes, same `free()`-able allocations, same return values and messages on malformed input.
 *
 * Formatter: the reference builds a record with ~12 snprintf calls and ~10 malloc/free pairs (cigar_t, cigar
 * string, VLAs, reverse-complement scratch); here one pass writes the bytes into the thread's output buffer.
 * Everything that decides a byte of the record (flag bits, POS of a reverse-strand hit, CIGAR runs, the IUPAC
 * rule of the SEQ column, Z0..Z6 through the same libm `log`) follows output.c line by line, cited below.
 */
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include <ctype.h>
#include <math.h>
#include <sched.h>
#include <unistd.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <emmintrin.h>

#include "gmapper/gmapper.h"
#include "gmapper/output.h"
#include "common/fasta.h"
#include "common/util.h"
#include "common/my-alloc.h"
#include "common/sw-full-common.h"

// the reference's own hit_output (integration/Makefile: objcopy --add-symbol on the copy of output.o)
extern "C" void shrimp_ref_hit_output(struct read_entry *, struct read_hit *, struct read_hit *, bool, int *, int, bool);

namespace {

// ================================================================================================================
// Reader
// ================================================================================================================
struct Reader {
  fasta_t owner = NULL;
  char *buf = NULL;
  size_t cap = 0, pos = 0, end = 0;
  bool eof = false;
  bool header = true;     // util.c:985: '#' lines before the first record are dropped by the line reader itself
  // one entry's text, assembled when a line spans block boundaries or an entry has several sequence lines
  char *acc = NULL;
  size_t acc_cap = 0;
};

enum { MAX_READERS = 16 };
Reader g_readers[MAX_READERS];
// SHRIMP_B200_VERBOSE: seconds spent in the reader (it runs inside gmapper.c's critical section: the serial part of a run)
unsigned long long g_reader_ticks, g_reader_calls, g_tick0;
double g_wall0;

Reader *reader_of(fasta_t f) {
  for (int i = 0; i < MAX_READERS; i++)
    if (g_readers[i].owner == f) return &g_readers[i];
  for (int i = 0; i < MAX_READERS; i++)
    if (g_readers[i].owner == NULL) {
      Reader *r = &g_readers[i];
      r->owner = f;
      r->cap = (size_t)16 << 20;
      if (const char *e = getenv("SHRIMP_B200_READER_BLOCK")) r->cap = (size_t)std::max(256L, atol(e));   // (tests: entries across block ends)
      r->buf = (char *)xmalloc(r->cap + 64);   // 16-byte loads of scan_line may start at the last byte
      r->pos = r->end = 0;
      r->eof = false;
      r->header = f->header;
      return r;
    }
  fprintf(stderr, "error: more than %d read files open\n", (int)MAX_READERS);
  exit(1);
}

// more bytes behind the unread tail; false at the end of the file
bool refill(Reader *r) {
  if (r->eof) return false;
  if (r->pos > 0) {
    memmove(r->buf, r->buf + r->pos, r->end - r->pos);
    r->end -= r->pos;
    r->pos = 0;
  }
  if (r->end == r->cap) return false;   // a line longer than the block: the caller takes it in pieces
  const int want = (int)((r->cap - r->end) > ((size_t)1 << 30) ? ((size_t)1 << 30) : (r->cap - r->end));
  const int got = gzread(r->owner->fp, r->buf + r->end, (unsigned)want);
  if (got <= 0) {
    r->eof = true;
    return false;
  }
  r->end += (size_t)got;
  return true;
}

// The next piece of input as fast_gzgets_safe (util.c:949) would hand it out: up to and including a newline, or what
// is left of an over-long line / of the file.  *nl tells whether the piece ends a line.  NULL at the end of the file.
const char *next_piece(Reader *r, size_t *len, bool *nl) {
  for (;;) {
    const char *s = r->buf + r->pos;
    const char *e = (const char *)memchr(s, '\n', r->end - r->pos);
    if (e) {
      *len = (size_t)(e - s);
      *nl = true;
      r->pos += *len + 1;
      if (r->header) {
        if (*len > 0 && s[0] == '#') continue;   // util.c:985-990
        r->header = false;
      }
      return s;
    }
    if (refill(r)) continue;
    if (r->end > r->pos) {   // block full without a newline, or the last line of a file that does not end in one
      const char *piece = r->buf + r->pos;   // (refill may have moved the tail to the front of the block)
      *len = r->end - r->pos;
      *nl = false;
      r->pos = r->end;
      return piece;
    }
    return NULL;
  }
}

// first byte of the next piece without taking it; -1 at the end of the file.  (Only used behind a name line, where
// the header comments of util.c:985 are over.)
int peek_first(Reader *r) {
  for (;;) {
    if (r->pos < r->end) return (unsigned char)r->buf[r->pos];
    if (!refill(r)) return -1;
  }
}

void acc_reserve(Reader *r, size_t n) {
  if (n <= r->acc_cap) return;
  size_t c = r->acc_cap ? r->acc_cap : 4096;
  while (c < n) c *= 2;
  r->acc = (char *)xrealloc(r->acc, c);
  r->acc_cap = c;
}

// fasta.c's loops stop copying a piece at a NUL byte (`fasta->buffer[i] != '\0'`)
inline size_t text_len(const char *s, size_t len) {
  const void *z = memchr(s, 0, len);
  return z ? (size_t)((const char *)z - s) : len;
}

char *dup17(const char *s, size_t len) {   // "allocate extra space to appease valgrind", fasta.c:381-383
  char *p = (char *)xmalloc(len + 17);
  memcpy(p, s, len);
  memset(p + len, 0, 17);
  return p;
}

// extract_name, fasta.c:243-281: the text after '>' / '@' up to the first tab, trimmed, cut at its first blank; a
// second tab-separated field is the range string
char *extract_name_fast(char *line, char **ranges) {
  char *save = NULL;
  char *tok = strtok_r(line + 1, "\t", &save);
  tok = strtrim(tok);
  const int full = (int)strlen(tok);
  int len = 0;
  while (len < full && tok[len] != ' ' && tok[len] != '\t') len++;
  char *ret = dup17(tok, (size_t)len);
  if (ranges != NULL && (tok = strtok_r(NULL, "\t", &save)) != NULL) *ranges = dup17(tok, strlen(tok));
  return ret;
}

}  // namespace

static bool next_read(fasta_t fasta, read_entry *re);
struct Views {   // the strings of one entry, in the reader's block (fast path) or in strings next_read() made
  const char *name, *seq, *plus, *qual, *range;
  size_t name_len, seq_len, plus_len, qual_len, range_len;
  bool is_rna;
};
static bool next_read_fast(Reader *r, fasta_t fasta, Views *V, char c);

// ---- read-ahead ------------------------------------------------------------------------------------------------------
// gmapper.c calls the reader inside `omp critical (fill_reads_buffer)` (gmapper.c:338): whatever an entry costs there
// is serial time of the whole run -- at 150-200 ns per entry a ceiling of 5-6 M reads/s however many threads map.  Once
// a file has delivered READ_AHEAD_AFTER entries (a read file, not a genome of a few contigs) a thread of this file's
// own parses ahead -- the same next_read_fast() / next_read() -- and lays the TEXT of every entry into a byte ring; the
// call inside the critical section allocates the entry's strings from that text and nothing else.  (The strings are
// allocated by the thread that will free them: handing malloc'd blocks from one producer thread to sixteen consumers
// was measured three times slower than parsing in place -- every block and its allocator metadata change cores.)
// Order, strings, allocations (free()-able, one per string) and the end-of-file / error outcome are those of calling
// next_read() in place.
struct Ahead {
  struct Rec {   // followed by the text: name, sequence, '+' line, qualities, range field
    uint32_t total;   // bytes to the next record
    uint32_t name_len, seq_len, plus_len, qual_len, range_len;
    uint8_t ok, is_rna, has_name, has_seq, has_plus, has_qual, has_range, pad;
  };
  fasta_t owner = NULL;
  char *slab = NULL;
  size_t size = 0;
  // bytes written / taken since the start.  Each side works on its own copy and publishes it every PUBLISH bytes (and
  // whenever it has to wait): a counter the other side polls changes cores with every store otherwise
  alignas(64) std::atomic<unsigned long long> w{0};
  alignas(64) std::atomic<unsigned long long> r{0};
  alignas(64) unsigned long long w_local = 0, w_published = 0;   // producer
  alignas(64) unsigned long long r_local = 0, r_published = 0, w_seen = 0;   // consumer
  bool ended = false;   // consumer: the record that ended the file has been seen
  alignas(64) std::atomic<bool> stop{false};
  std::atomic<bool> done{false};
  std::thread th;
};
enum { PUBLISH = 32 << 10 };
static const size_t READ_AHEAD_AFTER =
    getenv("SHRIMP_B200_READ_AHEAD_AFTER") ? (size_t)atol(getenv("SHRIMP_B200_READ_AHEAD_AFTER")) : 4096;   // (tests: 50)
static Ahead *g_ahead[MAX_READERS];
static size_t g_delivered[MAX_READERS];

// room for a record of `need` bytes that does not wrap; NULL when asked to stop
static char *ahead_reserve(Ahead *A, size_t need, double *waited) {
  for (;;) {
    const unsigned long long w = A->w_local;
    const size_t at = (size_t)(w % A->size), to_end = A->size - at;
    const size_t want = need <= to_end ? need : to_end + need;   // a pad record to the end of the slab first
    while (w + want - A->r.load(std::memory_order_acquire) > A->size) {
      if (A->stop.load(std::memory_order_relaxed)) return NULL;
      if (A->w_published != w) {
        A->w.store(w, std::memory_order_release);
        A->w_published = w;
      }
      const double t0 = omp_get_wtime();
      usleep(200);   // the slab holds far more than is taken in 200 us
      *waited += omp_get_wtime() - t0;
    }
    if (need <= to_end) return A->slab + at;
    Ahead::Rec pad;
    memset(&pad, 0, sizeof(pad));
    pad.total = (uint32_t)to_end;
    pad.pad = 1;
    memcpy(A->slab + at, &pad, sizeof(pad));   // to_end >= sizeof(Rec): records are multiples of 32 bytes
    A->w_local = w + to_end;
  }
}

static void ahead_main(Ahead *A) {
  Reader *r = reader_of(A->owner);
  const bool fastq = A->owner->fastq;
  const char c = fastq ? '@' : '>';
  unsigned long long n_entries = 0, n_general = 0;
  const double t_start = omp_get_wtime();
  double t_waited = 0;
  for (;;) {
    Views V;
    read_entry tmp;
    memset(&tmp, 0, sizeof(tmp));
    bool ok = true, general = false;
    if (!next_read_fast(r, A->owner, &V, c)) {   // the general path makes strings: their text goes into the slab too
      general = true;
      ok = next_read(A->owner, &tmp);
      memset(&V, 0, sizeof(V));
      V.name = tmp.name; V.name_len = tmp.name ? strlen(tmp.name) : 0;
      V.seq = tmp.seq; V.seq_len = tmp.seq ? strlen(tmp.seq) : 0;
      V.plus = tmp.plus_line; V.plus_len = tmp.plus_line ? strlen(tmp.plus_line) : 0;
      V.qual = tmp.qual; V.qual_len = tmp.qual ? strlen(tmp.qual) : 0;
      V.range = tmp.range_string; V.range_len = tmp.range_string ? strlen(tmp.range_string) : 0;
      V.is_rna = tmp.is_rna;
    }
    const size_t text = V.name_len + V.seq_len + V.plus_len + V.qual_len + V.range_len;
    const size_t need = (sizeof(Ahead::Rec) + text + 31) & ~(size_t)31;
    if (need > A->size / 2) {
      fprintf(stderr, "error: an entry of %zu bytes in a read file is more than the read-ahead ring takes; "
                      "run without SHRIMP_B200_READ_AHEAD\n", text);
      exit(1);
    }
    n_entries++;
    n_general += general;
    char *at = ahead_reserve(A, need, &t_waited);
    if (at) {
      Ahead::Rec R;
      memset(&R, 0, sizeof(R));
      R.total = (uint32_t)need;
      R.name_len = (uint32_t)V.name_len; R.seq_len = (uint32_t)V.seq_len; R.plus_len = (uint32_t)V.plus_len;
      R.qual_len = (uint32_t)V.qual_len; R.range_len = (uint32_t)V.range_len;
      R.ok = ok; R.is_rna = V.is_rna;
      R.has_name = V.name != NULL; R.has_seq = V.seq != NULL; R.has_plus = V.plus != NULL; R.has_qual = V.qual != NULL;
      R.has_range = V.range != NULL;
      memcpy(at, &R, sizeof(R));
      char *p = at + sizeof(R);
      if (V.name_len) memcpy(p, V.name, V.name_len);
      p += V.name_len;
      if (V.seq_len) memcpy(p, V.seq, V.seq_len);
      p += V.seq_len;
      if (V.plus_len) memcpy(p, V.plus, V.plus_len);
      p += V.plus_len;
      if (V.qual_len) memcpy(p, V.qual, V.qual_len);
      p += V.qual_len;
      if (V.range_len) memcpy(p, V.range, V.range_len);
      A->w_local += need;
      if (general || !ok || A->w_local - A->w_published >= PUBLISH) {
        A->w.store(A->w_local, std::memory_order_release);
        A->w_published = A->w_local;
      }
    }
    if (general) {
      free(tmp.name);
      free(tmp.seq);
      free(tmp.plus_line);
      free(tmp.qual);
      free(tmp.range_string);
    }
    if (!at || !ok) break;
  }
  A->w.store(A->w_local, std::memory_order_release);
  A->done.store(true, std::memory_order_release);
  if (getenv("SHRIMP_B200_VERBOSE"))
    fprintf(stderr, "[gmapper-b200] read-ahead thread: %llu entries (%llu through the general path), %.3f s of its own\n",
            n_entries, n_general, omp_get_wtime() - t_start - t_waited);
}

static bool ahead_take(Ahead *A, read_entry *re) {
  re->name = re->seq = NULL;
  re->paired = false;
  re->first_in_pair = false;
  re->mate_pair = NULL;
  if (A->ended) return false;   // past the end of the file: fasta.c:340-372 finds no line
  for (;;) {
    const unsigned long long rd = A->r_local;
    int spins = 0;
    while (A->w_seen == rd) {
      A->w_seen = A->w.load(std::memory_order_acquire);
      if (A->w_seen != rd) break;
      if (A->r_published != rd) {   // about to wait: the producer may be waiting for this room
        A->r.store(rd, std::memory_order_release);
        A->r_published = rd;
      }
      if (A->done.load(std::memory_order_acquire) && A->w.load(std::memory_order_acquire) == rd) {
        A->ended = true;
        return false;
      }
      if (++spins < 64) sched_yield();
      else usleep(50);
    }
    const char *at = A->slab + (size_t)(rd % A->size);
    Ahead::Rec R;
    memcpy(&R, at, sizeof(R));
    if (R.pad) {
      A->r_local = rd + R.total;
      continue;
    }
    const char *p = at + sizeof(R);
    if (R.has_name) re->name = dup17(p, R.name_len);
    p += R.name_len;
    if (R.has_seq) {
      re->seq = dup17(p, R.seq_len);
      re->orig_seq = re->seq;
    }
    p += R.seq_len;
    if (R.has_plus) re->plus_line = dup17(p, R.plus_len);
    p += R.plus_len;
    if (R.has_qual) {   // (already raised to '!' when it came through the general path; idempotent)
      re->qual = (char *)xmalloc((size_t)R.qual_len + 17);
      for (uint32_t i = 0; i < R.qual_len; i++) re->qual[i] = MAX((char)p[i], '!');
      memset(re->qual + R.qual_len, 0, 17);
      re->orig_qual = re->qual;
    }
    p += R.qual_len;
    if (R.has_range) re->range_string = dup17(p, R.range_len);
    if (R.ok) re->is_rna = R.is_rna;
    A->r_local = rd + R.total;
    if (A->r_local - A->r_published >= PUBLISH) {
      A->r.store(A->r_local, std::memory_order_release);
      A->r_published = A->r_local;
    }
    if (!R.ok) A->ended = true;
    return R.ok;
  }
}

static int reader_slot(fasta_t f) {
  for (int i = 0; i < MAX_READERS; i++)
    if (g_readers[i].owner == f) return i;
  return -1;
}

bool fasta_get_next_read_with_range(fasta_t fasta, read_entry *re) {   // fasta.c:316
  // gmapper.c fills re_buffer[] entry after entry (gmapper.c:344-378): the entries behind this one are the next to be
  // written, and the chunk buffer (tens of megabytes, zeroed just before) is no longer in the near caches
  __builtin_prefetch((const char *)(re + 2), 1);
  __builtin_prefetch((const char *)(re + 2) + 192, 1);
  static const bool verbose = getenv("SHRIMP_B200_VERBOSE") != NULL;
  // measured on a 16-core B200 box (8 M C2 reads, -N 16): 4.4 M reads/s with the read-ahead thread, 4.5-4.8 M
  // without -- the parser thread delivers an entry every ~150 ns, which is what parsing in place costs, so the ring
  // only adds a hand-over.  It stays an option (SHRIMP_B200_READ_AHEAD=1) for hosts with slow cores and many of them.
  static const bool no_ahead = getenv("SHRIMP_B200_READ_AHEAD") == NULL || getenv("SHRIMP_B200_NO_READ_AHEAD") != NULL;
  unsigned long long t0 = 0;
  if (verbose) {
    t0 = __builtin_ia32_rdtsc();
    if (g_tick0 == 0) {
      g_tick0 = t0;
      g_wall0 = omp_get_wtime();
    }
  }
  bool ok;
  int slot = reader_slot(fasta);
  if (slot >= 0 && g_ahead[slot]) {
    ok = ahead_take(g_ahead[slot], re);
  } else {
    ok = next_read(fasta, re);
    if (slot < 0) slot = reader_slot(fasta);
    if (ok && slot >= 0 && !no_ahead && ++g_delivered[slot] == READ_AHEAD_AFTER) {
      Ahead *A = new Ahead;
      A->owner = fasta;
      A->size = (size_t)8 << 20;   // small enough to stay in the last-level cache between the two threads
      A->slab = (char *)xmalloc(A->size);
      g_ahead[slot] = A;
      A->th = std::thread(ahead_main, A);
    }
  }
  if (verbose) {
    g_reader_ticks += __builtin_ia32_rdtsc() - t0;
    g_reader_calls++;
  }
  return ok;
}

// The common shape of an entry -- a name line without tabs, ONE sequence line, (FASTQ: a '+' line and ONE quality line
// of the right length), the next entry's marker behind it, all inside the current block, no NUL bytes -- is taken
// apart in place with memchr; anything else returns false without having consumed a byte and next_read's general
// path (the restatement of fasta.c:316-545 piece by piece) takes the entry.
// One SSE2 pass over a line: where it ends, and whether a NUL, a tab or a uracil lies before its end; the first blank.
// (16-byte loads may run past `end` into the slack behind the block; a newline found there does not count.)
enum { LN_NUL = 1, LN_TAB = 2, LN_U = 4 };
static inline char *scan_line(char *p, char *end, unsigned *flags, char **first_blank) {
  const __m128i vnl = _mm_set1_epi8('\n'), vz = _mm_setzero_si128(), vtab = _mm_set1_epi8('\t'), vsp = _mm_set1_epi8(' '),
                vu = _mm_set1_epi8('u'), v20 = _mm_set1_epi8(0x20);
  unsigned fl = 0;
  char *fb = NULL;
  for (; p < end; p += 16) {
    const __m128i v = _mm_loadu_si128((const __m128i *)p);
    const unsigned mnl = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, vnl));
    const unsigned before = mnl ? ((mnl & (0u - mnl)) - 1u) : 0xffffu;   // the bytes in front of the first newline
    const unsigned mz = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, vz)) & before;
    const unsigned mt = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, vtab)) & before;
    const unsigned ms = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, vsp)) & before;
    const unsigned mu = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_or_si128(v, v20), vu)) & before;
    fl |= (mz ? LN_NUL : 0) | (mt ? LN_TAB : 0) | (mu ? LN_U : 0);
    if (ms && !fb) fb = p + __builtin_ctz(ms);
    if (mnl) {
      char *at = p + __builtin_ctz(mnl);
      if (at >= end) return NULL;
      *flags = fl;
      *first_blank = fb;
      return at;
    }
  }
  return NULL;
}

static bool next_read_fast(Reader *r, fasta_t fasta, Views *V, char c) {
  if (r->header) return false;
  char *s = r->buf + r->pos, *const end = r->buf + r->end;
  if (s >= end || *s != c) return false;
  unsigned nfl = 0, sfl = 0, xfl = 0;
  char *nblank = NULL, *xblank = NULL;
  char *e = scan_line(s, end, &nfl, &nblank);
  if (!e || (nfl & (LN_NUL | LN_TAB))) return false;
  char *q = e + 1;                                  // the sequence line
  if (q >= end || *q == '#' || *q == '>' || *q == '+') return false;
  char *qe = scan_line(q, end, &sfl, &xblank);
  if (!qe || qe == q || (sfl & LN_NUL)) return false;   // (a NUL ends the reference's copy: general path)
  char *nx = qe + 1;                                // what follows the sequence line
  char *plus = NULL, *pe = NULL, *ql = NULL, *qle = NULL;
  if (!fasta->fastq) {
    if (nx >= end || *nx != '>') return false;      // more sequence lines, a comment, or the end of the block / file
  } else {
    if (nx >= end || *nx != '+') return false;
    plus = nx;
    pe = scan_line(plus, end, &xfl, &xblank);
    if (!pe || (xfl & LN_NUL)) return false;
    ql = pe + 1;
    if (ql >= end) return false;
    qle = scan_line(ql, end, &xfl, &xblank);
    if (!qle || (xfl & LN_NUL)) return false;
    const size_t want = fasta->space == LETTER_SPACE ? (size_t)(qe - q) : (size_t)(qe - q) - 1;
    if ((size_t)(qle - ql) != want || want == 0) return false;
    nx = qle + 1;
  }
  // the name: trimmed, cut at its first blank (extract_name, fasta.c:243-281, without a range field)
  char *b = s + 1, *t = e;
  while (b < t && isspace((unsigned char)*b)) b++;
  while (t > b && isspace((unsigned char)t[-1])) t--;
  if (b == t) return false;
  char *cut = nblank;
  if (cut && cut < b) cut = (char *)memchr(b, ' ', (size_t)(t - b));   // (blanks in front of the name)
  const size_t seq_len = (size_t)(qe - q);
  // uracil without thymine marks RNA (fasta.c:524-538)
  const bool ur = (sfl & LN_U) != 0;
  const bool th = ur && (memchr(q, 'T', seq_len) || memchr(q, 't', seq_len));
  V->name = b;
  V->name_len = (size_t)((cut && cut < t ? cut : t) - b);
  V->seq = q;
  V->seq_len = seq_len;
  V->plus = plus;
  V->plus_len = plus ? (size_t)(pe - plus) : 0;
  V->qual = ql;
  V->qual_len = ql ? (size_t)(qle - ql) : 0;
  V->range = NULL;
  V->range_len = 0;
  if (ur && th) fprintf(stderr, "WARNING: sequence has both uracil and thymine!?!\n");
  V->is_rna = ur && !th;
  r->pos = (size_t)(nx - r->buf);
  return true;
}

// the strings of an entry as fasta.c leaves them in a read_entry: one allocation each, 17 zero bytes behind the text,
// qualities raised to '!' (fasta.c:381-383, :508-513)
static void materialise(const Views &V, bool fastq, read_entry *re) {
  re->name = dup17(V.name, V.name_len);
  re->seq = dup17(V.seq, V.seq_len);
  re->orig_seq = re->seq;
  if (V.range) re->range_string = dup17(V.range, V.range_len);
  if (fastq) {
    re->plus_line = dup17(V.plus, V.plus_len);
    re->qual = (char *)xmalloc(V.qual_len + 17);
    for (size_t i = 0; i < V.qual_len; i++) re->qual[i] = MAX((char)V.qual[i], '!');
    memset(re->qual + V.qual_len, 0, 17);
    re->orig_qual = re->qual;
  }
  re->is_rna = V.is_rna;
}


static bool next_read(fasta_t fasta, read_entry *re) {
  Reader *r = reader_of(fasta);
  const char c = fasta->fastq ? '@' : '>';
  re->name = re->seq = NULL;
  re->paired = false;
  re->first_in_pair = false;
  re->mate_pair = NULL;
  Views V;
  if (next_read_fast(r, fasta, &V, c)) {
    materialise(V, fasta->fastq, re);
    return true;
  }

  // ---- the name line (fasta.c:340-384); comment lines ('#') in front of it are skipped
  size_t name_len = 0;
  bool end_of_line = false;
  size_t len;
  bool nl;
  const char *s;
  while ((s = next_piece(r, &len, &nl)) != NULL) {
    const size_t tl = text_len(s, len);
    acc_reserve(r, name_len + tl + 1);
    memcpy(r->acc + name_len, s, tl);
    name_len += tl;
    const char first = name_len ? r->acc[0] : '\0';
    if (first != '#' && first != c) {
      if (c == '>' && first == '@')
        fprintf(stderr, "Expecting \">\" but got \"%c\" are you sure it's not FASTQ format?\n", first);
      else if (c == '@' && first == '>')
        fprintf(stderr, "Expecting \"@\" but got \"%c\" are you sure it's not FASTA format?\n", first);
      else
        fprintf(stderr, "Expecting \"%c\" but got \"%c\" are you sure it's right format?\n", c, first);
      return false;
    }
    if (nl) {
      if (s[0] == '#' && len > 0) {   // fasta.c:366-370 tests the piece, not the assembled line
        name_len = 0;
        continue;
      }
      end_of_line = true;
      r->acc[name_len] = '\0';
      re->name = extract_name_fast(r->acc, &re->range_string);
      break;
    }
  }
  if (name_len == 0) return false;
  if (name_len <= 1 || !end_of_line) {
    r->acc[name_len] = '\0';
    fprintf(stderr, "error: Invalid read name! Are you sure this is a FASTA or FASTQ file?\n%s\n", r->acc);
    return false;
  }

  // ---- the sequence (fasta.c:389-417): every line up to the next '>' (FASTA) or the '+' line (FASTQ)
  size_t seq_len = 0;
  bool at_line_start = true;
  bool plus_seen = false;
  for (;;) {
    if (at_line_start) {
      const int f = peek_first(r);
      if (f < 0) break;
      if (fasta->fastq && f == '+') {
        plus_seen = true;
        break;
      }
      if (!fasta->fastq && f == '>') break;   // stays unread: the reference's `leftover`
    }
    s = next_piece(r, &len, &nl);
    if (s == NULL) break;
    const bool comment = at_line_start && len > 0 && s[0] == '#';
    // (a continuation piece of an over-long line is appended whatever it starts with)
    at_line_start = nl;
    if (comment) continue;
    const size_t tl = text_len(s, len);
    acc_reserve(r, seq_len + tl + 1);
    memcpy(r->acc + seq_len, s, tl);
    seq_len += tl;
  }
  if (seq_len == 0) {
    fprintf(stderr, "Read in sequence of length zero!\n");
    return false;
  }
  re->seq = dup17(r->acc, seq_len);
  re->orig_seq = re->seq;

  if (fasta->fastq) {
    // ---- the '+' line (fasta.c:418-447)
    size_t plus_len = 0;
    bool plus_nl = false;
    if (plus_seen) {
      while ((s = next_piece(r, &len, &nl)) != NULL) {
        const size_t tl = text_len(s, len);
        acc_reserve(r, plus_len + tl + 1);
        memcpy(r->acc + plus_len, s, tl);
        plus_len += tl;
        if (nl) {
          plus_nl = true;
          break;
        }
      }
    }
    if (plus_len < 1 || !plus_nl) {
      fprintf(stderr, "error: Error while readingin FASTQ entry!\n");
      return false;
    }
    re->plus_line = dup17(r->acc, plus_len);

    // ---- the qualities (fasta.c:452-506): lines until there are as many as bases (colour space: one less)
    const size_t want = fasta->space == LETTER_SPACE ? seq_len : seq_len - 1;
    size_t qual_len = 0;
    while ((s = next_piece(r, &len, &nl)) != NULL) {
      const size_t tl = text_len(s, len);
      acc_reserve(r, qual_len + tl + 1);
      memcpy(r->acc + qual_len, s, tl);
      qual_len += tl;
      if (qual_len == want) break;
      if (qual_len > seq_len) {
        fprintf(stderr, "There has been a problem reading in the read \"%s\", the quality length exceeds the sequence length!\n",
                re->name);
        fprintf(stderr, "Are you using the right executable? gmapper-cs for color space? and gmapper-ls for letter space?\n");
        exit(1);
      }
    }
    if (qual_len != want) {
      fprintf(stderr, "Read in quality string of wrong length!, %d vs %d\n", (int)qual_len, (int)seq_len);
      free(re->seq);
      free(re->plus_line);
      re->seq = re->orig_seq = re->plus_line = NULL;   // (the reference leaves them dangling; its caller drops the entry)
      return false;
    }
    re->qual = (char *)xmalloc(qual_len + 17);
    for (size_t i = 0; i < qual_len; i++) re->qual[i] = MAX((char)r->acc[i], '!');
    memset(re->qual + qual_len, 0, 17);
    re->orig_qual = re->qual;
  }

  // ---- RNA? (fasta.c:524-538)
  bool got_uracil = false, got_thymine = false;
  for (size_t j = 0; j < seq_len; j++) {
    const unsigned char chr = (unsigned char)re->seq[j];
    got_thymine |= (chr == 'T' || chr == 't');
    got_uracil |= (chr == 'U' || chr == 'u');
  }
  if (got_uracil && got_thymine) fprintf(stderr, "WARNING: sequence has both uracil and thymine!?!\n");
  re->is_rna = (got_uracil && !got_thymine);
  return true;
}

void fasta_close(fasta_t fasta) {   // fasta.c:208-220, plus this file's block
  if (getenv("SHRIMP_B200_VERBOSE") && g_reader_calls > 1000) {
    const double tps = (double)(__builtin_ia32_rdtsc() - g_tick0) / (omp_get_wtime() - g_wall0);
    fprintf(stderr, "[gmapper-b200] reader: %llu entries in %.3f s (%.0f ns each), inside the critical section of gmapper.c:338\n",
            g_reader_calls, (double)g_reader_ticks / tps, 1e9 * (double)g_reader_ticks / tps / (double)g_reader_calls);
    g_reader_calls = 0;
  }
  for (int i = 0; i < MAX_READERS; i++)
    if (g_readers[i].owner == fasta) {
      if (Ahead *A = g_ahead[i]) {   // stop the read-ahead; entries nobody took are released
        A->stop.store(true);
        A->th.join();
        free(A->slab);
        delete A;
        g_ahead[i] = NULL;
      }
      g_delivered[i] = 0;
      free(g_readers[i].buf);
      free(g_readers[i].acc);
      g_readers[i] = Reader();
    }
  gzclose(fasta->fp);
  free(fasta->file);
  free(fasta->parse_buffer);
  if (fasta->save_buf != NULL) free(fasta->save_buf);
  free(fasta);
}

// ================================================================================================================
// Packing a read (gmapper.c:475-488)
// ================================================================================================================
uint32_t *fasta_sequence_to_bitfield(fasta_t fasta, char *sequence) {   // fasta.c:610
  const uint32_t length = (uint32_t)strlen(sequence);
  if (length == 0) return NULL;
  const size_t words = BPTO32BW(length);
  uint32_t *bitfield = (uint32_t *)xmalloc(words * sizeof(uint32_t));
  memset(bitfield, 0, words * sizeof(uint32_t));
  uint32_t i = 0;
  if (fasta->space == COLOUR_SPACE) {   // the initial base is not part of the bitfield
    const char c = sequence[0];
    if (c != 'A' && c != 'a' && c != 'C' && c != 'c' && c != 'G' && c != 'g' && c != 'T' && c != 't') {
      free(bitfield);
      return NULL;
    }
    i = 1;
  }
  for (uint32_t idx = 0; i < length; i++, idx++) {
    const unsigned char ch = (unsigned char)sequence[i];
    const int a = ch < 128 ? fasta->translate[ch] : -1;
    if (a == -1) {   // the reference's message (it tests the translated value, so no character is printed)
      fprintf(stderr, "error: invalid character ");
      fprintf(stderr, "in input file [%s]\n", fasta->file);
      fprintf(stderr, "       (Did you mix up letter space and colour space programs?)\n");
      exit(1);
    }
    bitfield[idx >> 3] |= ((uint32_t)a & 0xfu) << (4 * (idx & 7));
  }
  return bitfield;
}

uint32_t *reverse_complement_read_cs(uint32_t *read, int8_t initbp, int8_t initbp_rc, uint32_t len, bool is_rna) {   // util.c:601
  const size_t words = BPTO32BW(len);
  uint32_t *read_rc = (uint32_t *)xmalloc(sizeof(uint32_t) * words);
  memset(read_rc, 0, sizeof(uint32_t) * words);   // (the reference leaves the unused nibbles of the last word unset)
  int8_t base = (int8_t)cstols(initbp, EXTRACT(read, 0), is_rna);
  for (uint32_t i = 1; i < len; i++) {
    const uint32_t c = EXTRACT(read, i);
    base = (int8_t)cstols(base, (int)c, is_rna);
    const uint32_t at = len - i;
    read_rc[at >> 3] |= c << (4 * (at & 7));
  }
  read_rc[0] |= (uint32_t)lstocs(base, complement_base(initbp_rc, is_rna), is_rna) & 0xfu;
  return read_rc;
}

#ifndef FAST_IO_READER_ONLY   // (tools: a reader-only build for the parser micro-benchmark)
// ================================================================================================================
// Formatter
// ================================================================================================================
namespace {

inline char *put_str(char *p, const char *s) {
  const size_t n = strlen(s);
  memcpy(p, s, n);
  return p + n;
}
inline char *put_mem(char *p, const char *s, size_t n) {
  memcpy(p, s, n);
  return p + n;
}
inline char *put_u64(char *p, unsigned long long v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}
inline char *put_int(char *p, int v) {   // %i / %d
  if (v < 0) {
    *p++ = '-';
    return put_u64(p, (unsigned long long)(-(long long)v));
  }
  return put_u64(p, (unsigned long long)v);
}
inline char *put_uint(char *p, int v) { return put_u64(p, (unsigned long long)(unsigned int)v); }   // %u of an int

// reverse(), output.c:167-220: reverse complement that keeps the case; 0 = a character it stops the run on
unsigned char g_rc[256];
// SEQ of an unmapped / clipped letter-space base, output.c:314-336: the ten two- and three-fold IUPAC codes become N
unsigned char g_ls_seq[256];
bool g_tables_ready = false;
void init_tables() {
  memset(g_rc, 0, sizeof(g_rc));
  const char *a = "ATCGRYKMBVDH", *b = "TAGCYRMKVBHD";
  for (int i = 0; a[i]; i++) {
    g_rc[(int)a[i]] = (unsigned char)b[i];
    g_rc[(int)a[i] + 32] = (unsigned char)(b[i] + 32);
  }
  for (const char *c = "-.NnSsWw"; *c; c++) g_rc[(int)*c] = (unsigned char)*c;
  for (int i = 0; i < 256; i++) {
    char c = (char)i;
    if (strchr("RYSWKMBDHV", c) && c) {
      g_ls_seq[i] = 'N';
    } else {
      if (c >= 'a') c -= 32;
      g_ls_seq[i] = (unsigned char)c;
    }
  }
  __atomic_store_n(&g_tables_ready, true, __ATOMIC_RELEASE);
}

}  // namespace

void hit_output(struct read_entry *re, struct read_hit *rh, struct read_hit *rh_mp, bool first_in_pair, int *hits,
                int satisfying_alignments, bool improper_mapping) {   // output.c:227
  // what this file does not restate goes to the reference's own code
  if (!Eflag || extra_sam_fields || (shrimp_mode == MODE_COLOUR_SPACE && Qflag && Bflag)) {
    shrimp_ref_hit_output(re, rh, rh_mp, first_in_pair, hits, satisfying_alignments, improper_mapping);
    return;
  }
  if (!__atomic_load_n(&g_tables_ready, __ATOMIC_ACQUIRE)) {
#pragma omp critical(shrimp_fast_io_tables)
    {
      if (!g_tables_ready) init_tables();
    }
  }
  const bool cs = shrimp_mode == MODE_COLOUR_SPACE;
  const size_t name_len = strlen(re->name);
  const size_t seq_str_len = strlen(re->seq);
  const size_t qual_str_len = (Qflag && re->qual) ? strlen(re->qual) : 0;
  struct read_entry *re_mp = re->mate_pair;
  const bool paired_read = re->paired;
  const size_t mp_seq_len = (sam_r2 && re_mp && re_mp->seq) ? strlen(re_mp->seq) : 0;
  const size_t rg_len = sam_read_group_name ? strlen(sam_read_group_name) : 0;
  const struct sw_full_results *sfr = rh ? rh->sfrp : NULL;
  const size_t al_len = sfr && sfr->qralign ? strlen(sfr->qralign) : 0;

  // an upper bound of the record; the reference's snprintf calls would cut it at the end of the buffer
  // (thread_output_buffer_safety = 500 KB is guaranteed below), so anything near that goes to the reference's code
  const size_t bound = name_len + 3 * seq_str_len + 3 * qual_str_len + mp_seq_len + rg_len + 16 * al_len + 1024;
  if (bound >= thread_output_buffer_safety || (sfr && (!sfr->qralign || !sfr->dbalign)) ||
      (paired_read && re_mp == NULL) || (sam_r2 && (re_mp == NULL || re_mp->seq == NULL))) {
    shrimp_ref_hit_output(re, rh, rh_mp, first_in_pair, hits, satisfying_alignments, improper_mapping);
    return;
  }

  // output.c:246-268: room for one record
  const int thread_id = omp_get_thread_num();
  while ((size_t)(thread_output_buffer[thread_id] + thread_output_buffer_sizes[thread_id] -
                  thread_output_buffer_filled[thread_id]) < thread_output_buffer_safety) {
    const size_t new_size = thread_output_buffer_sizes[thread_id] + thread_output_buffer_increment;
    const size_t filled = thread_output_buffer_filled[thread_id] - thread_output_buffer[thread_id];
    thread_output_buffer[thread_id] =
        (char *)my_realloc(thread_output_buffer[thread_id], new_size, thread_output_buffer_sizes[thread_id],
                           &mem_thread_buffer, "realloc thread_output_buffer");
    thread_output_buffer_sizes[thread_id] = new_size;
    thread_output_buffer_filled[thread_id] = thread_output_buffer[thread_id] + filled;
  }
  char *const rec = thread_output_buffer_filled[thread_id];
  char *p = rec;

  // ---- QNAME (output.c:363-378): for a paired read the common prefix of the two names without a trailing ':' / '/'
  size_t qname_len = name_len;
  bool mate_unmapped = false, reverse_strand_mp = false;
  int genome_start_mp = 0, genome_end_mp = 0, mpos = 0;
  const char *mrnm = "*";
  if (paired_read) {
    const size_t mp_name_len = strlen(re_mp->name);
    const size_t m = MIN(name_len, mp_name_len);
    size_t i = 0;
    while (i < m && re->name[i] == re_mp->name[i]) i++;
    // (the reference copies the whole name first and then cuts it at the first difference, or at the shorter length)
    if (i > 0 && (re->name[i - 1] == ':' || re->name[i - 1] == '/')) i--;
    qname_len = i;
    mate_unmapped = (rh_mp == NULL);
    if (!mate_unmapped) {   // output.c:381-398
      const struct sw_full_results *m_sfr = rh_mp->sfrp;
      const int read_start_mp = m_sfr->read_start + 1;
      const int read_end_mp = read_start_mp + m_sfr->rmapped - 1;
      const int genome_length_mp = genome_len[rh_mp->cn];
      reverse_strand_mp = (rh_mp->gen_st == 1);
      if (!reverse_strand_mp)
        genome_start_mp = m_sfr->genome_start + 1;
      else
        genome_start_mp = (genome_length_mp - m_sfr->genome_start) -
                          (read_end_mp - read_start_mp - m_sfr->deletions + m_sfr->insertions);
      genome_end_mp = genome_start_mp + m_sfr->gmapped - 1;
      mpos = genome_start_mp;
      mrnm = contig_names[rh_mp->cn];
    }
  }
  const bool second_in_pair = paired_read && !first_in_pair;
  const bool paired_alignment = paired_read && (rh != NULL && rh_mp != NULL && !improper_mapping);
  const bool query_unmapped = (rh == NULL);

  if (query_unmapped || (!half_paired && paired_read && mate_unmapped)) {   // output.c:409-468
    const int flag = (paired_read ? 0x0001 : 0) | (paired_alignment ? 0x0002 : 0) | (query_unmapped ? 0x0004 : 0) |
                     (mate_unmapped ? 0x0008 : 0) | (reverse_strand_mp ? 0x0020 : 0) | (first_in_pair ? 0x0040 : 0) |
                     (second_in_pair ? 0x0080 : 0);
    p = put_mem(p, re->name, qname_len);
    *p++ = '\t';
    p = put_int(p, flag);
    p = put_str(p, "\t*\t0\t0\t*\t");
    p = put_str(p, mrnm);
    *p++ = '\t';
    p = put_uint(p, mpos);
    p = put_str(p, "\t0\t");
    if (!cs) {
      for (int i = 0; i < re->read_len; i++) *p++ = (char)g_ls_seq[(unsigned char)re->seq[i]];
    } else {
      *p++ = '*';
    }
    *p++ = '\t';
    if (Qflag && !cs)
      p = put_mem(p, re->qual, qual_str_len);
    else
      *p++ = '*';
    if (cs) {
      p = put_str(p, "\tCQ:Z:");
      if (Qflag)
        p = put_mem(p, re->qual, qual_str_len);
      else
        *p++ = '*';
      p = put_str(p, "\tCS:Z:");
      p = put_mem(p, re->seq, seq_str_len);
    }
    if (sam_r2) {
      p = put_str(p, cs ? "\tX2:Z:" : "\tR2:Z:");
      p = put_mem(p, re_mp->seq, mp_seq_len);
    }
    if (sam_read_group_name != NULL) {
      p = put_str(p, "\tRG:Z:");
      p = put_mem(p, sam_read_group_name, rg_len);
    }
    *p++ = '\n';
    *p = '\0';
    thread_output_buffer_filled[thread_id] = p;
    return;
  }

  // ---- a mapped read (output.c:470-774)
  const char *rname = contig_names[rh->cn];
  const bool reverse_strand = (rh->gen_st == 1);
  const int read_length = re->read_len;
  const int read_start = sfr->read_start + 1;
  const int read_end = read_start + sfr->rmapped - 1;
  const int genome_length = genome_len[rh->cn];
  const char *qr = sfr->qralign, *db = sfr->dbalign;
  const int qralign_length = (int)al_len;

  // CIGAR runs (make_cigar, output.c:15-62): S, then D / I / M runs of the alignment strings, then S
  // (colour space: H, output.c:575-579); written later, reversed for a reverse-strand hit (output.c:629)
  const int max_ops = qralign_length + 2;
  char ops_small[256];
  uint32_t lens_small[256];
  char *ops = ops_small;
  uint32_t *lens = lens_small;
  if (max_ops > 256) {
    ops = (char *)xmalloc((size_t)max_ops);
    lens = (uint32_t *)xmalloc((size_t)max_ops * sizeof(uint32_t));
  }
  int n_ops = 0;
  const char clip = cs ? 'H' : 'S';
  if (read_start > 1) {
    ops[n_ops] = clip;
    lens[n_ops++] = (uint32_t)(read_start - 1);
  }
  for (int i = 0; i < qralign_length;) {
    int length;
    char op;
    if (qr[i] == '-') {
      for (length = 0; i + length < qralign_length && qr[i + length] == '-'; length++) {}
      op = 'D';
    } else if (db[i] == '-') {
      for (length = 0; i + length < qralign_length && db[i + length] == '-'; length++) {}
      op = 'I';
    } else {
      for (length = 0; i + length < qralign_length && db[i + length] != '-' && qr[i + length] != '-'; length++) {}
      op = 'M';
    }
    ops[n_ops] = op;
    lens[n_ops++] = (uint32_t)length;
    i += length;
  }
  if (read_end != read_length) {
    ops[n_ops] = clip;
    lens[n_ops++] = (uint32_t)(read_length - read_end);
  }

  // ---- SEQ (output.c:308-341, :484-536): letter space starts from the read's own letters and overlays the aligned
  // part; colour space prints the aligned part (the corrected base calls) only
  char seq_small[1280];
  const size_t seq_cap = (size_t)read_length + (size_t)qralign_length + 2;
  char *seq = seq_cap <= sizeof(seq_small) ? seq_small : (char *)xmalloc(seq_cap);
  int j = 0;
  if (!cs) {
    for (int i = 0; i < read_length; i++) seq[i] = (char)g_ls_seq[(unsigned char)re->seq[i]];
    seq[read_length] = '\0';
    j = read_start - 1;
  }
  for (int i = 0; i < qralign_length; i++) {
    char c = qr[i];
    if (c == '-') continue;
    if (c >= 'a') c -= 32;
    if (c != 'A' && c != 'G' && c != 'C' && c != 'T' && c != 'N') {
      // output.c:503-531: any other code is printed as N (the table that would resolve it against the genome
      // letter tests `c` after it was set to 'N'); a genome letter outside ACGT is reported on stderr
      c = 'N';
      if (db[i] != '-') {
        char g = db[i];
        if (g >= 'a') g -= 32;
        if (g != 'A' && g != 'C' && g != 'G' && g != 'T')
          fprintf(stderr, "There has been an error in printing an alignment, %c\n", g);
      }
    }
    seq[j++] = c;
  }
  int seq_len;
  bool bail = false;
  if (cs && j != read_end - read_start + 1) bail = true;   // output.c:541-543 would print what the stack held
  if (!cs) {
    seq_len = j + (read_length - read_end);
    seq[seq_len] = '\0';   // output.c:539 cuts the string here
    seq_len = (int)strlen(seq);
  } else {
    seq_len = j;
    seq[seq_len] = '\0';
  }

  // ---- QUAL (output.c:546-613)
  char qual_small[1280];
  const size_t qual_cap = MAX(qual_str_len, (size_t)read_length) + 16 + (sfr->qual ? strlen(sfr->qual) : 0);
  char *qual = qual_cap <= sizeof(qual_small) ? qual_small : (char *)xmalloc(qual_cap);
  size_t qual_len = 1;
  qual[0] = '*';
  qual[1] = '\0';
  if (!cs) {
    if (Qflag) {
      const int ql = (int)qual_str_len;
      if (!reverse_strand)
        memcpy(qual, re->qual, (size_t)ql + 1);
      else {
        for (int i = 0; i < ql; i++) qual[(ql - 1) - i] = re->qual[i];
        qual[ql] = '\0';
      }
      if (qual_delta != 33)
        for (int i = 0; i < ql; i++) qual[i] = (char)(qual[i] - qual_delta + 33);
      qual_len = strlen(qual);   // a shifted quality can become NUL; snprintf("%s") stops there
    }
  } else if (Qflag && compute_mapping_qualities) {
    if (sfr->qual == NULL) {
      bail = true;
    } else {
      strcpy(qual, sfr->qual);
      if (reverse_strand)
        for (int i = 0; i < sfr->rmapped / 2; i++) {
          const char t = qual[i];
          qual[i] = qual[sfr->rmapped - i - 1];
          qual[sfr->rmapped - i - 1] = t;
        }
      qual_len = strlen(qual);
    }
  }

  // ---- POS and strand (output.c:615-630)
  int genome_start;
  if (!reverse_strand) {
    genome_start = sfr->genome_start + 1;
  } else {
    genome_start = (genome_length - sfr->genome_start) - (read_end - read_start - sfr->deletions + sfr->insertions);
    for (int i = 0; i < seq_len && !bail; i++)
      if (g_rc[(unsigned char)seq[i]] == 0) bail = true;   // reverse() stops the run on such a letter: let it
    if (!bail) {
      for (int a = 0, b = seq_len - 1; a < b; a++, b--) {
        const char t = (char)g_rc[(unsigned char)seq[a]];
        seq[a] = (char)g_rc[(unsigned char)seq[b]];
        seq[b] = t;
      }
      if (seq_len & 1) seq[seq_len / 2] = (char)g_rc[(unsigned char)seq[seq_len / 2]];
    }
  }
  if (bail) {
    if (ops != ops_small) {
      free(ops);
      free(lens);
    }
    if (seq != seq_small) free(seq);
    if (qual != qual_small) free(qual);
    shrimp_ref_hit_output(re, rh, rh_mp, first_in_pair, hits, satisfying_alignments, improper_mapping);
    return;
  }
  const int genome_end = genome_start + sfr->gmapped - 1;

  // ---- mate columns (output.c:637-659)
  int isize = 0;
  if (paired_read && !mate_unmapped) {
    if (strcmp(rname, mrnm) == 0) {
      mrnm = "=";
      const int fivep = reverse_strand ? genome_end : genome_start - 1;
      const int fivep_mp = reverse_strand_mp ? genome_end_mp : genome_start_mp - 1;
      isize = fivep_mp - fivep;
    }
  }
  const int flag = (paired_read ? 0x0001 : 0) | (paired_alignment ? 0x0002 : 0) | (mate_unmapped ? 0x0008 : 0) |
                   (reverse_strand ? 0x0010 : 0) | (reverse_strand_mp ? 0x0020 : 0) | (first_in_pair ? 0x0040 : 0) |
                   (second_in_pair ? 0x0080 : 0);

  // ---- the record (output.c:676-760)
  p = put_mem(p, re->name, qname_len);
  *p++ = '\t';
  p = put_int(p, flag);
  *p++ = '\t';
  p = put_str(p, rname);
  *p++ = '\t';
  p = put_uint(p, genome_start);
  *p++ = '\t';
  p = put_int(p, sfr->mqv);
  *p++ = '\t';
  if (!reverse_strand)
    for (int i = 0; i < n_ops; i++) {
      p = put_int(p, (int)lens[i]);
      *p++ = ops[i];
    }
  else
    for (int i = n_ops - 1; i >= 0; i--) {
      p = put_int(p, (int)lens[i]);
      *p++ = ops[i];
    }
  *p++ = '\t';
  p = put_str(p, mrnm);
  *p++ = '\t';
  p = put_uint(p, mpos);
  *p++ = '\t';
  p = put_int(p, isize);
  *p++ = '\t';
  p = put_mem(p, seq, (size_t)seq_len);
  *p++ = '\t';
  p = put_mem(p, qual, qual_len);
  p = put_str(p, "\tAS:i:");
  p = put_int(p, rh->score_full);
  if (compute_mapping_qualities && !all_contigs) {
    if (pair_mode == PAIR_NONE) {
      p = put_str(p, "\tZ0:i:");
      p = put_int(p, double_to_neglog(sfr->z0));
      p = put_str(p, "\tZ1:i:");
      p = put_int(p, double_to_neglog(sfr->z1));
    } else if (rh != NULL && rh_mp != NULL && !improper_mapping) {
      p = put_str(p, "\tZ2:i:");
      p = put_int(p, double_to_neglog(sfr->z2));
      p = put_str(p, "\tZ3:i:");
      p = put_int(p, double_to_neglog(sfr->z3));
      p = put_str(p, "\tZ4:i:");
      p = put_int(p, double_to_neglog(sfr->pr_top_random_at_location));
      p = put_str(p, "\tZ6:i:");
      p = put_int(p, double_to_neglog(sfr->insert_size_denom));
    } else {
      p = put_str(p, "\tZ0:i:");
      p = put_int(p, double_to_neglog(sfr->z0));
      p = put_str(p, "\tZ1:i:");
      p = put_int(p, double_to_neglog(sfr->z1));
      p = put_str(p, "\tZ4:i:");
      p = put_int(p, double_to_neglog(sfr->pr_top_random_at_location));
      p = put_str(p, "\tZ5:i:");
      p = put_int(p, double_to_neglog(sfr->pr_missed_mp));
    }
  }
  p = put_str(p, "\tNM:i:");
  p = put_int(p, sfr->mismatches + sfr->deletions + sfr->insertions);
  if (cs) {
    if (Qflag) {
      p = put_str(p, "\tCQ:Z:");
      p = put_mem(p, re->qual, qual_str_len);
    }
    p = put_str(p, "\tCS:Z:");
    p = put_mem(p, re->seq, seq_str_len);
    p = put_str(p, "\tCM:i:");
    p = put_int(p, sfr->crossovers);
    p = put_str(p, "\tXX:Z:");
    p = put_mem(p, qr, (size_t)qralign_length);
  }
  if (sam_r2) {
    p = put_str(p, cs ? "\tX2:Z:" : "\tR2:Z:");
    p = put_mem(p, re_mp->seq, mp_seq_len);
  }
  if (sam_read_group_name != NULL) {
    p = put_str(p, "\tRG:Z:");
    p = put_mem(p, sam_read_group_name, rg_len);
  }
  *p++ = '\n';
  *p = '\0';
  thread_output_buffer_filled[thread_id] = p;
  if (ops != ops_small) {
    free(ops);
    free(lens);
  }
  if (seq != seq_small) free(seq);
  if (qual != qual_small) free(qual);
}
#endif   // FAST_IO_READER_ONLY
