// Micro-benchmark of the FASTA/FASTQ reader of integration/fast_io.cpp against the reference's (common/fasta.c:316):
// entries per second of fasta_get_next_read_with_range() alone, strings freed as gmapper.c's read_free does.
//
//   cd integration && F="-O2 -DNDEBUG -fopenmp -std=gnu++17 -w -D__STDC_FORMAT_MACROS -D__STDC_LIMIT_MACROS -I/root/reference -I../include -I."
//   g++ $F -D_MODULE_GMAPPER -c -o /tmp/rb.o ../tools/reader_bench.cpp
//   g++ $F -DFAST_IO_READER_ONLY -c -o /tmp/fio.o fast_io.cpp
//   g++ -fopenmp -pthread -o /tmp/rb_fast /tmp/rb.o /tmp/fio.o _build/weak/common_fasta.o _build/weak/common_util.o _build/ref/common_my-alloc.o -lz -lm
//   g++ -fopenmp -o /tmp/rb_ref /tmp/rb.o _build/ref/common_fasta.o _build/ref/common_util.o _build/ref/common_my-alloc.o -lz -lm
//   /tmp/rb_fast reads.csfa; /tmp/rb_ref reads.csfa        (SHRIMP_B200_READ_AHEAD=1: with the read-ahead thread)
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "gmapper/gmapper.h"
#include "common/fasta.h"
static double now() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}
int main(int argc, char **argv) {
  if (argc < 2) return 1;
  bool fq = false;
  fasta_t f = fasta_open(argv[1], argc > 2 && !strcmp(argv[2], "ls") ? LETTER_SPACE : COLOUR_SPACE, false, &fq);
  if (!f) return 1;
  read_entry re;
  long n = 0, bases = 0;
  const double t0 = now();
  for (;;) {
    memset(&re, 0, sizeof(re));
    if (!fasta_get_next_read_with_range(f, &re)) break;
    n++;
    bases += (long)strlen(re.seq);
    free(re.name);
    free(re.seq);
    if (fq) {
      free(re.qual);
      free(re.plus_line);
    }
  }
  const double t1 = now();
  printf("%ld entries, %ld bases, %.3f s, %.0f ns per entry\n", n, bases, t1 - t0, 1e9 * (t1 - t0) / (double)n);
  fasta_close(f);
  return 0;
}
