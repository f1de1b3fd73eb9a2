/* dbalign / qralign of struct sw_full_results from the edit script libshrimp_b200.so returns: what pretty_print
 * (common/sw-full-ls.c:524-560, common/sw-full-cs.c:945-1060) writes, and after post_sw what fix_base_calls
 * (common/sw-post.c:555-588) turns qralign into. */
#ifndef SHRIMP_SHIM_ALIGN_H
#define SHRIMP_SHIM_ALIGN_H

#include <ctype.h>
#include <stdint.h>
#include <vector>

namespace shrimp_shim {

static const char ALIGN_LETTERS[] = "ACGTUMRWSYKVHDBN";   /* base_translate, fasta.c:694-696 */

static inline int nib(const uint32_t *a, int64_t i) { return (int)((a[i / 8] >> (4 * (i % 8))) & 0xf); }

/* cstols, util.h:157-180 (is_rna false) */
static inline int cs_to_ls(int first_letter, int colour) {
  if (first_letter < 0 || first_letter > 3 || colour < 0 || colour > 3) return 15;
  return (first_letter % 2 == 0) ? (4 + first_letter + colour) % 4 : (4 + first_letter - colour) % 4;
}

/* ed[n]: one byte per alignment column -- bits 0-1: 1 genome base over '-', 2 read base over '-', 3 both; colour
 * space: bit 2 = crossover column (lower case), bits 4-5 = layer (letter translation starting from
 * (layer + initbp) % 4, sw-full-cs.c:1181-1196) or, with bit 3 set, the base post_sw called.
 * gen / rd: packed genome strand and read; gi / ri: first genome / read position of the alignment.
 * db, qr: n + 1 bytes each. */
static inline void edit_to_strings(const uint8_t *ed, int n, const uint32_t *gen, int64_t gi, const uint32_t *rd, int ri,
                                   bool colour_space, int initbp, int read_len, char *db, char *qr) {
  if (!colour_space) {
    for (int c = 0; c < n; c++) {
      const int ty = ed[c] & 3;
      db[c] = ty == 2 ? '-' : ALIGN_LETTERS[nib(gen, gi)];
      qr[c] = ty == 1 ? '-' : ALIGN_LETTERS[nib(rd, ri)];
      if (ty != 2) gi++;
      if (ty != 1) ri++;
    }
    db[n] = qr[n] = 0;
    return;
  }
  std::vector<uint8_t> lay;
  bool need_layers = false;
  for (int c = 0; c < n && !need_layers; c++) need_layers = (ed[c] & 3) != 1 && !(ed[c] & 8);
  if (need_layers) {
    lay.resize((size_t)4 * read_len);
    for (int k = 0; k < 4; k++) {
      int letter = (k + initbp) % 4;
      for (int j = 0; j < read_len; j++) {
        const int col = nib(rd, j);
        if (col == 15) {
          lay[(size_t)k * read_len + j] = 15;
          letter = (k + initbp) % 4;
        } else {
          letter = cs_to_ls(letter, col);
          lay[(size_t)k * read_len + j] = (uint8_t)letter;
        }
      }
    }
  }
  for (int c = 0; c < n; c++) {
    const int op = ed[c], ty = op & 3;
    if (ty == 1) {
      db[c] = ALIGN_LETTERS[nib(gen, gi++)];
      qr[c] = '-';
      continue;
    }
    char g = '-';
    if (ty == 3) g = ALIGN_LETTERS[nib(gen, gi++)];
    char q;
    if (op & 8) {
      q = ALIGN_LETTERS[(op >> 4) & 3];
      if (op & 4) q = (char)tolower(q);
    } else {
      q = ALIGN_LETTERS[lay[(size_t)((op >> 4) & 3) * read_len + ri] & 15];
      if (op & 4) q = (char)tolower(q);
      if ((q == 'N' || q == 'n') && ty == 3) q = (op & 4) ? (char)tolower(g) : g;   /* sw-full-cs.c:1029-1036 */
    }
    ri++;
    db[c] = g;
    qr[c] = q;
  }
  db[n] = qr[n] = 0;
}

}  // namespace shrimp_shim

#endif
