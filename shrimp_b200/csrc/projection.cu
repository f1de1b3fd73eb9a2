// Projection save in the reference's -S file format, so that CPU gmapper (-L) and this library
// share one projection.
//
// Replaces save_genome_map / save_genome_map_seed (gmapper/genome.c:185-272, :15-66).  The reference
// writes through zlib; its loader (load_genome_map :670-832, load_genome_map_seed :69-182) reads
// with gzread, which passes uncompressed files through unchanged, so the files are written plain:
// byte-identical to `gunzip` of what `gmapper -S` writes for the same genome and seeds.
//   <prefix>.genome : shrimp_mode, Hflag, num_contigs, genome_len[], contig_offsets[], per contig
//                     (name_len, name\0), total words, contigs fwd, contigs rc, [contigs colour]
//   <prefix>.seed.N : shrimp_mode, Hflag, seed_type{mask u64, span i32, weight i32},
//                     genomemap_len[4^W], total, position lists concatenated
#include <errno.h>
#include "genome.cuh"

using namespace shrimp;

namespace {

struct File {
  FILE *f = nullptr;
  ~File() {
    if (f) fclose(f);
  }
  bool put(const void *p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
};

// contig-local packed copy of nibbles [off, off+len) of a global packed array
void repack_contig(const std::vector<uint32_t> &G, uint64_t off, uint32_t len, std::vector<uint32_t> &out) {
  const size_t nw = ((size_t)len + 7) / 8;
  out.assign(nw, 0u);
  const uint64_t w0 = off >> 3;
  const unsigned sh = 4u * (unsigned)(off & 7);
  for (size_t w = 0; w < nw; w++) {
    uint32_t v = G[w0 + w] >> sh;
    if (sh) v |= G[w0 + w + 1] << (32u - sh);
    out[w] = v;
  }
  const unsigned tail = len & 7u;
  if (tail) out[nw - 1] &= (1u << (4u * tail)) - 1u;
}

}  // namespace

extern "C" int shrimp_gpu_projection_save(shrimp_gpu_ctx *ctx, const char *prefix, const char *const *contig_names) {
  DeviceGenome *g = ctx ? genome_of(ctx) : nullptr;
  if (!g || !g->have_index) {
    set_error("shrimp_gpu_projection_save: genome/index not resident");
    return SHRIMP_E_STATE;
  }
  if (!prefix || !contig_names) {
    set_error("shrimp_gpu_projection_save: invalid argument");
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  const uint32_t mode = g->colour_space ? 2u : 1u;  // MODE_LETTER_SPACE / MODE_COLOUR_SPACE, gmapper-definitions.h:31-32
  const uint32_t hflag = (uint32_t)g->seeds.hflag;
  std::string name = std::string(prefix) + ".genome";
  {
    File F;
    F.f = fopen(name.c_str(), "wb");
    if (!F.f) {
      set_error("shrimp_gpu_projection_save: cannot open %s: %s", name.c_str(), strerror(errno));
      return SHRIMP_E_ARG;
    }
    const uint32_t nc = (uint32_t)g->num_contigs;
    bool ok = F.put(&mode, 4) && F.put(&hflag, 4) && F.put(&nc, 4) && F.put(g->h_len.data(), 4 * (size_t)nc) &&
              F.put(g->h_off.data(), 4 * (size_t)nc);
    uint32_t total = 0;
    for (uint32_t c = 0; c < nc && ok; c++) {
      const uint32_t len = (uint32_t)strlen(contig_names[c]);
      ok = F.put(&len, 4) && F.put(contig_names[c], (size_t)len + 1);
      total += (g->h_len[c] + 7) / 8;
    }
    ok = ok && F.put(&total, 4);
    std::vector<uint32_t> G(g->words + 1, 0u), loc;
    const DevBuf *arr[3] = {&g->d_ls, &g->d_ls_rc, &g->d_cs};
    for (int a = 0; a < (g->colour_space ? 3 : 2) && ok; a++) {
      SH_CUDA(cudaMemcpy(G.data(), arr[a]->p, g->words * 4, cudaMemcpyDeviceToHost));
      for (uint32_t c = 0; c < nc && ok; c++) {
        repack_contig(G, g->h_off[c], g->h_len[c], loc);
        ok = F.put(loc.data(), loc.size() * 4);
      }
    }
    if (!ok) {
      set_error("shrimp_gpu_projection_save: write to %s failed", name.c_str());
      return SHRIMP_E_ARG;
    }
  }
  for (int sn = 0; sn < g->seeds.n_seeds; sn++) {
    name = std::string(prefix) + ".seed." + std::to_string(sn);
    File F;
    F.f = fopen(name.c_str(), "wb");
    if (!F.f) {
      set_error("shrimp_gpu_projection_save: cannot open %s: %s", name.c_str(), strerror(errno));
      return SHRIMP_E_ARG;
    }
    struct {
      uint64_t mask;
      int32_t span, weight;
    } seed = {g->seeds.mask[sn], g->seeds.span[sn], g->seeds.weight[sn]};  // seed_type, gmapper-definitions.h:59-63
    const uint32_t nb = g->nbuckets[sn];
    std::vector<uint32_t> offs((size_t)nb + 1), lens(nb);
    SH_CUDA(cudaMemcpy(offs.data(), g->d_offs[sn].p, offs.size() * 4, cudaMemcpyDeviceToHost));
    for (uint32_t m = 0; m < nb; m++) lens[m] = offs[m + 1] - offs[m];
    const uint32_t total = (uint32_t)g->total[sn];
    std::vector<uint32_t> pos(total);
    if (total) SH_CUDA(cudaMemcpy(pos.data(), g->d_pos[sn].p, (size_t)total * 4, cudaMemcpyDeviceToHost));
    const bool ok = F.put(&mode, 4) && F.put(&hflag, 4) && F.put(&seed, sizeof(seed)) &&
                    F.put(lens.data(), 4 * (size_t)nb) && F.put(&total, 4) && F.put(pos.data(), 4 * (size_t)total);
    if (!ok) {
      set_error("shrimp_gpu_projection_save: write to %s failed", name.c_str());
      return SHRIMP_E_ARG;
    }
  }
  return SHRIMP_OK;
}
