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

// ---- projection load (the other half) -------------------------------------------------------
// Replaces load_genome_map / load_genome_map_seed (gmapper/genome.c:670-832, :69-182) for the device: reads the files
// `gmapper -S <prefix>` (gzip) or shrimp_gpu_projection_save (plain) wrote, through zlib as the reference does, and
// puts the letter contigs and the per-seed CSR (genomemap_len -> offsets, position lists as they lie in the file)
// straight into HBM.  The reverse-complement and colour arrays are derived on the device (the ones in the file are
// the same bytes, tests/test_gpu_index.py).
#include <zlib.h>

namespace {
struct GzFile {
  gzFile f = nullptr;
  ~GzFile() {
    if (f) gzclose(f);
  }
  bool get(void *p, size_t n) {
    char *c = (char *)p;
    while (n > 0) {
      const unsigned chunk = (unsigned)std::min<size_t>(n, (size_t)1 << 30);
      const int r = gzread(f, c, chunk);
      if (r <= 0) return false;
      c += r;
      n -= (size_t)r;
    }
    return true;
  }
};
}  // namespace

extern "C" int shrimp_gpu_projection_load(shrimp_gpu_ctx *ctx, const char *prefix) {
  if (!ctx || !prefix) {
    set_error("shrimp_gpu_projection_load: invalid argument");
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  std::vector<std::string> names;
  uint32_t mode = 0, hflag = 0, nc = 0;
  {
    const std::string name = std::string(prefix) + ".genome";
    GzFile F;
    F.f = gzopen(name.c_str(), "rb");
    if (!F.f) {
      set_error("shrimp_gpu_projection_load: cannot open %s: %s", name.c_str(), strerror(errno));
      return SHRIMP_E_ARG;
    }
    bool ok = F.get(&mode, 4) && F.get(&hflag, 4) && F.get(&nc, 4);
    if (!ok || (mode != 1u && mode != 2u) || nc == 0 || nc > (1u << 24)) {
      set_error("shrimp_gpu_projection_load: %s is not a genome projection file", name.c_str());
      return SHRIMP_E_ARG;
    }
    std::vector<uint32_t> len(nc), off(nc);
    ok = F.get(len.data(), 4 * (size_t)nc) && F.get(off.data(), 4 * (size_t)nc);
    for (uint32_t c = 0; c < nc && ok; c++) {
      uint32_t nl = 0;
      ok = F.get(&nl, 4) && nl < (1u << 20);
      if (!ok) break;
      std::string s((size_t)nl + 1, '\0');
      ok = F.get(&s[0], (size_t)nl + 1);
      s.resize(nl);
      names.push_back(s);
    }
    uint32_t total = 0;
    ok = ok && F.get(&total, 4);
    std::vector<std::vector<uint32_t>> contigs(nc);
    std::vector<const uint32_t *> ptrs(nc);
    for (uint32_t c = 0; c < nc && ok; c++) {
      contigs[c].resize(((size_t)len[c] + 7) / 8);
      ok = F.get(contigs[c].data(), contigs[c].size() * 4);
      ptrs[c] = contigs[c].data();
    }
    if (!ok) {
      set_error("shrimp_gpu_projection_load: %s is truncated", name.c_str());
      return SHRIMP_E_ARG;
    }
    SH_TRY(shrimp_gpu_genome_load(ctx, (int)nc, ptrs.data(), len.data(), mode == 2u ? 1 : 0));
  }
  DeviceGenome *g = genome_of(ctx);
  g->contig_names = names;
  uint64_t masks[SHRIMP_MAX_SEEDS];
  int32_t spans[SHRIMP_MAX_SEEDS], weights[SHRIMP_MAX_SEEDS];
  int n_seeds = 0;
  for (int sn = 0; sn < SHRIMP_MAX_SEEDS; sn++) {
    const std::string name = std::string(prefix) + ".seed." + std::to_string(sn);
    FILE *probe = fopen(name.c_str(), "rb");   // the reference probes the same way (gmapper.c:2756-2775)
    if (!probe) break;
    fclose(probe);
    GzFile F;
    F.f = gzopen(name.c_str(), "rb");
    struct {
      uint64_t mask;
      int32_t span, weight;
    } seed;
    uint32_t m2 = 0, h2 = 0;
    bool ok = F.f && F.get(&m2, 4) && F.get(&h2, 4) && F.get(&seed, sizeof(seed));
    if (!ok || m2 != mode || h2 != hflag || seed.weight < 1 || (!hflag && seed.weight > 14)) {
      set_error("shrimp_gpu_projection_load: %s does not belong to this projection", name.c_str());
      return SHRIMP_E_ARG;
    }
    const uint32_t nb = 1u << (2 * (hflag ? 12 : seed.weight));
    std::vector<uint32_t> offs((size_t)nb + 2);
    ok = F.get(offs.data() + 1, 4 * (size_t)nb);   // genomemap_len -> inclusive prefix sums in place
    uint32_t total = 0;
    ok = ok && F.get(&total, 4);
    if (ok) {
      offs[0] = 0;
      unsigned long long run = 0;
      for (uint32_t m = 1; m <= nb; m++) {
        run += offs[m];
        offs[m] = (uint32_t)run;
      }
      offs[nb + 1] = (uint32_t)run;
      ok = run == total;
    }
    if (!ok) {
      set_error("shrimp_gpu_projection_load: %s is truncated or inconsistent", name.c_str());
      return SHRIMP_E_ARG;
    }
    SH_TRY(g->d_offs[sn].ensure(((size_t)nb + 2) * 4));
    SH_CUDA(cudaMemcpy(g->d_offs[sn].p, offs.data(), ((size_t)nb + 2) * 4, cudaMemcpyHostToDevice));
    SH_TRY(g->d_pos[sn].ensure(((size_t)total + 1) * 4));
    std::vector<uint32_t> buf((size_t)std::min<uint64_t>(total, 1u << 24));
    for (uint64_t done = 0; done < total;) {   // the lists lie in the file bucket after bucket: the CSR's own order
      const size_t n = (size_t)std::min<uint64_t>(buf.size(), total - done);
      if (!F.get(buf.data(), n * 4)) {
        set_error("shrimp_gpu_projection_load: %s is truncated", name.c_str());
        return SHRIMP_E_ARG;
      }
      SH_CUDA(cudaMemcpy(g->d_pos[sn].as<uint32_t>() + done, buf.data(), n * 4, cudaMemcpyHostToDevice));
      done += n;
    }
    g->nbuckets[sn] = nb;
    g->total[sn] = total;
    masks[sn] = seed.mask;
    spans[sn] = seed.span;
    weights[sn] = seed.weight;
    n_seeds = sn + 1;
  }
  if (n_seeds == 0) {
    set_error("shrimp_gpu_projection_load: no %s.seed.0", prefix);
    return SHRIMP_E_ARG;
  }
  SeedTable S{};
  SH_TRY(seed_table_init(S, n_seeds, masks, spans, weights, (int)hflag, "shrimp_gpu_projection_load"));
  g->seeds = S;
  g->have_index = true;
  return build_bucket_heads(ctx, g);
}

extern "C" int shrimp_gpu_num_contigs(shrimp_gpu_ctx *ctx) {
  DeviceGenome *g = ctx ? genome_of(ctx) : nullptr;
  return g ? g->num_contigs : 0;
}

extern "C" const char *shrimp_gpu_contig_name(shrimp_gpu_ctx *ctx, int cn) {
  DeviceGenome *g = ctx ? genome_of(ctx) : nullptr;
  if (!g || cn < 0 || cn >= (int)g->contig_names.size()) return nullptr;
  return g->contig_names[cn].c_str();
}
