/* Plain C99 driver of the engine through the sb_* entry points of include/scone_b200.h ONLY (no sbh_* host driver, no C++):
 * what a SCONE build calls through shim/sconeB200_mod.f90.  Reads a flat model written by sbh_model_dump (the arrays SCONE's own
 * objects hold after init), loads it, generates the initial source and runs eigenvalue cycles the way eigenPhysicsPackage%cycles
 * does (eigenPhysicsPackage_class.f90:203-307): the package generator strides by totalPop + 1 per cycle, normSize_Repr uses the
 * state after that stride, one more stride(1) follows.  Prints one line per cycle: k_cum, sites, segments.
 *
 *   gcc -std=c99 -pedantic -Wall -I include tests/c_driver/run_cycle.c -L scone_b200 -lscone_b200 -o run_cycle
 *   ./run_cycle tests/golden/c5g7_flat.bin <pop> <inactive> <active>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "scone_b200.h"

static FILE* g_f;
static void* blk(size_t sz, int64_t* n_out) {
  int64_t n = 0; void* p = NULL;
  if (fread(&n, 8, 1, g_f) != 1) { fprintf(stderr, "run_cycle: truncated model file\n"); exit(2); }
  if (n > 0) { p = malloc((size_t)n * sz); if (fread(p, sz, (size_t)n, g_f) != (size_t)n) { fprintf(stderr, "run_cycle: truncated model file\n"); exit(2); } }
  if (n_out) *n_out = n;
  return p;
}
static int32_t i32(void) { int32_t* p = (int32_t*)blk(4, NULL); int32_t v = *p; free(p); return v; }
static double f64(void) { double* p = (double*)blk(8, NULL); double v = *p; free(p); return v; }

/* RNG_class.f90:251-299: state after k steps of s -> (g s + 1) mod 2^63 */
static uint64_t rng_skip(uint64_t s, int64_t k_in) {
  const uint64_t M = 0x7fffffffffffffffULL;
  uint64_t k = (uint64_t)k_in & M, G = 1, C = 0, h = 2806196910506780709ULL, L = 1;
  while (k > 0) {
    if (k & 1ULL) { G = (G * h) & M; C = (C * h) & M; C = (C + L) & M; }
    L = (L * (h + 1)) & M; h = (h * h) & M; k >>= 1;
  }
  return (G * s + C) & M;
}
#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "run_cycle: %s failed: %s\n", #call, sb_last_error(eng)); return 1; } } while (0)

int main(int argc, char** argv) {
  char magic[8];
  sb_geom_flat g; sb_mg_flat d; sb_options opt; sb_engine* eng = NULL; sb_cycle_result res;
  sb_clerk* clerks[2]; int32_t n_clerks[2], norm_clerk[2]; double norm_val[2];
  int ph, c, m, cyc, pop, n_inactive, n_active, file_pop; uint64_t pRNG; double k_eff; int32_t* bc;
  if (argc < 5) { fprintf(stderr, "usage: run_cycle <flat model> <pop> <inactive> <active>\n"); return 2; }
  g_f = fopen(argv[1], "rb");
  if (!g_f || fread(magic, 1, 8, g_f) != 8 || memcmp(magic, "SBFLAT1", 8) != 0) { fprintf(stderr, "run_cycle: not a flat model file\n"); return 2; }
  pop = atoi(argv[2]); n_inactive = atoi(argv[3]); n_active = atoi(argv[4]);
  memset(&g, 0, sizeof g); memset(&d, 0, sizeof d); memset(&opt, 0, sizeof opt);
  g.n_surf = i32(); g.surf_type = (int32_t*)blk(4, NULL); g.surf_par = (double*)blk(8, NULL);
  g.n_cell = i32(); g.cell_off = (int32_t*)blk(4, NULL); g.cell_surf = (int32_t*)blk(4, NULL);
  g.n_uni = i32(); g.uni_type = (int32_t*)blk(4, NULL); g.uni_ipar = (int32_t*)blk(4, NULL); g.uni_dpar = (double*)blk(8, NULL);
  g.n_aux_d = i32(); g.aux_d = (double*)blk(8, NULL); g.n_aux_i = i32(); g.aux_i = (int32_t*)blk(4, NULL);
  g.n_graph = i32(); g.graph_idx = (int32_t*)blk(4, NULL); g.graph_id = (int32_t*)blk(4, NULL);
  g.root_idx = i32(); g.border_idx = i32(); bc = (int32_t*)blk(4, NULL); memcpy(g.bc, bc, sizeof g.bc); free(bc);
  d.n_mat = i32(); d.n_g = i32(); d.data = (double*)blk(8, NULL); d.P0 = (double*)blk(8, NULL); d.prod = (double*)blk(8, NULL);
  d.P1 = (double*)blk(8, NULL); d.chi = (double*)blk(8, NULL); d.fissile = (int32_t*)blk(4, NULL); d.majorant = (double*)blk(8, NULL); d.collision_xs = f64();
  for (ph = 0; ph < 2; ++ph) {
    n_clerks[ph] = i32(); norm_clerk[ph] = i32(); norm_val[ph] = f64();
    clerks[ph] = (sb_clerk*)calloc((size_t)(n_clerks[ph] > 0 ? n_clerks[ph] : 1), sizeof(sb_clerk));
    for (c = 0; c < n_clerks[ph]; ++c) {
      sb_clerk* k = &clerks[ph][c]; int32_t* mt;
      k->n_maps = i32(); k->n_resp = i32(); mt = (int32_t*)blk(4, NULL); memcpy(k->resp_mt, mt, sizeof k->resp_mt); free(mt);
      k->handle_virtual = i32(); k->kind = i32(); k->cycles = i32();
      for (m = 0; m < k->n_maps; ++m) {
        sb_map1d* q = &k->maps[m];
        q->type = i32(); q->axis = i32(); q->grid = i32(); q->n_bins = i32(); q->first = f64(); q->step = f64(); q->default_bin = i32();
        q->bounds = (double*)blk(8, NULL); q->mat_bin = (int32_t*)blk(4, NULL);
      }
    }
  }
  opt.tracking = i32(); opt.ht_cutoff = f64(); opt.st_cache = i32(); (void)i32(); opt.max_pop = pop;
  file_pop = i32(); (void)file_pop;
  { uint64_t* s = (uint64_t*)blk(8, NULL); pRNG = *s; free(s); }
  k_eff = f64();
  fclose(g_f);

  if (sb_create(&eng, 0) != 0) { fprintf(stderr, "run_cycle: sb_create failed: %s\n", sb_last_error(NULL)); return 1; }
  CHECK(sb_load_geometry(eng, &g));
  CHECK(sb_load_mg_data(eng, &d));
  for (ph = 0; ph < 2; ++ph) CHECK(sb_define_tallies(eng, ph, clerks[ph], n_clerks[ph], norm_clerk[ph], norm_val[ph]));
  CHECK(sb_set_options(eng, &opt));
  CHECK(sb_source_generate(eng, pop, pRNG, 0));                  /* generateInitialState, then pRNG%stride(totalPop) */
  pRNG = rng_skip(pRNG, 152917LL * pop);
  for (cyc = 0; cyc < n_inactive + n_active; ++cyc) {
    const int phase = cyc >= n_inactive;
    const uint64_t rng0 = pRNG;
    pRNG = rng_skip(pRNG, 152917LL * (pop + 1));                 /* pRNG%stride(totalPop + 1) after the history loop */
    CHECK(sb_run_cycle(eng, rng0, 0, k_eff, phase, &res));
    CHECK(sb_resample(eng, pop, pRNG));                          /* normSize_Repr(totalPop, pRNG) */
    pRNG = rng_skip(pRNG, 152917LL);                             /* pRNG%stride(1) */
    k_eff = res.k_cum;
    printf("cycle %d k_cum %.17g sites %d segments %lld\n", cyc + 1, res.k_cum, (int)res.n_sites, (long long)res.n_segments);
  }
  sb_destroy(eng);
  return 0;
}
