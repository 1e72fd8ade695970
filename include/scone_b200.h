/* scone_b200 -- C ABI of the B200 transport engine.
 *
 * SCONE (Fortran 2008) has no FFI today. The seam this library plugs into is the body of the
 * `gen` loop of the physics packages: instead of calling transOp%transport / collOp%collide /
 * tally%report* once per particle from inside an OpenMP loop, the package calls this library
 * once per cycle through a thin iso_c_binding shim (INTEGRATION.md shows the shim).
 *   PhysicsPackages/eigenPhysicsPackage_class.f90:203-307   (cycle body that is replaced)
 *   TransportOperator/transportOperator_inter.f90:105-130    (transport)
 *   CollisionOperator/CollisionProcessors/collisionProcessor_inter.f90:114-195 (collide)
 *   Tallies/tallyAdmin_class.f90:466-794                     (report*, reportCycleEnd)
 *   ParticleObjects/particleDungeon_class.f90:431-602        (normSize_Repr)
 *
 * Conventions: plain pointers and sizes only; all indices that cross the boundary are 1-based
 * exactly as SCONE holds them (matIdx, uniIdx, surfIdx, localID, uniqueID, group G);
 * reals are IEEE binary64, integers int32 unless stated; arrays are caller-owned and copied.
 * Every function returns 0 on success, non-zero on error; sb_last_error() gives the text
 * (the shim forwards it to fatalError, SharedModules/errors_mod.f90:32-69).
 * One engine handle drives one GPU and is not re-entrant.
 */
#ifndef SCONE_B200_H
#define SCONE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb_engine sb_engine;

/* ---- special material codes: SharedModules/universalVariables.f90:43-47 ---- */
#define SB_OUTSIDE_MAT 0
#define SB_VOID_MAT    2147483647
#define SB_UNDEF_MAT   2147483646
#define SB_OVERLAP_MAT 2147483645

/* ---- surfaces: Geometry/Surfaces/ ------------------------------------------------------
 * par[8] per surface:
 *   x/y/zPlane      : par[0] = a0
 *   plane           : par[0..2] = unit normal, par[3] = offset
 *   sphere          : par[0..2] = origin, par[3] = r, par[4] = r*r
 *   x/y/zCylinder   : par[0..2] = origin, par[3] = r, par[4] = r*r
 *   box             : par[0..2] = origin, par[3..5] = halfwidth
 *   x/y/zSquareCyl. : par[0..2] = origin, par[3..5] = halfwidth (axis component unused)
 *   par[6] = surface tolerance (surface_inter.f90 surf_tol)                                  */
enum { SB_SURF_XPLANE = 1, SB_SURF_YPLANE = 2, SB_SURF_ZPLANE = 3, SB_SURF_PLANE = 4, SB_SURF_SPHERE = 5,
       SB_SURF_XCYL = 6, SB_SURF_YCYL = 7, SB_SURF_ZCYL = 8, SB_SURF_BOX = 9,
       SB_SURF_XSQCYL = 10, SB_SURF_YSQCYL = 11, SB_SURF_ZSQCYL = 12,
       SB_SURF_XTCYL = 13, SB_SURF_YTCYL = 14, SB_SURF_ZTCYL = 15 /* truncCylinder: par = origin[3], radius, radius^2, axial halfwidth, tolerance; BC { a_min, a_max } */ };
#define SB_SURF_NPAR 8

/* ---- universes: Geometry/Universes/ ------------------------------------------------------
 * dpar[32] per universe: [0..2] origin, [3..11] rotation matrix (row major, r' = R r),
 *   lattice: [12..14] pitch, [15..17] corner, [18..20] a_bar, [21..23] outline halfwidth
 * ipar[8] per universe: [0] rotated?, [1] global transform?,
 *   root   : [2] border surfIdx
 *   pin    : [2] number of annuli N, [3] offset into aux_d of { r_sq[N], tol[N] }
 *   lattice: [2..4] sizeN, [5] outLocalID, [6] offset mode (0 none, 1 all cells, 2 per-cell map),
 *            [7] offset into aux_i of the per-cell map (mode 2; outLocalID entries)
 *   cell   : [2] number of cells, [3] offset into aux_i of the cellIdx list, [4] checkOverlap */
enum { SB_UNI_ROOT = 1, SB_UNI_PIN = 2, SB_UNI_LAT = 3, SB_UNI_CELL = 4 };
#define SB_UNI_NDPAR 32
#define SB_UNI_NIPAR 8

typedef struct sb_geom_flat {
  int32_t n_surf;  const int32_t* surf_type;  const double* surf_par;      /* [n_surf][8]            */
  int32_t n_cell;  const int32_t* cell_off;   const int32_t* cell_surf;    /* CSR, signed surfIdx     */
  int32_t n_uni;   const int32_t* uni_type;   const int32_t* uni_ipar;  const double* uni_dpar;
  int32_t n_aux_d; const double*  aux_d;
  int32_t n_aux_i; const int32_t* aux_i;
  int32_t n_graph; const int32_t* graph_idx;  const int32_t* graph_id;     /* geomGraph_class.f90:21-24 */
  int32_t root_idx, border_idx;                                            /* csg_class.f90 rootIdx, borderIdx */
  int32_t bc[6];                                                           /* (-x +x -y +y -z +z): 0 vacuum 1 reflective 2 periodic */
} sb_geom_flat;

/* ---- multigroup data: NuclearData/mgNeutronData/baseMgNeutron/ ---------------------------
 * data  : [n_mat][n_g][6] = Fortran data(6,nG) of each material, rows TOTAL, IESCATTER, CAPTURE,
 *         FISSION, NU_FISSION, KAPPA_FISSION (baseMgNeutronMaterial_class.f90:32-37)
 * P0/prod/P1 : [n_mat][n_g][n_g] = Fortran P0(G_out,G_in) storage order (G_out fastest);
 *         P1 already normalised as multiScatterP1MG does (P1/P0*3); NULL for P0 scattering
 * chi   : [n_mat][n_g] ; fissile : [n_mat] ; majorant : [n_g] (initMajorant)                 */
typedef struct sb_mg_flat {
  int32_t n_mat, n_g;
  const double* data; const double* P0; const double* prod; const double* P1; const double* chi;
  const int32_t* fissile; const double* majorant; double collision_xs;
} sb_mg_flat;

/* ---- continuous-energy data: NuclearData/ceNeutronData/aceDatabase/ ---------------------------
 * What aceNeutronDatabase holds after init (aceNeutronDatabase_class.f90:873-1163): per nuclide the energy grid
 * eGrid(N) and mainData(rows, N) (rows = 4 non-fissile / 8 fissile: TOTAL, ESCATTER, IESCATTER, CAPTURE, FISSION,
 * NU_FISSION, KAPPA, PROMPT_NU_FISSION; aceNeutronNuclide_class.f90:44-54), Fortran storage order (rows of one
 * energy point contiguous); per material the nuclide indices (1-based) and atomic densities in composition order
 * (ceNeutronMaterial_class.f90). grid / data are the concatenations over nuclides.                             */
typedef struct sb_ce_flat {
  int32_t n_nuc; const int32_t* grid_size; const int32_t* rows; const double* grid; const double* data;
  int32_t n_mat; const int32_t* mat_off /*[n_mat+1]*/; const int32_t* mat_nuc; const double* mat_dens;
} sb_ce_flat;

/* ---- continuous-energy transport model ------------------------------------------------------------------------------
 * One ACE card per nuclide, exactly the arrays aceCard holds after readFromFile (NuclearData/DataDecks/ACE/aceCard_class.f90:
 * 96-98 NXS(16), JXS(32), XSS(:); :69-71 ZAID, AW, TZ): the engine builds from them what aceNeutronNuclide%init builds
 * (aceNeutronNuclide_class.f90:737-948: main data, reaction list) and samples the reaction laws from XSS in place.
 * Materials as ceNeutronMaterial holds them (nuclide indices 1-based in composition order, atomic densities) plus the material
 * temperature [K] of materialMenu.  active_mats = the materials nuclearDatabase%activate receives (present in the geometry):
 * the unionised majorant covers those (aceNeutronDatabase_class.f90:1330-1621).
 * Settings: collision_xs = 1/avgDist, energy_per_fission = H235 (aceNeutronDatabase init); min/max_energy, thresh_energy,
 * thresh_mass = neutronCEstd minEnergy/maxEnergy/energyThreshold/massThreshold (neutronCEstd_class.f90:110-150);
 * source_energy = fissionSource E (fissionSource_class.f90:128).
 * Not supported (refused with an error): S(a,b), URR tables, TMS, DBRC, correlated laws, 32-bin angular pdfs.          */
typedef struct sb_ace_card {
  const char* zaid; double aw, tz;
  const int32_t* nxs /*[16]*/; const int32_t* jxs /*[32]*/; const double* xss; int64_t n_xss;
} sb_ace_card;
typedef struct sb_ce_model {
  int32_t n_nuc; const sb_ace_card* cards;
  int32_t n_mat; const int32_t* mat_off /*[n_mat+1]*/; const int32_t* mat_nuc; const double* mat_dens; const double* mat_temp /*[n_mat]*/;
  int32_t n_active; const int32_t* active_mats;
  double collision_xs, energy_per_fission;
  double min_energy, max_energy, thresh_energy, thresh_mass;
  double source_energy;
} sb_ce_model;

/* ---- tallies: Tallies/TallyClerks/collisionClerk_class.f90, TallyMaps/, TallyResponses/ ---- */
enum { SB_MAP_SPACE = 1, SB_MAP_MATERIAL = 2, SB_MAP_ENERGY = 3 };
enum { SB_GRID_LIN = 1, SB_GRID_LOG = 2, SB_GRID_UNSTRUCT = 3 };
typedef struct sb_map1d {
  int32_t type, axis /*0 x,1 y,2 z*/, grid, n_bins;
  double first, step;             /* grid%bins(1), grid%step (lin/log)                        */
  const double* bounds;           /* n_bins+1 boundaries (unstruct) or NULL                   */
  const int32_t* mat_bin;         /* materialMap: bin of matIdx m at [m-1], size n_mat; NULL  */
  int32_t default_bin;            /* materialMap default (0 = not scored)                     */
} sb_map1d;
#define SB_MAX_MAPS 4
#define SB_MAX_RESP 8
/* kind: collisionClerk (maps x responses), or one of the k-eff clerks as a USER clerk of a tally block (the engine always runs
 * the attachment clerks of eigenPhysicsPackage itself): keffAnalogClerk = 3 bins { start weight, end weight, k }
 * (keffAnalogClerk_class.f90:40-60,132-176), keffImplicitClerk = 5 bins { IMP_PROD, IMP_ABS, SCATTER_PROD, ANA_LEAK, K_EFF }
 * (keffImplicitClerk_class.f90:60-75,292-312): the first bins pass through closeCycle (normalised), the k bin is accumulated as is */
enum { SB_CLERK_COLLISION = 0, SB_CLERK_KEFF_ANALOG = 1, SB_CLERK_KEFF_IMPLICIT = 2,
       SB_CLERK_TRACK = 3 /* trackClerk (trackClerk_class.f90:185-232): maps on the pre-path state, score = response * w * path length; surface tracking only */,
       SB_CLERK_SHANNON = 4 /* shannonEntropyClerk (shannonEntropyClerk_class.f90:117-190): maps only; N + 1 + cycles bins { weight, N bin weights, entropy of
                               cycle 1..cycles }: the fission bank of every cycle end is binned by weight, -sum p log2 p is accumulated in the cycle's own bin */ };
typedef struct sb_clerk {
  int32_t n_maps; sb_map1d maps[SB_MAX_MAPS];   /* multiMap order; 0 maps = single bin        */
  int32_t n_resp; int32_t resp_mt[SB_MAX_RESP]; /* 0 = fluxResponse, else SCONE macro MT (-1..) */
  int32_t handle_virtual;
  int32_t kind;                                 /* SB_CLERK_*; maps and responses are ignored for the k-eff clerks */
  int32_t cycles;                               /* shannonEntropyClerk: number of cycles that are scored */
} sb_clerk;

enum { SB_TRACK_DT = 0, SB_TRACK_ST = 1, SB_TRACK_HT = 2 };
typedef struct sb_options {
  int32_t tracking; double ht_cutoff; int32_t st_cache;
  int32_t max_pop;          /* capacity of the device banks is 2*max_pop as eigenPP allocates dungeons */
  int32_t threads_per_block, blocks_per_sm;   /* 0 = engine default                            */
} sb_options;

/* result of one cycle, all of it computed on the device */
typedef struct sb_cycle_result {
  int32_t n_start;            /* histories transported                                         */
  int32_t n_sites;            /* fission sites banked before normalisation                     */
  double  start_wgt, end_wgt; /* popWeight of this / next cycle dungeon (keffAnalogClerk)      */
  double  imp_prod, imp_abs, scatter_prod, ana_leak;  /* keffImplicitClerk bins of the cycle   */
  double  k_analog, k_implicit;                        /* per-cycle estimates                   */
  double  k_cum, k_cum_std;   /* cumulative mean of this phase's attachment clerk = k_new      */
  int64_t n_segments, n_collisions, n_scores;   /* flights, real collisions, tally scores (f64 accumulations) */
  int32_t error;              /* device-side fatal condition (SB_ERR_*), 0 if none             */
  int32_t max_history_segments; /* flights of the longest history (>= 256; the cycle's critical path)  */
  int64_t n_xs_terms;         /* continuous energy: nuclide terms summed by the total-cross-section lookups of the flights */
} sb_cycle_result;
enum { SB_ERR_BANK_OVERFLOW = 1, SB_ERR_UNDEF_MAT = 2, SB_ERR_OVERLAP_MAT = 3, SB_ERR_SAMPLING = 4,
       SB_ERR_NEST = 5, SB_ERR_SOURCE = 6, SB_ERR_NORM = 7, SB_ERR_FILE_SOURCE = 10, SB_ERR_PEER_TIMEOUT = 11, SB_ERR_BALANCE = 12, SB_ERR_MAT_SOURCE = 13, SB_ERR_MAT_SOURCE_VOID = 14,
       SB_ERR_CE_ENERGY = 8 /* energy outside the bounds of the CE data */, SB_ERR_CE_DATA = 9 /* failed search / rejection loop in the reaction data */ };

/* ---- life cycle -------------------------------------------------------------------------- */
int  sb_create(sb_engine** h, int device);
void sb_destroy(sb_engine* h);
const char* sb_last_error(sb_engine* h);
/* number of kernels of this library launched so far by this handle */
int64_t sb_launch_count(sb_engine* h);

/* ---- model load (replaces nothing at run time; consumes what csg%init / database%init built) */
int sb_load_geometry(sb_engine* h, const sb_geom_flat* g);
int sb_load_mg_data(sb_engine* h, const sb_mg_flat* d);
/* continuous-energy transport data (instead of sb_load_mg_data); load the geometry first */
int sb_load_ce_model(sb_engine* h, const sb_ce_model* m);
/* what the engine built from the cards, for parity tests: grid size / rows of nuclide nuc_idx, then the arrays */
int sb_ce_nuclide_info(sb_engine* h, int nuc_idx, int32_t* grid_size, int32_t* rows, int32_t* n_mt);
int sb_ce_nuclide_data(sb_engine* h, int nuc_idx, double* grid, double* main_data, int32_t* mt_list);
/* phase: 0 inactive, 1 active. norm_clerk = 1-based clerk whose first bin normalises (0 = none) */
int sb_define_tallies(sb_engine* h, int phase, const sb_clerk* clerks, int n_clerks, int norm_clerk, double norm_val);
int sb_set_options(sb_engine* h, const sb_options* o);

/* ---- banks (particleDungeon) ---------------------------------------------------------------
 * host SoA <-> device "this cycle" bank; r, dir are [n][3]                                   */
int sb_bank_upload(sb_engine* h, int n, const double* r, const double* dir, const double* w, const int32_t* G);
int sb_bank_download(sb_engine* h, int cap, int* n, double* r, double* dir, double* w, int32_t* G);
int sb_bank_size(sb_engine* h);
/* broodID of every site of the current bank (1-based index of the parent history in the cycle that made it; 0 for source and uploaded banks):
 * the ninth column of particleDungeon%printToFile (particleDungeon_class.f90:1077-1112)                                        */
int sb_bank_brood(sb_engine* h, int cap, int32_t* brood);
/* continuous-energy banks carry E [MeV] instead of G */
int sb_bank_upload_ce(sb_engine* h, int n, const double* r, const double* dir, const double* w, const double* E);
int sb_bank_download_ce(sb_engine* h, int cap, int* n, double* r, double* dir, double* w, double* E);
/* fissionSource%generate on the device (ParticleObjects/Source/fissionSource_class.f90:149-271) */
int sb_source_generate(sb_engine* h, int n, uint64_t rng_state, int history_offset);

/* ---- fixed-source calculations: PhysicsPackages/fixedSourcePhysicsPackage_class.f90:149-289 ---------------------------------
 * sb_set_fixed_source(on, buffer_size): particles banked by a collision (implicit fission sites) are not the next generation but
 * secondaries of the SAME history: they go to the history's private buffer (`buffer`, default 50, :318) and are followed, last in
 * first out, when the current particle dies, on the history's random stream (the bufferLoop, :198-246). Exceeding buffer_size is
 * the reference's "Run out of space for particles" error. A cycle is then sb_source_point (or sb_bank_upload of any source the
 * caller sampled) followed by sb_run_cycle with phase 1 and k_eff 1: transport + tallies + closeCycle; no normalisation of a bank.
 * sb_source_point = pointSource%sampleParticle for n particles (ParticleObjects/Source/pointSource_class.f90:142-200,
 * configSource_inter.f90:75-90): particle i uses rng_state skipped by 152917*(history_offset+i).                              */
typedef struct sb_point_source {
  double r[3], dir[3]; int32_t isotropic;          /* dir is used (already normalised) when isotropic == 0 */
  int32_t is_mg; double E; int32_t G;               /* E [MeV] for continuous energy; G, or prob_g[n_prob] (normalised), for multigroup */
  int32_t n_prob; const double* prob_g;
} sb_point_source;
int sb_set_fixed_source(sb_engine* h, int on, int buffer_size);
int sb_source_point(sb_engine* h, int n, uint64_t rng_state, int history_offset, const sb_point_source* s);
/* fileSource (ParticleObjects/Source/fileSource_class.f90): rows[n_rows][10] are the records of a printToFile dump
 * (r, dir, E, G, broodID, wgt; particleDungeon_class.f90:1077-1112), kept on the device by sb_set_file_source.
 * sb_source_file = fileSource%sampleParticle for n particles (:151-196): row int(rand * n_rows) + 1, position checked
 * against OUTSIDE / undefined regions, weight and E (or G) from the row.                                                      */
/* materialSource (ParticleObjects/Source/materialSource_class.f90:136-210): uniform points of the box [bottom, top] by rejection on the
 * material at the point (at most 200 attempts of 4 random numbers each), isotropic direction, weight 1, E or G as given.                */
typedef struct sb_material_source { int32_t mat_idx, is_mg, G, pad; double E, bottom[3], top[3]; } sb_material_source;
int sb_source_material(sb_engine* h, int n, uint64_t rng_state, int history_offset, const sb_material_source* s);
int sb_geometry_bounds(sb_engine* h, double* bounds6);      /* geometry%bounds(): AABB of the boundary surface, { min[3], max[3] } */
int sb_set_file_source(sb_engine* h, int64_t n_rows, const double* rows, int is_mg);
int sb_source_file(sb_engine* h, int n, uint64_t rng_state, int history_offset);

/* ---- the cycle ------------------------------------------------------------------------------
 * Transports every history of the current bank to its death (transport + collide + tallies +
 * fission-site banking), closes the cycle's tallies (reduceBins, closeCycle with normalisation,
 * k estimators) and leaves the brood-sorted next-cycle bank on the device.
 * rng_state = state of the package RNG at cycle start (pRNG); history n of this engine uses
 * rng_state skipped by 152917*(history_offset+n), eigenPhysicsPackage_class.f90:216-218.       */
int sb_run_cycle(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase, sb_cycle_result* res);
/* particleDungeon%normSize_Repr(totPop, pRNG) on the next-cycle bank, then swap banks.
 * rng_state = pRNG state after the stride(totalPop+1) of the cycle.                            */
int sb_resample(sb_engine* h, int tot_pop, uint64_t rng_state);

/* sb_run_cycle followed by sb_resample with ONE host synchronisation (single rank): nothing the host decides lies between
 * reportCycleEnd and normSize_Repr (eigenPhysicsPackage_class.f90:254-275) except the deterministic stride of pRNG.
 * rng_state_resample = pRNG state after stride(totalPop + 1).                                                    */
int sb_run_cycle_resample(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase, int tot_pop,
                          uint64_t rng_state_resample, sb_cycle_result* res);

/* The same for a caller that keeps its dungeons in host memory (what bench.py's `e2e` figure times): upload of the bank (n sites:
 * r, dir [n][3], w, and G - or E for continuous energy, G then NULL), the cycle, normSize_Repr, and the read-back of the normalised
 * bank (tot_pop sites) and of the cycle's BIN column (bins_out, may be NULL; sb_tally_size doubles), enqueued together and
 * synchronised ONCE. Output arrays may be the input arrays.                                                        */
int sb_run_cycle_resample_host(sb_engine* h, int n, const double* r, const double* dir, const double* w, const int32_t* G, const double* E,
                               uint64_t rng_state, int history_offset, double k_eff, int phase, int tot_pop, uint64_t rng_state_resample,
                               int* n_out, double* r_out, double* dir_out, double* w_out, int32_t* G_out, double* E_out, double* bins_out,
                               sb_cycle_result* res);

/* ---- the cycle split for several ranks (one engine = one GPU = one rank) ---------------------
 * Ranks own contiguous shares of the bank (getWorkshare/getOffset, SharedModules/mpi_func.f90:133-159) and
 * exchange once per cycle what SCONE exchanges over MPI; the transport between ranks (NCCL, CUDA-aware MPI,
 * peer copies) is the caller's, the engine only exposes device buffers:
 *   sb_cycle_begin : transport + brood order; writes this rank's 6 score sums {implicit production, implicit
 *                    absorption, analog leakage, scatter production, start weight, end weight}, its fission-bank size
 *                    (as a double) and a zero to dev_sums (device pointer, 8 doubles) -> caller all-reduces the sums /
 *                    all-gathers the 8 doubles (scoreMemory%reduceBins with
 *                    mpiSync, scoreMemory_class.f90:404-431; mpi_bcast of k, eigenPhysicsPackage_class.f90:302)
 *   sb_cycle_end   : k estimators and closeCycle from the reduced sums (identical on every rank)
 *   sb_resample_ranked : normSize_Repr with the bank sizes of all ranks (replaces mpi_gather + 3 mpi_bcast,
 *                    particleDungeon_class.f90:464,516-518: the threshold is recomputed on every rank from
 *                    master_rng_state = state of the MASTER's pRNG); returns the new local size
 *   sb_bank_export / sb_bank_splice : loadBalancing (particleDungeon_class.f90:607-698): pack k sites of the
 *                    front / back of the bank into device buffers of sb_site_buffer_bytes(k) bytes; then rebuild
 *                    the bank as [add_front] + bank[drop_front : n - drop_back] + [add_back]                */
int sb_cycle_begin(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase, double* dev_sums, int32_t* n_sites);
int sb_cycle_end(sb_engine* h, const double* dev_sums, sb_cycle_result* res);
/* sb_cycle_end + sb_resample_ranked with one synchronisation; host_sums = the 6 reduced sums in HOST memory; new_sizes[n_ranks] =
 * bank size of every rank after normalisation (each rank computes all of them: no second all-gather)                */
int sb_cycle_end_resample_ranked(sb_engine* h, const double* host_sums, int tot_pop, uint64_t master_rng_state, int n_ranks, int rank,
                                 const int32_t* pop_sizes, int32_t* new_sizes, sb_cycle_result* res);
int sb_resample_ranked(sb_engine* h, int tot_pop, uint64_t master_rng_state, int n_ranks, int rank, const int32_t* pop_sizes, int32_t* new_local_pop);
size_t sb_site_buffer_bytes(int k);

/* ---- the ranks of one node over peer memory (NVLink / NVSwitch) --------------------------------
 * The same cycle with no host and no library collective in the loop: every rank owns a small mailbox + site staging buffers in
 * its own HBM, exported with CUDA IPC; the other ranks' kernels store into it (score sums, bank sizes, the sites loadBalancing
 * hands to the neighbours) and release a flag, the owner's kernels spin on their local flags.  Replaces, per cycle, the
 * mpi_reduce / mpi_bcast of scoreMemory%reduceBins, the gathers and broadcasts of normSize_Repr (particleDungeon_class.f90:464,
 * 516-518,593) and the mpi_send / mpi_recv of loadBalancing (:607-698).
 *   sb_peer_create  after sb_set_options: allocates the region, returns its 64-byte cudaIpcMemHandle_t; stage_cap (sites) is the
 *                   same on every rank and at least the largest sb_bank_capacity of the ranks
 *   sb_peer_attach  handles of all ranks in rank order (the caller exchanges them once, e.g. MPI_Allgather); caps[r] = sb_peer_capacity
 *                   (the stage_cap) of rank r (must be equal), may be NULL
 *   sb_run_cycle_ranked_peer  = sb_cycle_begin + reduction + sb_cycle_end_resample_ranked + loadBalancing with ONE host
 *                   synchronisation; final_sizes[n_ranks] = every rank's bank size afterwards.  All ranks must call it for the
 *                   same cycles; a rank that does not show up within the timeout (default 20 s) gives SB_ERR_PEER_TIMEOUT.     */
int sb_bank_capacity(sb_engine* h);          /* sites the banks of this engine hold: 2 * max_pop, the dungeons of eigenPhysicsPackage_class.f90:355-356 */
int sb_peer_create(sb_engine* h, int n_ranks, int rank, int stage_cap, void* ipc_handle_64_bytes);
int sb_peer_attach(sb_engine* h, const void* ipc_handles, const int32_t* caps);
int sb_peer_capacity(sb_engine* h);
int sb_peer_set_timeout(sb_engine* h, double seconds);
int sb_run_cycle_ranked_peer(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase, int tot_pop,
                             uint64_t master_rng_state_resample, int32_t* final_sizes, sb_cycle_result* res);
int sb_bank_export(sb_engine* h, int k_front, void* dev_buf_front, int k_back, void* dev_buf_back);
int sb_bank_splice(sb_engine* h, int drop_front, int drop_back, int add_front, const void* dev_buf_front, int add_back, const void* dev_buf_back);

/* ---- results (scoreMemory) ---------------------------------------------------------------- */
int64_t sb_tally_size(sb_engine* h, int phase);
int sb_tally_read(sb_engine* h, int phase, double* csum, double* csum2, int32_t* batch_n);
int sb_tally_last_bins(sb_engine* h, int phase, double* bins);   /* BIN column of the last closed cycle, before normalisation */

/* ---- measurement ------------------------------------------------------------------------------
 * sb_timer_*: CUDA events on the engine's stream around whatever is issued in between.
 * sb_profile_*: per-launch CUDA-event time of the history kernel, accumulated over cycles.          */
int sb_timer_begin(sb_engine* h);
int sb_timer_end(sb_engine* h, double* ms);
int sb_profile_enable(sb_engine* h, int on);
int sb_profile_read(sb_engine* h, double* ms_histories, int64_t* n_launches, int64_t* n_segments, int64_t* n_scores);
/* ranked cycles over peer memory, while profiling is enabled: accumulated device time [ms] between the end of the brood
   ordering and the arrival of every rank's sums (waiting for the slowest rank), and of the cycle close + normSize_Repr +
   load balancing that follow */
int sb_profile_peer_stages(sb_engine* h, double* ms_wait, double* ms_tail);
int sb_flush_l2(sb_engine* h, size_t bytes);
/* page-locked host memory for the banks the caller keeps (so that uploads/downloads are true async DMA) */
void* sb_pinned_alloc(size_t bytes);
void  sb_pinned_free(void* p);

/* ---- continuous-energy cross-section lookup (the XS-lookup event kernel) -----------------------
 * sb_load_ce_data builds the unionised grid + majorant (initMajorant, aceNeutronDatabase_class.f90:1330-1621) and the
 * per-nuclide index table on the device. sb_ce_lookup: for n particles (E [MeV], matIdx 1-based) any of
 *   total    [n]    Sigma_t of the material (updateTotalMatXS / getTrackMatXS, :402-443,509-571)
 *   macro    [n][8] full neutronMacroXSs set (updateMacroXSs, :579-644), order as mainData rows
 *   majorant [n]    unionised majorant at E (updateMajorantXS, :346-394)
 * NULL outputs are skipped. Host pointers: copied in and out (the end-to-end path); *_device: device pointers, no copies.
 * sb_ce_nuclide_index: the index aceNeutronNuclide%search returns for nuclide nuc_idx (parity: bit-exact).        */
int sb_load_ce_data(sb_engine* h, const sb_ce_flat* d);
int sb_ce_union_size(sb_engine* h);
int sb_ce_union(sb_engine* h, double* grid, double* majorant);
int sb_ce_lookup(sb_engine* h, int64_t n, const double* E, const int32_t* mat, double* total, double* macro, double* majorant);
int sb_ce_lookup_device(sb_engine* h, int64_t n, const double* dE, const int32_t* dMat, double* dTotal, double* dMacro, double* dMajorant);
int sb_ce_nuclide_index(sb_engine* h, int nuc_idx, int64_t n, const double* E, int32_t* idx);
/* the same with the lookups binned by (material, energy) first - the north star's "sorting by material and energy": a counting
 * sort of the lookup numbers on the device (histogram, scan, scatter), then the same lookup kernel with lane j working on the j-th
 * lookup of that order, so that the lanes of a warp gather from the same sectors.  Results land in the caller's order and are
 * bit-identical to sb_ce_lookup_device's.  sb_ce_last_kernel_ms then covers sort + lookup.                               */
int sb_ce_lookup_sorted_device(sb_engine* h, int64_t n, const double* dE, const int32_t* dMat, double* dTotal, double* dMacro, double* dMajorant);
/* CUDA-event time of the last sb_ce_lookup_device launch, in ms */
int sb_ce_last_kernel_ms(sb_engine* h, double* ms);
/* device memory of the loaded CE tables: raw_bytes = the nuclide grids and mainData as the reference holds them; index_bytes = what
 * the engine adds to find and interpolate them (pair records of the total cross section, the per-nuclide hashed index, the
 * unionised grid and majorant, and - only while it is small, SB_CE_IDXTAB_MAX_MB, default 256 - the [union interval][nuclide]
 * index table; has_union_table tells whether that one was built). */
int sb_ce_memory(sb_engine* h, int64_t* raw_bytes, int64_t* index_bytes, int32_t* has_union_table);

/* ---- batch queries used by the parity tests (same device functions as the cycle kernel) ---- */
/* placeCoord / whatIsAt for n points; if dist != NULL teleport by dist[i] first (geometryStd teleport) */
int sb_geom_query(sb_engine* h, int64_t n, double* r, double* dir, const double* dist, int32_t* mat, int32_t* unique_id);
int sb_mg_query(sb_engine* h, int64_t n, const int32_t* mat, const int32_t* G, double* total, double* majorant);
int sb_rng_query(int64_t n, const uint64_t* state, const int64_t* skip, uint64_t* out_state, double* out_real);
int sb_math_query(int64_t n, const double* x, double* log_x, double* sin_x, double* cos_x);
/* self-check of the engine's in-line division / square root (the sequences nvcc expands `/` and sqrt into, without
   their per-operation range branch) against the plain operators on n generated operand pairs with binary exponents in
   [-exp_span, exp_span]: *mismatches = results that differ in any bit (must be 0). */
int sb_fastmath_check(int64_t n, uint64_t seed, int exp_span, int64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif
