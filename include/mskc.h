/*
 * mskc.h -- C ABI of the B200-native MulSKIPS KMC event-loop library (libmskc.so).
 *
 * The reference (MulSKIPS, /root/reference) has no FFI for this path: the loop
 * body lives in the Fortran host (mulskips-source/main.f:304-371) and works on
 * module globals (mulskips-source/modules/defsystem.f).  These entry points sit
 * exactly where that loop and the Set* -> AddMC/AddBulk/AddVoid calls sit, so a
 * Fortran host binds them with one ISO_C_BINDING interface block
 * (fortran/mskc_bind.f90).
 *
 * Conventions: extern "C", plain pointers and sizes, caller owns host buffers,
 * the library owns device memory behind an opaque handle, every call returns
 * int (0 = ok, <0 = error; mskc_last_error() gives the text, which mirrors the
 * reference's STOP messages where one exists).  One handle per GPU, one host
 * thread per handle (the reference is single-threaded).  No CPU fallback: if
 * no CUDA device is usable mskc_create fails.
 */
#ifndef MSKC_H
#define MSKC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mskc_ctx mskc_ctx;

/* event codes of the PVD/CVD tree (main.f:335-357): 1..3 deposit species k, 4 evaporate,
 * 5 desorb, 5+k adsorb coverage species k */
enum { MSKC_EV_DEP1 = 1, MSKC_EV_DEP2 = 2, MSKC_EV_DEP3 = 3, MSKC_EV_EVAP = 4, MSKC_EV_DES = 5, MSKC_EV_ABS1 = 6 };

/* why mskc_run_* returned */
enum {
  MSKC_STOP_BUDGET = 0,   /* event_budget / n_sweeps exhausted; call again to continue            */
  MSKC_STOP_MAXITER = 1,  /* Iter > MaxIter            (main.f:362-363)                           */
  MSKC_STOP_MAXTIME = 2,  /* Time > MaxTime            (main.f:364-365)                           */
  MSKC_STOP_FRAME = 3,    /* output boundary reached AFTER the pick/time update and BEFORE the
                             event is applied (main.f:316-333): write the frame, call again       */
  MSKC_STOP_FEWMC = 4,    /* "few MC particles: stop MC" (main.f:307-310)                         */
  MSKC_STOP_RESOLIDIFIED = 5, /* LaserAnnealing fork: NumAtoms+NumAdAtomOcc > NumOcc0
                             (mulskips-source-LaserAnnealing/main.f:714-723)                      */
  MSKC_STOP_ERROR = -1
};

/* Replaces: static allocation in modules/defsystem.f:27-68 and mcevent/AllocateArrays.f:24-46.
 * nummcmax<=0 selects the reference's NumMcMax=500000 (defsystem.f:34) -> Levels=20. */
int mskc_create(int lenx, int leny, int lenz, int ncrystal, int ncov,
                const int32_t *listcrystal /*[ncrystal]*/, const int32_t *listcov /*[ncov]*/,
                int device_id, int nummcmax, mskc_ctx **out);
int mskc_destroy(mskc_ctx *h);
/* text of the last error on this handle (h may be NULL: last mskc_create failure) */
const char *mskc_last_error(const mskc_ctx *h);

/* Replaces: the tables filled by mcevent/SetProbability.f:21-143.  Fortran column-major
 * flattening, exactly the module arrays: PtransE(3,0:3,0:3,0:3,0:3,0:3,0:3) = 12288 doubles,
 * PtransD(3,1:3) = 9, PtransAbs(3,0:3) = 12, PtransDes(3,0:3) = 12 (defsystem.f:60-65). */
int mskc_set_rates(mskc_ctx *h, const double *PtransE, const double *PtransD,
                   const double *PtransAbs, const double *PtransDes, double PtransZig);
/* Replaces: the LA rate law of mcevent/Get_Prob.f:127-149 (InitSt='LA').  Per-temperature tables are
 * built on the host with libm exp/erf (bit-identical to the CPU reference) for every integer kelvin
 * in [tmin,tmax].  enable=0 switches back to plain tables. */
int mskc_set_la(mskc_ctx *h, int enable, double Tm, double P0, double sigma, double Tflex, int tmin, int tmax);

/* Replaces: the 'ST' tabulated-temperature mode -- mcevent/SetProbabilityFromField.f:21-171 (ntemp blocks of the four
 * tables at the temperatures reftemp[], Fortran flattening per block as in mskc_set_rates: refPE [ntemp][12288],
 * refPD [ntemp][9], refPAbs / refPDes [ntemp][12], NULL = zeros) and mcevent/InterpProbFromField.f:22-57 (linear
 * interpolation with INTEGER distance weights at the site's temperature, Get_Prob.f:76-91).  The interpolated tables
 * are built on the host for every integer kelvin in [tmin,tmax] (must lie inside reftemp[0]..reftemp[ntemp-1], where
 * the reference would index out of bounds otherwise), so device rates are the reference's bit for bit.
 * fixed_temperature > 0 is the Ttop == Tbottom case (initial/SetTrenchSi.f:65-67): interpolate once at that
 * temperature and use plain tables.  Either way a cleared cell keeps its temperature bits (Evaporation.f:70-74). */
int mskc_set_rate_field(mskc_ctx *h, int ntemp, const int32_t *reftemp, const double *refPE, const double *refPD,
                        const double *refPAbs, const double *refPDes, double PtransZig, int tmin, int tmax, int fixed_temperature);

/* The older LaserAnnealing fork (mulskips-source-LaserAnnealing, "la/"; what examples/ex4-LA-SiGe drives through
 * sockets) differs from mulskips-source in three ways this library honours:
 *  (1) layout: species at bits 28-29 (la/modules/defsystem.f:99-100) and event codes 1-2 deposit, 3 evaporate,
 *      4 desorb, 4+k adsorb (la/main.f:671-699).  mskc_set_tree_variant(h, MSKC_TREE_LA) makes every entry point
 *      that carries lattice words or event codes (upload_state, download_lattice, patch_lattice, download_lists,
 *      download_active, replay_events, extract_frame) speak that layout; the device keeps one internal layout.
 *  (2) rate law (la/mcevent/Get_Prob.f:142-152): Tm, P0, sigma, Tflex per crystal species and the solidification
 *      rate scaled by the liquid fraction FLOAT(CountLiquid(k))/SUM(CountLiquid), CountLiquid = InitCountCrystal -
 *      CountCrystal re-evaluated before every event (la/main.f:636).  Factors come from host-built tables (glibc
 *      exp/erf) and are multiplied in the reference's order, so serial-parity rates are bit-identical.
 *  (3) loop (la/main.f:675-681,714-723): in a Si/Ge melt a deposition of an exhausted species becomes the other one;
 *      the run stops (MSKC_STOP_RESOLIDIFIED) once NumAtoms + NumAdAtomOcc exceeds num_occ0.  num_adatom_occ is the
 *      builder's count of occupied ad-atoms (la/initial/SetCAD.f:544,595); the serial-parity loop updates it where
 *      la/mcevent/deposition.f:219,246,253 and Evaporation.f:47,192 do, mskc_get_counters returns it.
 * The sampler (mskc_run_parallel) uses CountLiquid as of the start of the call and derives the occupied count from a
 * lattice census; call it with few sweeps per call when the melt composition changes quickly.
 * Call after mskc_set_rates (and again after a new mskc_set_rates); enable = 0 switches the law off. */
enum { MSKC_TREE_SRC = 0, MSKC_TREE_LA = 1 };
int mskc_set_tree_variant(mskc_ctx *h, int variant);
int mskc_set_la_fork(mskc_ctx *h, int enable, const double *Tm, const double *P0, const double *sigma, const double *Tflex /* [ncrystal] each */,
                     int tmin, int tmax, const int32_t *InitCountCrystal /* [3] */, int num_occ0, int num_adatom_occ);

/* Replaces: what the Set* builders leave in module globals (LattCoo + the three particle lists, in
 * REFERENCE LIST ORDER, built through mcevent/Add.f).  Coordinates are absolute, neighbours are
 * [n][4][3] with (0,0,0) = "not fixed" (defdertype.f:32-37).  LattCoo is [(lenz+1)][leny][lenx]. */
int mskc_upload_state(mskc_ctx *h, const int32_t *LattCoo,
                      int n_mc, const int32_t *mc_xyz, const int32_t *mc_nn, const int32_t *mc_event, const double *mc_rate,
                      int n_bulk, const int32_t *bulk_xyz, const int32_t *bulk_nn,
                      int n_void, const int32_t *void_xyz, const int32_t *void_nn);
/* Device-side flat (001) builder for lattices too large to stage through the host: same geometry as
 * initial/SetSi.f (walls=1, one species) / initial/SetSiC3C.f (walls=0, Si/C), parallel mode only
 * (list order is not the reference's). */
int mskc_init_flat(mskc_ctx *h, int len1, int walls, int two_species);

/* Replaces: IDUM / counter / the SAVEd state of mcevent/random.f:32-34.  runtype is 'R' or 'T'
 * (PickTreeEvent.f:43-49); ran3_state = {iff, inext, inextp, ma[1..55]} (58 ints) or NULL for a
 * fresh generator seeded by idum; counter0 = numberslist cursor (definitions.f:35). */
int mskc_seed(mskc_ctx *h, int idum, int runtype, int64_t counter0, const int32_t *ran3_state);
/* read the generator back (58 ints + counter) so a host can continue the stream */
int mskc_get_rng(mskc_ctx *h, int32_t *ran3_state, int64_t *counter, int32_t *idum);
/* Iter/Time/species counters of main.f:31,54,75-76 (for continuation) */
int mskc_set_clock(mskc_ctx *h, int64_t iter, double time, const int32_t *CountCrystal, const int32_t *CountCov);

typedef struct {
  int64_t max_iter;      /* MaxIter                                        */
  double  max_time;      /* MaxTime                                        */
  int32_t exit_on_time;  /* ExitStrategy=='Time'                           */
  int32_t pad;
  int64_t out_every;     /* OutMolMol ('Iter' strategy), 0 = no frames     */
  double  out_time;      /* OutTime ('Time' strategy), 0 = no frames       */
  int64_t framecounter;  /* frames already written in 'Time' strategy      */
  int64_t event_budget;  /* events to execute in this call at most, 0=inf  */
} mskc_parity_args;

typedef struct {
  int64_t iter;          /* Iter (1-based index of the NEXT event; main.f:31,359) */
  int64_t iter_out;      /* IterOut                                               */
  double  time;          /* Time                                                  */
  int32_t last_site[3];  /* site of the last picked event                         */
  int32_t stop_reason;
  int32_t num_atoms, num_voids, num_adatom;
  int32_t pad;
  int64_t events_done;   /* events executed by this call                          */
  int64_t site_updates;  /* Get_Prob evaluations by this call                     */
  double  tree_root;     /* Tree(1)                                               */
  float   kernel_ms;     /* device time of this call (CUDA events on the library's stream) */
  int32_t launches;      /* kernels launched by this call */
} mskc_parity_result;

/* Replaces: the DO loop body main.f:304-371, serial-parity mode (single GPU): replays the reference's
 * n-fold-way selection, RNG streams, list order and tree updates; lattice words, event sequence and
 * Time are bit-identical to the reference. */
int mskc_run_parity(mskc_ctx *h, const mskc_parity_args *a, mskc_parity_result *r);

typedef struct {
  int64_t n_sweeps;          /* sweeps (4 colour steps each) to run in this call                    */
  double  events_per_tile;   /* target mean events per active tile per colour step (sets tau)       */
  double  tau_fixed;         /* >0: use this time quantum instead of the adaptive one (from the first sweep of this call on) */
  uint64_t seed;             /* Philox key                                                          */
  int32_t  tile;             /* tile edge in super-lattice units (24), 0 = default                  */
  int32_t  flags;            /* bits 0-15: events one tile may execute per colour step before the rest of its
                              * quantum is carried to its next step (0 = default 3*events_per_tile,
                              * 0xFFFF = no cap); bit 16: MSKC_PAR_TIME_KERNELS; other bits reserved, 0   */
} mskc_parallel_args;
/* time every colour-step launch with CUDA events on the library's stream (n_sweeps <= 64): sweep_kernel_ms of the result */
#define MSKC_PAR_TIME_KERNELS 0x10000

typedef struct {
  int64_t events_done;       /* events executed by this call                                        */
  int64_t site_updates;      /* rate re-evaluations by this call                                    */
  int64_t sweeps_done;
  double  time;              /* KMC time after the call                                             */
  double  total_rate;        /* sum of all rates at the last sweep                                  */
  int64_t n_active;          /* active (listed) sites at the last sweep                             */
  int64_t count_dep[3], count_evap, count_abs, count_des;
  float   kernel_ms;         /* device time of this call (CUDA events on the library's stream)      */
  int32_t launches;          /* kernels launched by this call                                       */
  float   sweep_kernel_ms;   /* part of kernel_ms spent in the sampler kernels (colour steps / sweep batches) */
  int32_t stop_reason;       /* MSKC_STOP_BUDGET (the sweeps asked for were run) or MSKC_STOP_FEWMC (main.f:307-310) */
  int32_t sampler_launches;  /* how many sampler kernels those were (short calls time each one)     */
  int32_t max_tile_events;   /* diagnostic: most events one tile executed in one colour step        */
  double  tau;               /* the adaptive quantum (events_per_tile / mean rate of a tile that carries rate) as of the last sweep */
} mskc_parallel_result;

/* Synchronous-sublattice sampler: x-y tiles coloured 2x2, every tile of one colour runs its own
 * n-fold way for a time quantum tau with counter-based Philox RNG.  Statistically equivalent to the
 * loop of main.f:304-371 (see DESIGN.md for the tolerances). */
int mskc_run_parallel(mskc_ctx *h, const mskc_parallel_args *a, mskc_parallel_result *r);
/* Dense re-rate pass ("rate kernel"): recompute (event class, rate) of every listed site from the
 * packed lattice and rebuild the per-tile sums.  Needed after a temperature upload. */
int mskc_rerate_all(mskc_ctx *h, int64_t *n_sites, double *total_rate, float *kernel_ms);

/* Replaces: initial/SetT.f:25-50 -- integer kelvin per cell, [lenz][leny*lenx], value 10 = wall
 * (left untouched).  Writes bits 9..21 of every lattice word on the device. */
int mskc_set_temperature(mskc_ctx *h, const int32_t *T);
/* the same field as 16-bit kelvin (the lattice word holds 13 bits): half the bytes over PCIe per coupling cycle */
int mskc_set_temperature_u16(mskc_ctx *h, const uint16_t *T);
/* Replaces: la/main.f:797-817 (the `c` array the socket's GETPHASES returns): per cell [lenz][leny*lenx] the coordination
 * of an occupied cell, 10 for a wall, 0 otherwise -- computed on the device, one byte per cell comes down instead of
 * the whole lattice */
int mskc_download_phases(mskc_ctx *h, uint8_t *phases);
/* device time (ms, CUDA events) of the dense kernel the last of the three calls above launched */
int mskc_last_dense_ms(mskc_ctx *h, float *ms);

/* Replaces: the module globals read by inpout/WriteMolMolXYZFile.f, the checkpoint writer
 * (main.f:385-403) and the Coor dump (main.f:407-430). */
int mskc_download_lattice(mskc_ctx *h, int32_t *LattCoo);
/* patch lattice words changed by the host (the _w.xyz pass of WriteMolMolXYZFile.f:267-398 flips
 * species bits in place): n (cell index, word) pairs */
int mskc_patch_lattice(mskc_ctx *h, int n, const int64_t *cell, const int32_t *word);
typedef struct {
  int32_t num_atoms, num_voids, num_adatom;
  int32_t num_adatom_occ;     /* LaserAnnealing fork: occupied ad-atoms (NumAdAtomOcc), 0 otherwise */
  int32_t CountCrystal[3], CountCov[3];
  int64_t iter; double time;
} mskc_counters;
int mskc_get_counters(mskc_ctx *h, mskc_counters *c);
/* list download; pass NULL for what is not wanted.  Sizes from mskc_get_counters.
 * MC and void lists come in reference list order (parity mode); bulk in z,y,x scan order. */
int mskc_download_lists(mskc_ctx *h,
                        int32_t *mc_xyz, int32_t *mc_nn, int32_t *mc_event, double *mc_rate,
                        int32_t *void_xyz, int32_t *void_nn,
                        int32_t *bulk_xyz, int32_t *bulk_nn, int64_t bulk_cap, int64_t *n_bulk_found);
/* 64-bit FNV-1a over the lattice words + list sizes + bits of Time + Iter: same function as the
 * oracle's orc_state_hash, computed on the device lattice */
int mskc_state_hash(mskc_ctx *h, uint64_t *hash);
/* device census for output/statistics: counts used by WriteMolMolXYZFile.f:77-119 */
typedef struct { int64_t undercoordinated, coverage, noepi_bulk, bulk, occupied, walls; double mean_height, sq_height; } mskc_census;
int mskc_census_lattice(mskc_ctx *h, mskc_census *c, float *kernel_ms);

/* ---- several GPUs: x-decomposed slabs with halo exchange (no counterpart in the reference, which is one
 * serial process).  The global box global_lenx x leny x lenz is cut into nranks blocks along x;
 * every rank's context holds its block plus one 24-unit tile column of each neighbour (local LenX =
 * global_lenx/nranks + 48, periodic across the ranks).  mskc_run_parallel then exchanges the changed records
 * of the boundary strips after every colour step (NCCL send/recv) and all-reduces the total rate once per
 * sweep.  With nranks == 1 the rank is its own neighbour: the same pack/apply kernels, no NCCL.  Only the
 * sampler runs on a slab (serial parity mode is single-GPU by construction).
 *   rank 0: mskc_comm_unique_id(id)  ->  the host broadcasts the 128 bytes  ->  all: mskc_comm_connect(h, id). */
int mskc_create_slab(int global_lenx, int leny, int lenz, int ncrystal, int ncov,
                     const int32_t *listcrystal, const int32_t *listcov,
                     int rank, int nranks, int device_id, mskc_ctx **out);
int mskc_comm_unique_id(void *id128);
int mskc_comm_connect(mskc_ctx *h, const void *id128);

/* Test/diagnostic entry points of the sampler's transition code (no counterpart in the reference).
 * mskc_replay_events applies n GIVEN events one after the other through the same warp-cooperative
 * transition + re-rate code mskc_run_parallel uses: site_xyz [n][3], event codes as in main.f:335-357,
 * arm[i] = 0 let IF_Arm_Choice.f decide (ties -> zig-zag), 1 force zig-zag, 2 force arm-chair.  tree_root
 * stands for Tree(1) in the isolated-atom rate (Get_Prob.f:305).  tests/ replay the CPU oracle's event
 * trace with it and compare lattice words, bonds and rates bit for bit. */
int mskc_replay_events(mskc_ctx *h, int64_t n, const int32_t *site_xyz, const int32_t *event, const int32_t *arm, double tree_root);
/* the same trace through the instantiation mskc_run_parallel launches: four 8-lane teams of one warp in lock step,
 * feature set (coverage / per-kelvin tables) as the sampler picks it.  Consecutive events more than 48 units apart
 * in x or y touch disjoint cells and share a lock-step iteration (up to four), so teams run concurrently and idle
 * exactly as tiles do in a colour step; *n_iterations / *n_shared = iterations run / iterations with >= 2 events. */
int mskc_replay_teams(mskc_ctx *h, int64_t n, const int32_t *site_xyz, const int32_t *event, const int32_t *arm, double tree_root, int64_t *n_iterations, int64_t *n_shared);
/* the sampler's active-site lists, unordered: site, event class (1..5 as main.f; 100 = deposit/adsorb with
 * the species drawn when the event fires), rate.  cap = capacity of the arrays, *n = entries found. */
int mskc_download_active(mskc_ctx *h, int64_t cap, int32_t *xyz, int32_t *event, double *rate, int64_t *n);
/* the four stored bond targets (absolute coordinates, (0,0,0) = not fixed) of n given cells */
int mskc_download_bonds(mskc_ctx *h, int64_t n, const int32_t *xyz, int32_t *nn);

/* Frame extraction on the device: (cell index, lattice word) of every cell the reference's frame writers print
 * (inpout/WriteMolMolXYZFile.f:130-215: under-coordinated atoms and coverage sites for .xyz, non-epitaxial
 * 4-coordinated atoms for _d.xyz), so that a frame costs one 4 B/cell streaming pass and a download of the
 * surface only instead of the whole lattice.  Unordered; sort by cell for the reference's z,y,x order.
 * cap = capacity of the two arrays (host memory), *n = cells found (may exceed cap: call again with more). */
int mskc_extract_frame(mskc_ctx *h, int64_t cap, uint32_t *cell, uint32_t *word, int64_t *n, float *kernel_ms);
/* The complete frame without the lattice, for the host program's frame writer (inpout/WriteMolMolXYZFile.f:77-260): the
 * cells above plus every cell one of the writer's tests touches -- coverage ids 1 / 17 whatever else the word holds
 * (:169-176), atoms of species 1..3 with coordination <= 3 (:177-215), the cells of the two counting scans (:77-92) and
 * the empty 4-coordinated sites, i.e. the vacancies of _v.xyz (Add.f:82-100 lists a site exactly when it reaches four
 * neighbours while empty).  Same calling convention as mskc_extract_frame; the plane ranges of the individual scans are
 * the caller's (it has the coordinates). */
int mskc_extract_frame_all(mskc_ctx *h, int64_t cap, uint32_t *cell, uint32_t *word, int64_t *n, float *kernel_ms);
/* The antisite scan of the SiC frame writer (inpout/WriteMolMolXYZFile.f:267-398) on the device: (cell, word) of every
 * 4-coordinated atom of planes 1..LenZ-1 whose number of bonded neighbours of the other species equals target, 0 (the
 * atoms the reference flips, :330-349) or 3 (the atoms it prints into _w.xyz, :352-372).  Unordered, as above. */
int mskc_extract_antisites(mskc_ctx *h, int target, int64_t cap, uint32_t *cell, uint32_t *word, int64_t *n, float *kernel_ms);
/* lattice words of n given cells (cell = x + LenX*(y + LenY*z)): what a writer needs of the neighbours of a vacancy
 * (WriteMolMolXYZFile.f:233-256) without downloading the lattice */
int mskc_gather_lattice(mskc_ctx *h, int64_t n, const int64_t *cell, int32_t *word);

/* diagnostics: the sampler keeps the neighbour census of every site incrementally (two flag bits per bond: species of
 * the occupied, reciprocal neighbour -- what Get_Prob.f:216-291 recounts at every call).  This recomputes the census of
 * every cell that carries bonds in one dense pass and counts the cells whose stored flags differ (must be 0), and the
 * cells whose coordination differs from their census (the reference's own consistency stop, Get_Prob.f:293). */
int mskc_check_flags(mskc_ctx *h, int64_t *n_sites, int64_t *n_flag_mismatch, int64_t *n_coor_mismatch, float *kernel_ms);

/* diagnostics: raw per-tile counters of the sublattice sampler (80 B records, colour-major list order) */
int mskc_debug_tile_stats(mskc_ctx *h, void *out, int64_t max_tiles, int64_t *ntiles);

#ifdef __cplusplus
}
#endif
#endif
