/* b200md.h -- C ABI of the B200-native force path for the LAMMPS pair styles
 * `rebomos` (REBO Mo/S + tapered LJ) and `aeam` (angular EAM).
 *
 * Plain C, plain pointers and sizes, no C++/torch types.  The library behind it
 * (lammps_plugins_b200/libb200md.so) is hand-written CUDA for sm_100a; there is
 * no CPU fallback: every entry point returns B200MD_ERR_CUDA when no device or
 * kernel image is usable, and the host pair classes turn that into error->one().
 *
 * What each entry point replaces in the reference (lammps/lammps-plugins):
 *
 *   b200md_rebomos_init        PairREBOMoS::read_file/init_one parameter block
 *                              USER-REBOMOS/pair_rebomos.cpp:857-1107, 244-274
 *   b200md_set_neighbor_list   the NeighList handed to Pair::init_list / used at
 *                              pair_rebomos.cpp:304-307, 476-479; pair_aeam.cpp:150-153
 *   b200md_rebomos_compute     PairREBOMoS::compute           pair_rebomos.cpp:102-111
 *                                = REBO_neigh :281-352 + FREBO :358-447
 *                                + bondorder :571-847 + FLJ :453-558
 *                                + Pair::virial_fdotr_compute (LAMMPS-core)
 *   b200md_rebomos_neigh       PairREBOMoS::REBO_neigh output (REBO_numneigh,
 *                              REBO_firstneigh, nM, nS)        pair_rebomos.cpp:281-352
 *   b200md_aeam_init           PairAEAM::file2array/array2spline tables
 *                              USER-AEAM/pair_aeam.cpp:752-942
 *   b200md_aeam_compute        PairAEAM::compute               pair_aeam.cpp:110-479
 *   b200md_aeam_get_rho_fp     per-atom rho/fp (the data of pack/unpack_*_comm,
 *                              pair_aeam.cpp:946-990)
 *   b200md_neigh_build         LAMMPS-core NBinStandard + NPairFullBin[Ghost]
 *                              ("pair build: full/bin/ghost", log.rebomos-bulk.1:49)
 *   b200md_system_*            the LAMMPS run loop around compute(): Verlet::run,
 *                              fix nve, Neighbor::decide, CommBrick forward/reverse/
 *                              borders/exchange, thermo (GPU-resident driver)
 *   b200md_local_group_create  the MPI communicator LAMMPS runs in ("2 by 2 by 1 MPI processor
 *   b200md_system_comm_init*   grid", log.rebomos-bulk.4:22): NCCL ranks or in-process ranks
 *
 * Conventions: positions/forces are AoS double[n][3] exactly like atom->x /
 * atom->f; `type` is the 1-based LAMMPS atom type; `tag` the int32 atom ID;
 * neighbor indices are local indices (owned first, then ghosts), the top 3 bits
 * are masked with NEIGHMASK as the reference does.  All functions return 0 on
 * success, a negative B200MD_ERR_* otherwise; b200md_last_error() gives the text.
 */
#ifndef B200MD_H
#define B200MD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MD_VERSION 100

#define B200MD_OK 0
#define B200MD_ERR_CUDA (-1)      /* no device / kernel launch or runtime failure */
#define B200MD_ERR_ARG (-2)       /* bad argument or call order */
#define B200MD_ERR_OVERFLOW (-3)  /* a per-atom row exceeded its capacity */
#define B200MD_ERR_NCCL (-4)

#define B200MD_ENERGY_GLOBAL 1    /* == Pair::ENERGY_GLOBAL */
#define B200MD_VIRIAL_PAIR 1      /* == Pair::VIRIAL_PAIR   */
#define B200MD_VIRIAL_FDOTR 2     /* == Pair::VIRIAL_FDOTR  */

typedef struct b200md_ctx b200md_ctx;

/* ---- lifecycle ----------------------------------------------------------- */
int b200md_create(int device, b200md_ctx **out);
/* number of visible devices this library can drive (all must be sm_100; 0 if none or a mixed box) -- what the host
 * classes take a rank's node-local index modulo of */
int b200md_device_count(void);
void b200md_destroy(b200md_ctx *ctx);
/* text of the last error on this context (ctx may be NULL: last create error) */
const char *b200md_last_error(const b200md_ctx *ctx);
int b200md_version(void);

/* ---- REBOMoS parameters (MoS.REBO.set5b after PairREBOMoS::read_file) ----
 * 2x2 tables are row-major [itype][jtype] with 0 = Mo, 1 = S, exactly the
 * member arrays of pair_rebomos.h:54-60.                                      */
typedef struct {
  double rcmin[4], rcmax[4];
  double Q[4], alpha[4], A[4], BIJc[4], Beta[4];
  double b[7][2];  /* b0..b6  [order][elem]   pair_rebomos.h:56 */
  double bg[7][2]; /* bg0..bg6                pair_rebomos.h:57 */
  double a[4][2];  /* a0..a3                  pair_rebomos.h:58 */
  double rcLJmin[4], rcLJmax[4];
  double epsilon[4], sigma[4];
} b200md_rebomos_params;

/* map[1..ntypes] = 0 (Mo), 1 (S) or -1 (NULL), as PairREBOMoS::coeff builds it */
int b200md_rebomos_init(b200md_ctx *ctx, const b200md_rebomos_params *p, int ntypes, const int *map);

/* ---- AEAM tables (AlSi.aeam after PairAEAM::read_file + file2array) -------
 * Raw tabulated values, 0-based, as stored in the file; the library builds the
 * 7-coefficient splines itself (PairAEAM::interpolate, pair_aeam.cpp:915-942),
 * bit-identically, and keeps them on the device.                              */
typedef struct {
  int nelements;        /* == ntypes (the reference requires it, pair_aeam.cpp:568-572) */
  int nnonangular;      /* types 1..nnonangular are plain EAM, the rest angular */
  const int *nrho;      /* [nelements]                         */
  const double *drho;   /* [nelements]                         */
  const int *nr;        /* [nelements*nelements] row-major i,j */
  const double *dr;     /* [nelements*nelements]               */
  const double *cut;    /* [nelements*nelements]               */
  const double *const *frho; /* [nelements] -> nrho[i] values               */
  const double *const *rhor; /* [nelements*nelements] -> nr[i][j] values    */
  const double *const *z2r;  /* [nelements*nelements] -> nr[i][j] values; only j<=i read */
} b200md_aeam_tables;

int b200md_aeam_init(b200md_ctx *ctx, const b200md_aeam_tables *t);
/* read back one built spline table (testing: bit-parity with array2spline).
 * kind: 0 = frho[i], 1 = rhor[i*nel+j], 2 = z2r[type2z2r]; out[(n+1)*7]      */
int b200md_aeam_get_spline(b200md_ctx *ctx, int kind, int index, double *out, int nrows);

/* ---- neighbor list hand-over (host CSR, LAMMPS NeighList layout) ----------
 * inum owned rows + gnum ghost rows (gnum = 0 for AEAM); row i has numneigh[i]
 * entries at firstneigh[i].  `skin` is neighbor->skin: the list is valid until
 * some atom has moved more than skin/2 (LAMMPS' rebuild rule).  Call again
 * whenever LAMMPS rebuilt the list (neighbor->ago == 0).                      */
int b200md_set_neighbor_list(b200md_ctx *ctx, int inum, int gnum, const int *numneigh,
                             const int *const *firstneigh, double skin);
/* the same list described the way LAMMPS' NeighList really is (neigh_list.h: numneigh[] and firstneigh[] indexed by
 * ATOM index, ilist[0..inum+gnum) naming the atoms that have a row; pair_rebomos.cpp:304-310 walks it that way).  The
 * list must give every owned atom -- and every ghost when gnum > 0 -- exactly one row: skip lists and the sub-style
 * lists of pair hybrid are refused with B200MD_ERR_ARG.  ilist == NULL means the identity. */
int b200md_set_neighbor_list_ilist(b200md_ctx *ctx, int inum, int gnum, const int *ilist, const int *numneigh,
                                   const int *const *firstneigh, double skin);
/* same, from a flat CSR (offsets[inum+gnum+1], values) */
int b200md_set_neighbor_csr(b200md_ctx *ctx, int inum, int gnum, const int64_t *offsets,
                            const int *values, double skin);

/* ---- GPU neighbor build (replaces the host-built list) --------------------
 * Builds the full list (and ghost rows if ghost_rows) on the device from host
 * positions, reproducing LAMMPS' bin/stencil traversal order bit-for-bit.
 * cutneighsq / cutneighghostsq are (ntypes+1)^2 row-major like Neighbor's.
 * box: boxlo[3], boxhi[3], tilt xy xz yz, sub-domain lo/hi (lamda if triclinic),
 * cutghost[3] (Comm::cutghost).                                               */
typedef struct {
  int triclinic;
  double boxlo[3], boxhi[3];
  double xy, xz, yz;
  double sublo[3], subhi[3];
  double cutghost[3];
  double cutneighmax;
} b200md_box;

int b200md_neigh_build(b200md_ctx *ctx, const b200md_box *box, int ntypes, const double *cutneighsq,
                       const double *cutneighghostsq, int nlocal, int nghost, const double *x,
                       const int *type, int ghost_rows, double skin);
/* size / download of the device-resident list (testing and plugin hand-back) */
int b200md_neigh_size(b200md_ctx *ctx, int *nrows, int64_t *nentries);
int b200md_neigh_download(b200md_ctx *ctx, int *numneigh, int64_t *offsets, int *values);

/* ---- REBOMoS compute (host buffers in, host buffers out) ------------------
 * f is ACCUMULATED into (like atom->f after force_clear()); eng_vdwl and
 * virial[6] are overwritten with this call's contribution (caller adds).
 * type[] and tag[] are read on the first call after a neighbor-list hand-over (b200md_set_neighbor_* or
 * b200md_neigh_build) and kept on the device until the next one; x[] is read every call.
 * Forces are complete after LAMMPS' reverse_comm; the split of a pair's force
 * between an owner and its ghost image differs from the reference. */
int b200md_rebomos_compute(b200md_ctx *ctx, int nlocal, int nghost, const double *x, const int *type,
                           const int *tag, int eflag, int vflag, double *f, double *eng_vdwl,
                           double *virial);

/* same, plus per-atom energy eatom[nlocal+nghost] and per-atom virial vatom[(nlocal+nghost)*6] (xx,yy,zz,xy,xz,yz),
 * both ACCUMULATED into like Pair::eatom / Pair::vatom; either may be NULL.  Distribution follows the reference's
 * tallies: ev_tally halves (pair_rebomos.cpp:444,554), v_tally3 thirds (:710), v_tally2 halves (:725).  Ghost entries
 * carry the share of ghost atoms (LAMMPS reverse-communicates them in compute pe/atom, stress/atom). */
int b200md_rebomos_compute_peratom(b200md_ctx *ctx, int nlocal, int nghost, const double *x, const int *type,
                                   const int *tag, int eflag, int vflag, double *f, double *eng_vdwl,
                                   double *virial, double *eatom, double *vatom);

/* REBO short-range sub-list for owned AND ghost atoms from the current list and
 * positions: numneigh[nall], rows packed with stride `stride`, nM[nall], nS[nall] */
int b200md_rebomos_neigh(b200md_ctx *ctx, int nlocal, int nghost, const double *x, const int *type,
                         int stride, int *rebo_numneigh, int *rebo_rows, double *nM, double *nS);

/* ---- AEAM compute ----------------------------------------------------------
 * The pair term is evaluated in gather form (both directed visits of a pair from the owned atom's
 * side), which needs fp = F'(rho) of GHOST neighbors: the fp forward exchange that the reference
 * declares (comm_forward = 1, pair_aeam.cpp:56,307,946-965) but never reads becomes live here.
 *
 * One-shot (single rank; every ghost is a periodic image of an owned atom, matched by atom ID): */
int b200md_aeam_compute(b200md_ctx *ctx, int nlocal, int nghost, const double *x, const int *type,
                        const int *tag, int eflag, int vflag, double *f, double *eng_vdwl, double *virial);
/* Two-phase, for hosts that own the halo exchange (PairAEAM::compute with LAMMPS' Comm):
 *   density phase  = density pass + embedding  (pair_aeam.cpp:164-303): rho[nlocal], fp[nlocal] out
 *   [host: comm->forward_comm(this) fills fp (and rho) of ghosts]
 *   force phase    = force pass                (pair_aeam.cpp:309-478): rho/fp [nlocal+nghost] in     */
int b200md_aeam_density_phase(b200md_ctx *ctx, int nlocal, int nghost, const double *x, const int *type,
                              double *rho, double *fp);
int b200md_aeam_force_phase(b200md_ctx *ctx, const double *rho_all, const double *fp_all, int eflag,
                            int vflag, double *f, double *eng_vdwl, double *virial);
/* per-atom variants (Pair::eatom / Pair::vatom, accumulated into; either may be NULL): embedding energy F to the atom
 * (F/3 for angular atoms, pair_aeam.cpp:295-300), phi/2 of visit (i,j) to i (:389), ev_tally halves (:393), ev_tally3
 * thirds (:472).  In the two-phase form set option "peratom" = 1 BEFORE the density phase. */
int b200md_aeam_compute_peratom(b200md_ctx *ctx, int nlocal, int nghost, const double *x, const int *type,
                                const int *tag, int eflag, int vflag, double *f, double *eng_vdwl, double *virial,
                                double *eatom, double *vatom);
int b200md_aeam_force_phase_peratom(b200md_ctx *ctx, const double *rho_all, const double *fp_all, int eflag, int vflag,
                                    double *f, double *eng_vdwl, double *virial, double *eatom, double *vatom);
/* rho[nlocal], fp[nlocal] of the last compute */
int b200md_aeam_get_rho_fp(b200md_ctx *ctx, int nlocal, double *rho, double *fp);

/* ---- tuning / introspection ------------------------------------------------ */
/* option names (value is an integer):
 *  "deterministic"   0/1: rebomos bond forces are written to a (center, slot) table and summed by destination in a
 *                    fixed order instead of FP64 atomics -- forces are then bitwise reproducible run to run
 *  "margin"          inner-list skin in 1e-3 A, 0 = default skin/2; clamped to skin
 *  "margin_tight"    rebomos, GPU-resident loop: margin of the tight rows the force kernels stream, in 1e-3 A
 *                    (default 400; 0 = no third list level).  Re-derived from the inner rows at margin_tight/2
 *  "lj_pairs"        0/1 (default 1): LJ over pairs of neighboring centers sharing one union row / one row per center
 *  "f_overwrite"     0/1: f is written, not accumulated -- valid when the caller guarantees f == 0 on entry, as right
 *                    after LAMMPS' force_clear()
 *  "h2d_chunks"      plugin mode: pieces of the pipelined position upload (default 6, 1 = one copy, no overlap)
 *  "h2d_ramp"        plugin mode: piece k of the upload holds (ramp + k) / sum_j (ramp + j) of the atoms (default 4: the
 *                    first kernels start sooner, the later launches are longer); 0 = equal pieces
 *  "d2h_chunks"      plugin mode: atom ranges of the pipelined force download (default 4, 1 = one copy)
 *  "d2h_min_atoms"   plugin mode: below this many owned atoms both pipelines are off (default 65536)
 *  "p2p_halo"        0/1 (default 1): multi-GPU halos go through peer-memory windows mapped with CUDA IPC -- the
 *                    sender packs straight into the receiver's HBM over NVLink -- falling back to NCCL send/recv
 *  "aeam_cluster"    AEAM row form: 2 (default) one row per center, the density pass hands f'(r) of every entry to the force
 *                    pass; 1 clusters of 4 consecutive centers share one union row; 0 the force pass re-gathers rho'
 *  "force_rebuild"   resident loop: the NEXT step takes the reneighboring path whatever the displacements (set it on
 *                    every rank; used to time a master rebuild)
 *  "overlap_halo"    0/1 (default 1), resident loop: the owned centers are split into interior and boundary ones and the
 *                    halos run on a stream of their own beside the interior kernels
 *  "split_elems"     2 (default) / 3, resident loop with halo overlap: only the S pair rows (the larger LJ launch) are
 *                    launched as interior + boundary parts, the Mo rows in one launch beside the reverse halo / both elements
 *  "flat_halo"       0/1 (default 1), resident loop on one rank: every ghost is a periodic image of an owned atom, so each
 *                    halo is ONE gather (or fold) through a ghost -> owned-atom map instead of three staged launches
 *  "peer_vote"       0/1 (default 1), resident loop: the per-step reneighbor vote is exchanged through the peer-memory
 *                    windows and reported to a mapped host word (no ncclAllReduce, no stream synchronisation per step);
 *                    0, or no peer windows: ncclAllReduce of one int + copy + cudaStreamSynchronize
 *  "fuse_integrate"  0/1 (default 1), resident loop: the second half kick of a step that is followed by another step (no
 *                    thermo output, no thermostat in between) is applied by the next step's first integrate launch
 *  "one_pass_neigh"  0/1 (default 1), resident loop: master rebuilds after the first walk the stencil once into rows of a
 *                    fixed (sticky) stride; a row longer than the stride falls back to count + scan + fill
 *  "neigh_unroll", "aeam_variant", "aeam_sort_rows"   tuning knobs of measured-and-dropped variants (see profiles/):
 *                    trips of candidates in flight in the row fill when the list cutoff differs by type pair, lane
 *                    layouts of the cluster-row kernels, index-sorted cluster rows
 *  "ang_ctas"        AEAM angular launches: CTAs per SM (default 10)
 *  "fp_gated"        0/1, AEAM two-phase API: the density phase hands out (rho > minrho ? fp : 0) per owned atom, what a
 *                    neighbor needs from it (pair_aeam.cpp:329-332), and rho_out / rho_all may be NULL: the host ships one
 *                    double per ghost and tests nothing itself
 *  "peratom"         0/1, AEAM two-phase API only;  "sync_timing" 0/1: per-launch CUDA events for b200md_kernel_stats */
int b200md_set_option(b200md_ctx *ctx, const char *name, long long value);
/* counters: "kernel_launches", "list_uploads" (master lists received: handed over or built on the device), "compute_calls", "inner_rebuilds", "tight_refreshes", "h2d_bytes", "d2h_bytes",
 * "lj_entries", "short_entries", "num_sms", "p2p_exchanges", "pipelined_calls", "pipelined_redos", "upload_stragglers"
 * (atoms whose positions travel ahead of the pieces of the pipelined upload);
 * row statistics summed on the device when asked (measurement): "master_entries", "lj_entries_tight",
 * "short_entries_tight", "short_entries_owned", "aeam_entries" (-1 while those rows do not exist) */
long long b200md_get_counter(b200md_ctx *ctx, const char *name);
/* device time of the kernels of the last compute call, ms, by name
 * ("rebo_center_mo","rebo_center_s","lj","fdotr","aeam_density","aeam_force",...) */
double b200md_last_kernel_ms(b200md_ctx *ctx, const char *name);
/* accumulated per-kernel device time since the last reset (while "sync_timing" is 1): iterate index from 0
 * until the return value is 1 */
int b200md_kernel_stats(b200md_ctx *ctx, int index, char *name_out, int name_cap, double *total_ms,
                        long long *count);
int b200md_kernel_stats_reset(b200md_ctx *ctx);
/* raw CUDA stream the context launches on (cudaStream_t), for external event timing */
void *b200md_stream(b200md_ctx *ctx);
/* CUDA-event timing on that stream: record into slot 0..7, then elapsed ms between two slots
 * (synchronises on the later event) */
int b200md_event_record(b200md_ctx *ctx, int slot);
double b200md_event_elapsed_ms(b200md_ctx *ctx, int slot_a, int slot_b);
/* roofline denominators measured in place on this context's device: a DFMA-saturating kernel (TFLOP/s FP64) and
 * a 2 GiB device-to-device copy (GB/s, read + write bytes).  Either pointer may be NULL. */
int b200md_measure_peaks(b200md_ctx *ctx, double *fp64_tflops, double *hbm_gbs);
/* one plain copy of `bytes` between a host block and device staging memory, timed with CUDA events (ms): what the link
 * alone needs for a per-step position upload (to_device = 1) or force download (0) */
int b200md_copy_probe(b200md_ctx *ctx, void *host, size_t bytes, int to_device, double *ms);
/* page-locked host memory for callers that want DMA-speed x/f transfers */
void *b200md_host_alloc(size_t bytes);
void b200md_host_free(void *p);
/* page-lock memory the CALLER owns (LAMMPS' atom->x / atom->f blocks, Memory::create'd: pair_rebomos.cpp reads them
 * through the Pointers members): the piecewise x upload and ranged f download of the compute entry points only overlap
 * with the kernels when the host side is pinned.  Returns 0 on success; a failure is not fatal (transfers are then
 * staged by the driver).  Unregister before the block is freed or reallocated. */
int b200md_host_register(void *p, size_t bytes);
int b200md_host_unregister(void *p);

/* =============================================================================
 * GPU-resident MD system (the benchmark driver's run loop; one per GPU/rank)
 * ============================================================================ */
typedef struct {
  int style;             /* 0 = rebomos, 1 = aeam (potential set with *_init before) */
  int ntypes;
  const double *mass;    /* [ntypes+1], 1-based */
  b200md_box box;        /* global box; sublo/subhi/cutghost are computed by the library */
  int procgrid[3];       /* brick decomposition; product = number of ranks */
  int rank;
  double skin;           /* neighbor skin */
  double dt;             /* timestep (ps) */
  double ftm2v, mvv2e, boltz, nktv2p; /* unit constants (metal: Update::set_units) */
  int sort_every;        /* atom_modify sort frequency (0 = never) */
} b200md_system_desc;

/* upload this rank's owned atoms; builds ghosts + lists + initial forces (Verlet::setup) */
int b200md_system_create(b200md_ctx *ctx, const b200md_system_desc *d, int nlocal, const double *x,
                         const double *v, const int *type, const int *tag);
/* multi-GPU: 128-byte NCCL unique id from rank 0, then every rank joins */
int b200md_nccl_unique_id(void *id128);
int b200md_system_comm_init(b200md_ctx *ctx, const void *id128, int nranks, int rank);
/* single-process variant (tests on a one-GPU box, or several GPUs driven by host threads of one process):
 * ranks are contexts of THIS process, one host thread each; messages are device-to-device copies.
 * b200md_local_group_create returns a group id > 0 (negative B200MD_ERR_* on failure).             */
int b200md_local_group_create(int nranks);
int b200md_system_comm_init_local(b200md_ctx *ctx, int group, int nranks, int rank);
/* host-only helper (no device needed): the order in which CommBrick::exchange packs leaving atoms and fills the
 * holes, replayed on indices -- what makes per-rank atom order equal to LAMMPS' after migration.  leavers[]
 * ascending; order[nleave]; moves[2*nleave] receives (dst, src) pairs; returns 0 or B200MD_ERR_ARG. */
int b200md_exchange_plan(int n, const int *leavers, int nleave, int *order, int *moves, int *nmoves, int *nlocal_out);
/* advance n NVE steps entirely on the device; thermo quantities are evaluated on the last step
 * (and every thermo_every steps, retrievable with b200md_system_thermo) */
int b200md_system_run(b200md_ctx *ctx, int nsteps, int thermo_every);
/* `fix ID all nvt temp Tstart Tstop Tdamp` for the resident loop (USER-AEAM/sample.in:23): Nose-Hoover chain with LAMMPS'
 * defaults (tchain 3, tloop 1, no drag), restating FixNH [LAMMPS-core src/fix_nh.cpp] setup / initial_integrate /
 * final_integrate / nhc_temp_integrate.  The ramp runs over each b200md_system_run call like over a LAMMPS `run`.
 * Tdamp <= 0 switches back to plain NVE.  b200md_system_nh_energy = thermostat part of the conserved quantity. */
int b200md_system_set_nvt(b200md_ctx *ctx, double t_start, double t_stop, double t_period);
double b200md_system_nh_energy(b200md_ctx *ctx);
/* thermo of the most recent evaluation: out[0..11] = step, temp, press, pe, ke, vol, virial[6]
 * (global sums over ranks) */
int b200md_system_thermo(b200md_ctx *ctx, double *out);
int b200md_system_thermo_count(b200md_ctx *ctx);
int b200md_system_thermo_row(b200md_ctx *ctx, int i, double *out);
/* sizes: out[0]=nlocal, out[1]=nghost, out[2]=neighbor builds, out[3]=dangerous builds,
 * out[4]=atoms this rank sent away in CommBrick::exchange so far, out[5]=global atom count,
 * out[6]=inner-list refreshes between master rebuilds, out[7]=force evaluations with the halos overlapped (out[8]) */
int b200md_system_sizes(b200md_ctx *ctx, long long *out);
/* download owned+ghost state (any pointer may be NULL) */
int b200md_system_download(b200md_ctx *ctx, double *x, double *v, double *f, int *type, int *tag);

#ifdef __cplusplus
}
#endif
#endif
