/* ----------------------------------------------------------------------
   pair_style rebomos -- B200-native host class (see pair_rebomos.h).

   Mirrors the reference's interface: settings/coeff/init_style/init_one
   semantics and error texts follow USER-REBOMOS/pair_rebomos.cpp:144-274,
   the file is read in the order of pair_rebomos.cpp:884-948.  compute()
   replaces REBO_neigh + FREBO + bondorder + FLJ + virial_fdotr_compute
   (pair_rebomos.cpp:102-111) by one call into the CUDA library.
------------------------------------------------------------------------- */

#include "pair_rebomos.h"

#include "b200md_host.h"
#include "force.h"
#include "modify.h"
#include "memory.h"
#include "potential_file_reader.h"
#include "text_file_reader.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

using namespace LAMMPS_NS;

/* ---------------------------------------------------------------------- */

PairREBOMoS::PairREBOMoS(LAMMPS *lmp) : Pair(lmp)
{
  single_enable = 0;
  restartinfo = 0;
  one_coeff = 1;
  ghostneigh = 1;
  manybody_flag = 1;
  centroidstressflag = CENTROID_NOTAVAIL;
  // The kernels evaluate every pair from both ends (gather form) and give a ghost partner nothing, so sum(x . f) over
  // owned + ghost atoms taken by SOMEBODY ELSE (PairHybrid::compute calls virial_fdotr_compute() itself after clearing
  // VIRIAL_FDOTR for its sub-styles) is not this style's virial.  With this flag the caller asks for VIRIAL_PAIR
  // instead and receives the virial the device computed.
  no_virial_fdotr_compute = 1;

  ctx = nullptr;
  cut3rebo = 0.0;
  last_list_step = -1;
  uploaded_nlocal = uploaded_nghost = -1;
  memset(&params, 0, sizeof(params));
}

/* ----------------------------------------------------------------------
   the class can be destructed when incomplete
------------------------------------------------------------------------- */

PairREBOMoS::~PairREBOMoS()
{
  B200MDHost::write_stats(ctx, "rebomos", comm->me);
  if (ctx) b200md_destroy(ctx);
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
    memory->destroy(cutghost);
  }
}

/* ---------------------------------------------------------------------- */

void PairREBOMoS::compute(int eflag, int vflag)
{
  ev_init(eflag, vflag);
  pinned.refresh(atom);
  {
    const int ow = B200MDHost::forces_zero_on_entry(this, force, modify) ? 1 : 0;
    if (ow != f_overwrite) {
      B200MDHost::check(error, ctx, b200md_set_option(ctx, "f_overwrite", ow), "option");
      f_overwrite = ow;
    }
  }

  const int nlocal = atom->nlocal;
  const int nghost = atom->nghost;
  // a rank whose brick holds no owned and no ghost atoms (vacuum in a slab run, many ranks): nothing to do, and
  // atom->x / atom->f need not even be allocated
  if (nlocal + nghost == 0) {
    vflag_fdotr = 0;
    return;
  }

  // LAMMPS rebuilt its neighbor list on this step (or atom counts changed): refresh the device list
  if (neighbor->ago == 0 || uploaded_nlocal != nlocal || uploaded_nghost != nghost) {
    int rc = B200MDHost::sync_neighbor_list(ctx, atom, neighbor, comm, domain, list, 1);
    B200MDHost::check(error, ctx, rc, "neighbor list hand-over");
    uploaded_nlocal = nlocal;
    uploaded_nghost = nghost;
  }

  double eng = 0.0, vir[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const int want_virial = (vflag_fdotr || vflag_global) ? B200MD_VIRIAL_FDOTR : 0;
  // per-atom tallies (compute pe/atom, stress/atom): Pair::eatom / Pair::vatom are accumulated on the device with the
  // reference's distribution (ev_tally halves, v_tally3 thirds, v_tally2 halves) and added here
  double *ea = eflag_atom ? eatom : nullptr;
  double *va = (vflag_atom && vatom) ? &vatom[0][0] : nullptr;
  int rc = b200md_rebomos_compute_peratom(ctx, nlocal, nghost, nlocal + nghost ? &atom->x[0][0] : nullptr, atom->type,
                                          atom->tag, eflag_global ? B200MD_ENERGY_GLOBAL : 0, want_virial,
                                          nlocal + nghost ? &atom->f[0][0] : nullptr, &eng, vir, ea, va);
  B200MDHost::check(error, ctx, rc, "rebomos force computation");

  if (eflag_global) eng_vdwl += eng;
  if (want_virial)
    for (int k = 0; k < 6; k++) virial[k] += vir[k];
  // the device already summed x (x) f over owned and ghost atoms: nothing left for virial_fdotr_compute()
  vflag_fdotr = 0;
}

/* ----------------------------------------------------------------------
   allocate all arrays
------------------------------------------------------------------------- */

void PairREBOMoS::allocate()
{
  allocated = 1;
  int n = atom->ntypes;

  memory->create(setflag, n + 1, n + 1, "pair:setflag");
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) setflag[i][j] = 0;

  memory->create(cutsq, n + 1, n + 1, "pair:cutsq");
  memory->create(cutghost, n + 1, n + 1, "pair:cutghost");
  map = new int[n + 1];
}

/* ----------------------------------------------------------------------
   global settings: pair_style rebomos takes no arguments
------------------------------------------------------------------------- */

void PairREBOMoS::settings(int narg, char ** /* arg */)
{
  if (narg != 0) error->all(FLERR, "Illegal pair_style command");
}

/* ----------------------------------------------------------------------
   pair_coeff * * <file> <element per atom type | NULL>
------------------------------------------------------------------------- */

void PairREBOMoS::coeff(int narg, char **arg)
{
  if (!allocated) allocate();

  if (narg != 3 + atom->ntypes) error->all(FLERR, "Incorrect args for pair coefficients");
  if (strcmp(arg[0], "*") != 0 || strcmp(arg[1], "*") != 0)
    error->all(FLERR, "Incorrect args for pair coefficients");

  // map[i] = which element (0 = Mo, 1 = S) the Ith atom type is, -1 if NULL
  for (int i = 3; i < narg; i++) {
    const std::string el(arg[i]);
    if (el == "NULL") map[i - 2] = -1;
    else if (el == "Mo" || el == "M") map[i - 2] = 0;    // "M": backward compatibility
    else if (el == "S") map[i - 2] = 1;
    else error->all(FLERR, "Incorrect args for pair coefficients");
  }

  read_file(arg[2]);

  int n = atom->ntypes;
  int count = 0;
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) {
      setflag[i][j] = 0;
      if (map[i] >= 0 && map[j] >= 0) {
        setflag[i][j] = 1;
        count++;
      }
    }
  if (count == 0) error->all(FLERR, "Incorrect args for pair coefficients");
}

/* ----------------------------------------------------------------------
   init specific to this pair style
------------------------------------------------------------------------- */

void PairREBOMoS::init_style()
{
  if (atom->tag_enable == 0) error->all(FLERR, "Pair style REBOMoS requires atom IDs");
  if (force->newton_pair == 0) error->all(FLERR, "Pair style REBOMoS requires newton pair on");
  if (atom->ntypes > 8) error->all(FLERR, "Pair style rebomos (B200) supports at most 8 atom types");

  // a full neighbor list, including neighbors of ghosts (same request as the reference) -- but only when LAMMPS' own
  // list is the one handed over (B200MD_NEIGH=host).  By default the list is rebuilt on the device from LAMMPS'
  // positions, bins and cutoffs, so the host need not spend its cores on 496-entry rows: no request, like the GPU
  // package's styles when the neighbor build runs on the device; Neighbor still decides WHEN to rebuild
  if (!B200MDHost::device_neighbor_build()) neighbor->add_request(this, NeighConst::REQ_FULL | NeighConst::REQ_GHOST);

  if (!ctx) {
    int rc = b200md_create(B200MDHost::pick_device(comm->me), &ctx);
    if (rc != B200MD_OK) error->one(FLERR, "Cannot open the B200 device: {}", b200md_last_error(nullptr));
  }
  int rc = b200md_rebomos_init(ctx, &params, atom->ntypes, map);
  B200MDHost::check(error, ctx, rc, "parameter upload");
  uploaded_nlocal = uploaded_nghost = -1;
}

/* ----------------------------------------------------------------------
   init for one type pair i,j and corresponding j,i
------------------------------------------------------------------------- */

double PairREBOMoS::init_one(int i, int j)
{
  if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");

  // cut3rebo = 3 REBO distances (Mo-Mo, the largest), returned for every type pair
  cut3rebo = 3.0 * params.rcmax[0];

  // cutghost = REBO cutoff used for the neighbors of ghosts
  cutghost[i][j] = cutghost[j][i] = params.rcmax[map[i] * 2 + map[j]];
  return cut3rebo;
}

/* ----------------------------------------------------------------------
   read the REBO potential file: one "value name" line per parameter, 61 of them
------------------------------------------------------------------------- */

void PairREBOMoS::read_file(char *filename)
{
  // order in the file: pair tables as MM, MS, SS triples; then per-element polynomials; then LJ
  enum { NPAIR = 7, NPOLY = 7, NCOORD = 4 };
  double pairv[NPAIR][3];           // rcmin rcmax Q alpha A BIJc Beta
  double bpoly[2][NPOLY], bgpoly[2][NPOLY], acoord[2][NCOORD];
  double eps[2], sig[2];
  std::vector<double *> slots;
  for (int t = 0; t < NPAIR; t++)
    for (int k = 0; k < 3; k++) slots.push_back(&pairv[t][k]);
  for (int e = 0; e < 2; e++) {
    for (int o = 0; o < NPOLY; o++) slots.push_back(&bpoly[e][o]);
    for (int o = 0; o < NPOLY; o++) slots.push_back(&bgpoly[e][o]);
  }
  for (int e = 0; e < 2; e++)
    for (int o = 0; o < NCOORD; o++) slots.push_back(&acoord[e][o]);
  slots.push_back(&eps[0]);
  slots.push_back(&eps[1]);
  slots.push_back(&sig[0]);
  slots.push_back(&sig[1]);

  std::vector<double> values(slots.size(), 0.0);
  if (comm->me == 0) {
    PotentialFileReader reader(lmp, filename, "rebomos");
    try {
      for (auto &v : values) v = reader.next_double();
    } catch (TokenizerException &e) {
      error->one(FLERR, "reading rebomos potential file {}\nREASON: {}\n", filename, e.what());
    } catch (FileReaderException &fre) {
      error->one(FLERR, "reading rebomos potential file {}\nREASON: {}\n", filename, fre.what());
    }
  }
  MPI_Bcast(values.data(), (int) values.size(), MPI_DOUBLE, 0, world);
  for (size_t k = 0; k < slots.size(); k++) *slots[k] = values[k];

  // symmetric 2x2 tables, row-major [itype][jtype], 0 = Mo, 1 = S
  auto fill = [](double *dst, const double *mm_ms_ss) {
    dst[0] = mm_ms_ss[0];
    dst[1] = dst[2] = mm_ms_ss[1];
    dst[3] = mm_ms_ss[2];
  };
  fill(params.rcmin, pairv[0]);
  fill(params.rcmax, pairv[1]);
  fill(params.Q, pairv[2]);
  fill(params.alpha, pairv[3]);
  fill(params.A, pairv[4]);
  fill(params.BIJc, pairv[5]);
  fill(params.Beta, pairv[6]);
  for (int e = 0; e < 2; e++) {
    for (int o = 0; o < NPOLY; o++) {
      params.b[o][e] = bpoly[e][o];
      params.bg[o][e] = bgpoly[e][o];
    }
    for (int o = 0; o < NCOORD; o++) params.a[o][e] = acoord[e][o];
  }
  // LJ: arithmetic sigma, geometric epsilon; window [rcmin, 2.5 sigma]
  const double sigma3[3] = {sig[0], (sig[0] + sig[1]) / 2, sig[1]};
  const double eps3[3] = {eps[0], sqrt(eps[0] * eps[1]), eps[1]};
  const double ljmax3[3] = {2.5 * sigma3[0], 2.5 * sigma3[1], 2.5 * sigma3[2]};
  fill(params.sigma, sigma3);
  fill(params.epsilon, eps3);
  fill(params.rcLJmin, pairv[0]);
  fill(params.rcLJmax, ljmax3);
}

/* ----------------------------------------------------------------------
   memory usage: everything per-atom lives on the device
------------------------------------------------------------------------- */

double PairREBOMoS::memory_usage()
{
  return 0.0;
}
