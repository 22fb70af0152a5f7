/* ----------------------------------------------------------------------
   pair_style aeam -- B200-native host class (see pair_aeam.h).

   Interface and file format follow USER-AEAM/pair_aeam.cpp: settings :513-517,
   coeff :523-595 (element names must match the file, in file order), read_file
   :627-746 (12 raw header lines, nelements "nrho drho mass" lines, nelements^2
   "nr dr cut" lines, then frho / rhor / z2r blocks), init_one :615-621.
   compute() replaces the density, embedding and force passes (:110-479).
------------------------------------------------------------------------- */

#include "pair_aeam.h"

#include "b200md_host.h"
#include "force.h"
#include "modify.h"
#include "memory.h"
#include "text_file_reader.h"
#include "tokenizer.h"
#include "utils.h"

#include <cmath>
#include <cstring>

using namespace LAMMPS_NS;

static constexpr int MAXLINE = 1024;

/* ---------------------------------------------------------------------- */

PairAEAM::PairAEAM(LAMMPS *lmp) : Pair(lmp)
{
  restartinfo = 0;
  manybody_flag = 1;
  one_coeff = 1;
  // gather-form kernels: the global virial must come from the device, never from somebody else's sum(x . f) over
  // atom->f (see pair_rebomos.cpp)
  no_virial_fdotr_compute = 1;

  nmax = 0;
  rho = fp = nullptr;
  setfl = nullptr;
  ctx = nullptr;
  uploaded_nlocal = uploaded_nghost = -1;

  // per-atom quantity exchanged by this Pair: fp forward (live), rho reverse (kept for interface parity)
  comm_forward = 1;
  comm_reverse = 1;
}

PairAEAM::~PairAEAM()
{
  B200MDHost::write_stats(ctx, "aeam", comm->me);
  if (ctx) b200md_destroy(ctx);
  memory->destroy(rho);
  if (fp) b200md_host_free(fp);
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
  }
  delete setfl;
}

/* ---------------------------------------------------------------------- */

void PairAEAM::compute(int eflag, int vflag)
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

  if (atom->nmax > nmax) {
    // fp travels between host and device twice per step: page-locked (the class owns it, unlike atom->x / atom->f);
    // rho stays on the device (the library hands out fp with the minrho test applied, option "fp_gated")
    memory->destroy(rho);
    if (fp) b200md_host_free(fp);
    nmax = atom->nmax;
    memory->create(rho, nmax, "pair:rho");
    fp = static_cast<double *>(b200md_host_alloc((size_t) nmax * sizeof(double)));
    if (!fp) error->one(FLERR, "Cannot allocate page-locked memory for the fp halo");
    for (int i = 0; i < nmax; i++) rho[i] = fp[i] = 0.0;
  }

  const int nlocal = atom->nlocal;
  const int nghost = atom->nghost;
  const int nall = nlocal + nghost;
  // a rank without owned and ghost atoms still takes part in the forward communication (it is collective)
  if (nall == 0) {
    comm->forward_comm(this);
    vflag_fdotr = 0;
    return;
  }

  if (neighbor->ago == 0 || uploaded_nlocal != nlocal || uploaded_nghost != nghost) {
    int rc = B200MDHost::sync_neighbor_list(ctx, atom, neighbor, comm, domain, list, 0);
    B200MDHost::check(error, ctx, rc, "neighbor list hand-over");
    uploaded_nlocal = nlocal;
    uploaded_nghost = nghost;
  }

  // per-atom tallies (compute pe/atom, stress/atom) start in the density phase: the embedding energy is an atom's own
  const int want_atom = (eflag_atom || vflag_atom) ? 1 : 0;
  B200MDHost::check(error, ctx, b200md_set_option(ctx, "peratom", want_atom), "option");

  // phase 1 on the device: density + embedding of owned atoms
  // what a neighbor needs from atom j is (rho_j > minrho ? fp_j : 0): the library hands out exactly that (option
  // "fp_gated", set in init_style), one double per atom, and rho never leaves the device
  int rc = b200md_aeam_density_phase(ctx, nlocal, nghost, nall ? &atom->x[0][0] : nullptr, atom->type, nullptr, fp);
  B200MDHost::check(error, ctx, rc, "aeam density pass");
  comm->forward_comm(this);

  // phase 2 on the device: pair + embedding forces, angular 3-body forces, energy, virial
  double eng = 0.0, vir[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const int want_virial = (vflag_fdotr || vflag_global) ? B200MD_VIRIAL_FDOTR : 0;
  rc = b200md_aeam_force_phase_peratom(ctx, nullptr, fp, eflag_global ? B200MD_ENERGY_GLOBAL : 0, want_virial,
                                       nall ? &atom->f[0][0] : nullptr, &eng, vir, eflag_atom ? eatom : nullptr,
                                       (vflag_atom && vatom) ? &vatom[0][0] : nullptr);
  B200MDHost::check(error, ctx, rc, "aeam force pass");

  if (eflag_global) eng_vdwl += eng;
  if (want_virial)
    for (int k = 0; k < 6; k++) virial[k] += vir[k];
  vflag_fdotr = 0;
}

/* ---------------------------------------------------------------------- */

void PairAEAM::allocate()
{
  allocated = 1;
  int n = atom->ntypes;

  memory->create(setflag, n + 1, n + 1, "pair:setflag");
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) setflag[i][j] = 0;
  memory->create(cutsq, n + 1, n + 1, "pair:cutsq");

  delete[] map;
  map = new int[n + 1];
  for (int i = 1; i <= n; i++) map[i] = -1;
}

void PairAEAM::settings(int narg, char ** /*arg*/)
{
  if (narg != 0) error->all(FLERR, "Illegal pair_style command");
}

/* ----------------------------------------------------------------------
   pair_coeff * * <file> <element per atom type>
------------------------------------------------------------------------- */

void PairAEAM::coeff(int narg, char **arg)
{
  if (!allocated) allocate();

  if (narg != 3 + atom->ntypes) error->all(FLERR, "Incorrect args for pair coefficients");
  if (strcmp(arg[0], "*") != 0 || strcmp(arg[1], "*") != 0)
    error->all(FLERR, "Incorrect args for pair coefficients");

  delete setfl;    // a repeated pair_coeff simply re-reads the file
  setfl = new Setfl();
  read_file(arg[2]);

  // map atom types to elements of the file
  for (int i = 3; i < narg; i++) {
    if (strcmp(arg[i], "NULL") == 0) {
      map[i - 2] = -1;
      continue;
    }
    int j;
    for (j = 0; j < setfl->nelements; j++)
      if (setfl->elements[j] == arg[i]) break;
    if (j < setfl->nelements) map[i - 2] = j;
    else error->all(FLERR, "No matching element in AEAM potential file");
  }
  // the tables are indexed by atom type directly, so types must list the elements in file order
  for (int i = 3; i < narg; i++)
    if (i - 3 >= setfl->nelements || setfl->elements[i - 3] != arg[i])
      error->all(FLERR, "no matching atom order of input file and potential file");

  int n = atom->ntypes;
  int count = 0;
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) {
      setflag[i][j] = 0;
      if (map[i] >= 0 && map[j] >= 0) {
        setflag[i][j] = 1;
        if (i == j) atom->set_mass(FLERR, i, setfl->mass[map[i]]);
        count++;
      }
    }
  if (count == 0) error->all(FLERR, "Incorrect args for pair coefficients");
}

/* ---------------------------------------------------------------------- */

void PairAEAM::init_style()
{
  if (force->newton_pair == 0) error->all(FLERR, "Pair style aeam requires newton pair on");
  if (!setfl) error->all(FLERR, "All pair coeffs are not set");
  if (setfl->nelements > 4) error->all(FLERR, "Pair style aeam (B200) supports at most 4 elements");

  if (!ctx) {
    int rc = b200md_create(B200MDHost::pick_device(comm->me), &ctx);
    if (rc != B200MD_OK) error->one(FLERR, "Cannot open the B200 device: {}", b200md_last_error(nullptr));
    B200MDHost::check(error, ctx, b200md_set_option(ctx, "fp_gated", 1), "option");
  }

  // hand the raw tables over; the library builds the splines (array2spline) and keeps them on the device
  const int nel = setfl->nelements;
  std::vector<const double *> pf(nel), pr(nel * nel), pz(nel * nel);
  for (int i = 0; i < nel; i++) pf[i] = setfl->frho[i].data();
  for (int k = 0; k < nel * nel; k++) {
    pr[k] = setfl->rhor[k].data();
    pz[k] = setfl->z2r[k].empty() ? nullptr : setfl->z2r[k].data();
  }
  b200md_aeam_tables t;
  t.nelements = nel;
  t.nnonangular = setfl->nnonangular;
  t.nrho = setfl->nrho.data();
  t.drho = setfl->drho.data();
  t.nr = setfl->nr.data();
  t.dr = setfl->dr.data();
  t.cut = setfl->cut.data();
  t.frho = pf.data();
  t.rhor = pr.data();
  t.z2r = pz.data();
  int rc = b200md_aeam_init(ctx, &t);
  B200MDHost::check(error, ctx, rc, "table upload");
  uploaded_nlocal = uploaded_nghost = -1;

  // LAMMPS' own full list is requested only when it is the one handed over (B200MD_NEIGH=host); see pair_rebomos.cpp
  if (!B200MDHost::device_neighbor_build()) neighbor->add_request(this, NeighConst::REQ_FULL);
}

double PairAEAM::init_one(int i, int j)
{
  if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
  return setfl->cut[(i - 1) * setfl->nelements + (j - 1)];
}

/* ----------------------------------------------------------------------
   read the multi-element AEAM file
------------------------------------------------------------------------- */

void PairAEAM::read_file(char *filename)
{
  Setfl *file = setfl;
  const int me = comm->me;
  FILE *fptr = nullptr;
  char line[MAXLINE];

  if (me == 0) {
    fptr = utils::open_potential(filename, lmp, nullptr);
    if (fptr == nullptr)
      error->one(FLERR, "Cannot open AEAM potential file {}: {}", filename, utils::getsyserror());
  }

  // 11 free-form header lines, line 12 = "nelements nnonangular nangular name..."
  int n = 0;
  if (me == 0) {
    line[0] = '\0';
    for (int i = 0; i < 12; i++)
      if (fgets(line, MAXLINE, fptr) == nullptr) error->one(FLERR, "AEAM potential file {} is truncated", filename);
    n = strlen(line) + 1;
  }
  MPI_Bcast(&n, 1, MPI_INT, 0, world);
  MPI_Bcast(line, n, MPI_CHAR, 0, world);

  try {
    ValueTokenizer values(line);
    file->nelements = values.next_int();
    file->nnonangular = values.next_int();
    file->nangular = values.next_int();
    for (int i = 0; i < file->nelements; i++) file->elements.push_back(values.next_string());
  } catch (std::exception &e) {
    error->all(FLERR, "AEAM potential file parser error: {}", e.what());
  }
  const int nel = file->nelements;
  if (nel < 1) error->all(FLERR, "AEAM potential file parser error: no elements");

  file->nrho.assign(nel, 0);
  file->drho.assign(nel, 0.0);
  file->mass.assign(nel, 0.0);
  file->nr.assign(nel * nel, 0);
  file->dr.assign(nel * nel, 0.0);
  file->cut.assign(nel * nel, 0.0);

  if (me == 0) {
    try {
      for (int i = 0; i < nel; i++) {
        if (fgets(line, MAXLINE, fptr) == nullptr) throw FileReaderException("unexpected end of file");
        ValueTokenizer values(line);
        file->nrho[i] = values.next_int();
        file->drho[i] = values.next_double();
        file->mass[i] = values.next_double();
      }
      for (int k = 0; k < nel * nel; k++) {
        if (fgets(line, MAXLINE, fptr) == nullptr) throw FileReaderException("unexpected end of file");
        ValueTokenizer values(line);
        file->nr[k] = values.next_int();
        file->dr[k] = values.next_double();
        file->cut[k] = values.next_double();
      }
    } catch (std::exception &e) {
      error->one(FLERR, "AEAM potential file parser error: {}", e.what());
    }
  }
  MPI_Bcast(file->nrho.data(), nel, MPI_INT, 0, world);
  MPI_Bcast(file->drho.data(), nel, MPI_DOUBLE, 0, world);
  MPI_Bcast(file->mass.data(), nel, MPI_DOUBLE, 0, world);
  MPI_Bcast(file->nr.data(), nel * nel, MPI_INT, 0, world);
  MPI_Bcast(file->dr.data(), nel * nel, MPI_DOUBLE, 0, world);
  MPI_Bcast(file->cut.data(), nel * nel, MPI_DOUBLE, 0, world);

  // tabulated values: frho per element, rhor for all (i,j) row-major, z2r for j <= i
  file->frho.assign(nel, {});
  file->rhor.assign(nel * nel, {});
  file->z2r.assign(nel * nel, {});
  for (int i = 0; i < nel; i++) file->frho[i].assign(file->nrho[i], 0.0);
  for (int k = 0; k < nel * nel; k++) file->rhor[k].assign(file->nr[k], 0.0);
  for (int i = 0; i < nel; i++)
    for (int j = 0; j <= i; j++) file->z2r[i * nel + j].assign(file->nr[i * nel + j], 0.0);

  if (me == 0) {
    try {
      TextFileReader reader(fptr, "AEAM");
      for (int i = 0; i < nel; i++) reader.next_dvector(file->frho[i].data(), file->nrho[i]);
      for (int k = 0; k < nel * nel; k++) reader.next_dvector(file->rhor[k].data(), file->nr[k]);
      for (int i = 0; i < nel; i++)
        for (int j = 0; j <= i; j++) reader.next_dvector(file->z2r[i * nel + j].data(), file->nr[i * nel + j]);
    } catch (std::exception &e) {
      error->one(FLERR, "AEAM potential file parser error: {}", e.what());
    }
    fclose(fptr);
  }
  for (int i = 0; i < nel; i++) MPI_Bcast(file->frho[i].data(), file->nrho[i], MPI_DOUBLE, 0, world);
  for (int k = 0; k < nel * nel; k++) MPI_Bcast(file->rhor[k].data(), file->nr[k], MPI_DOUBLE, 0, world);
  for (int i = 0; i < nel; i++)
    for (int j = 0; j <= i; j++)
      MPI_Bcast(file->z2r[i * nel + j].data(), file->nr[i * nel + j], MPI_DOUBLE, 0, world);
}

/* ----------------------------------------------------------------------
   halo hooks (same signatures and packing as the reference, pair_aeam.cpp:946-990)
------------------------------------------------------------------------- */

int PairAEAM::pack_forward_comm(int n, int *list, double *buf, int /*pbc_flag*/, int * /*pbc*/)
{
  int m = 0;
  for (int i = 0; i < n; i++) buf[m++] = fp[list[i]];
  return m;
}

void PairAEAM::unpack_forward_comm(int n, int first, double *buf)
{
  int m = 0;
  const int last = first + n;
  for (int i = first; i < last; i++) fp[i] = buf[m++];
}

int PairAEAM::pack_reverse_comm(int n, int first, double *buf)
{
  int m = 0;
  const int last = first + n;
  for (int i = first; i < last; i++) buf[m++] = rho[i];
  return m;
}

void PairAEAM::unpack_reverse_comm(int n, int *list, double *buf)
{
  int m = 0;
  for (int i = 0; i < n; i++) rho[list[i]] += buf[m++];
}

double PairAEAM::memory_usage()
{
  return 2.0 * nmax * sizeof(double);
}
