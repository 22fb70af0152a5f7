// oracle/port/portplugin.cpp -- TEST INFRASTRUCTURE.  Wraps the plain restatement (port_kernels.c) as LAMMPS
// pair styles "rebomos" and "aeam" so the mini engine can host it exactly like the verbatim reference build.
#include "lammpsplugin.h"
#include "version.h"

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neigh_list.h"
#include "neighbor.h"
#include "pair.h"
#include "potential_file_reader.h"
#include "text_file_reader.h"
#include "tokenizer.h"
#include "utils.h"

#include "port_kernels.h"

#include <cstring>
#include <string>
#include <vector>

namespace LAMMPS_NS {

class PairPortBase : public Pair {
 public:
  explicit PairPortBase(LAMMPS *l) : Pair(l) {}
  void settings(int narg, char **) override
  {
    if (narg != 0) error->all(FLERR, "Illegal pair_style command");
  }
  void alloc_flags()
  {
    allocated = 1;
    int n = atom->ntypes;
    memory->create(setflag, n + 1, n + 1, "pair:setflag");
    memory->create(cutsq, n + 1, n + 1, "pair:cutsq");
    for (int i = 1; i <= n; i++)
      for (int j = i; j <= n; j++) setflag[i][j] = 0;
    delete[] map;
    map = new int[n + 1];
  }
  void begin_tally(port_tally &t)
  {
    t.eflag_global = eflag_global;
    t.vflag_global = vflag_global;    // explicit tallies only when LAMMPS did not choose fdotr
    t.eng_vdwl = 0.0;
    for (double &v : t.virial) v = 0.0;
  }
  void end_tally(const port_tally &t)
  {
    eng_vdwl += t.eng_vdwl;
    for (int k = 0; k < 6; k++) virial[k] += t.virial[k];
    if (vflag_fdotr) {
      int nall = atom->nlocal + atom->nghost;
      if (nall) port_virial_fdotr(nall, &atom->x[0][0], &atom->f[0][0], virial);
      vflag_fdotr = 0;
    }
  }
};

// ---------------------------------------------------------------- rebomos
class PairREBOMoSPort : public PairPortBase {
  port_rebomos_par par;
  std::vector<int> elem, rebo_num, store;
  std::vector<long> rebo_first;
  std::vector<double> nM, nS;

 public:
  explicit PairREBOMoSPort(LAMMPS *l) : PairPortBase(l)
  {
    single_enable = 0;
    restartinfo = 0;
    one_coeff = 1;
    ghostneigh = 1;
    manybody_flag = 1;
    centroidstressflag = CENTROID_NOTAVAIL;
    memset(&par, 0, sizeof(par));
  }
  ~PairREBOMoSPort() override
  {
    if (allocated) {
      memory->destroy(setflag);
      memory->destroy(cutsq);
      memory->destroy(cutghost);
    }
  }
  void coeff(int narg, char **arg) override
  {
    if (!allocated) {
      alloc_flags();
      memory->create(cutghost, atom->ntypes + 1, atom->ntypes + 1, "pair:cutghost");
    }
    if (narg != 3 + atom->ntypes) error->all(FLERR, "Incorrect args for pair coefficients");
    if (strcmp(arg[0], "*") != 0 || strcmp(arg[1], "*") != 0) error->all(FLERR, "Incorrect args for pair coefficients");
    for (int i = 3; i < narg; i++) {
      std::string e(arg[i]);
      if (e == "NULL") map[i - 2] = -1;
      else if (e == "Mo" || e == "M") map[i - 2] = 0;
      else if (e == "S") map[i - 2] = 1;
      else error->all(FLERR, "Incorrect args for pair coefficients");
    }
    double v[61];
    if (comm->me == 0) {
      PotentialFileReader reader(lmp, arg[2], "rebomos");
      try {
        for (double &d : v) d = reader.next_double();
      } catch (std::exception &e) {
        error->one(FLERR, "reading rebomos potential file {}\nREASON: {}\n", arg[2], e.what());
      }
    }
    MPI_Bcast(v, 61, MPI_DOUBLE, 0, world);
    port_rebomos_setup(&par, v);
    int n = atom->ntypes, count = 0;
    for (int i = 1; i <= n; i++)
      for (int j = i; j <= n; j++) {
        setflag[i][j] = (map[i] >= 0 && map[j] >= 0) ? 1 : 0;
        count += setflag[i][j];
      }
    if (count == 0) error->all(FLERR, "Incorrect args for pair coefficients");
  }
  void init_style() override
  {
    if (atom->tag_enable == 0) error->all(FLERR, "Pair style REBOMoS requires atom IDs");
    if (force->newton_pair == 0) error->all(FLERR, "Pair style REBOMoS requires newton pair on");
    neighbor->add_request(this, NeighConst::REQ_FULL | NeighConst::REQ_GHOST);
  }
  double init_one(int i, int j) override
  {
    if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
    cutghost[i][j] = cutghost[j][i] = par.rcmax[map[i]][map[j]];
    return 3.0 * par.rcmax[0][0];
  }
  void compute(int eflag, int vflag) override
  {
    ev_init(eflag, vflag);
    const int nlocal = atom->nlocal, nall = nlocal + atom->nghost;
    const int nrows = list->inum + list->gnum;
    elem.resize(nall);
    for (int i = 0; i < nall; i++) elem[i] = map[atom->type[i]];
    rebo_num.resize(nall);
    rebo_first.resize(nall);
    nM.resize(nall);
    nS.resize(nall);
    store.resize((size_t) nall * 24 + 1024);
    double *x = nall ? &atom->x[0][0] : nullptr, *f = nall ? &atom->f[0][0] : nullptr;
    if (port_rebo_neigh(&par, nrows, x, elem.data(), list->numneigh, list->firstneigh, rebo_num.data(),
                        rebo_first.data(), store.data(), (long) store.size(), nM.data(), nS.data()))
      error->one(FLERR, "Neighbor list overflow, boost neigh_modify one");
    port_tally t;
    begin_tally(t);
    t.vflag_global = vflag_either;    // the reference guards its v_tally calls with vflag_either
    port_frebo(&par, nlocal, x, elem.data(), atom->tag, rebo_num.data(), rebo_first.data(), store.data(), nM.data(),
               nS.data(), f, &t);
    port_flj(&par, nlocal, x, elem.data(), atom->tag, list->numneigh, list->firstneigh, f, &t);
    end_tally(t);
  }
};

// ---------------------------------------------------------------- aeam
class PairAEAMPort : public PairPortBase {
  int nel = 0, nnon = 0;
  std::vector<std::string> names;
  std::vector<int> nrho, nr;
  std::vector<double> drho, mass, dr, cut;
  std::vector<std::vector<double>> frho, rhor, z2r, sf, sr, sz;
  std::vector<double *> pf, pr, pz;
  std::vector<double> rho, fp;
  port_aeam_par par;

 public:
  explicit PairAEAMPort(LAMMPS *l) : PairPortBase(l)
  {
    restartinfo = 0;
    manybody_flag = 1;
    one_coeff = 1;
    comm_forward = comm_reverse = 1;
  }
  ~PairAEAMPort() override
  {
    if (allocated) {
      memory->destroy(setflag);
      memory->destroy(cutsq);
    }
  }
  void coeff(int narg, char **arg) override
  {
    if (!allocated) alloc_flags();
    if (narg != 3 + atom->ntypes) error->all(FLERR, "Incorrect args for pair coefficients");
    if (strcmp(arg[0], "*") != 0 || strcmp(arg[1], "*") != 0) error->all(FLERR, "Incorrect args for pair coefficients");
    // every rank reads the file itself (identical result to read-on-0 + broadcast)
    FILE *fp_ = utils::open_potential(arg[2], lmp, nullptr);
    if (!fp_) error->one(FLERR, "Cannot open AEAM potential file {}: {}", arg[2], utils::getsyserror());
    char line[1024];
    for (int i = 0; i < 12; i++)
      if (!fgets(line, 1024, fp_)) error->one(FLERR, "AEAM potential file is truncated");
    try {
      ValueTokenizer v(line);
      nel = v.next_int();
      nnon = v.next_int();
      v.next_int();
      names.clear();
      for (int i = 0; i < nel; i++) names.push_back(v.next_string());
      nrho.assign(nel, 0); drho.assign(nel, 0); mass.assign(nel, 0);
      nr.assign(nel * nel, 0); dr.assign(nel * nel, 0); cut.assign(nel * nel, 0);
      for (int i = 0; i < nel; i++) {
        if (!fgets(line, 1024, fp_)) throw FileReaderException("truncated");
        ValueTokenizer w(line);
        nrho[i] = w.next_int(); drho[i] = w.next_double(); mass[i] = w.next_double();
      }
      for (int k = 0; k < nel * nel; k++) {
        if (!fgets(line, 1024, fp_)) throw FileReaderException("truncated");
        ValueTokenizer w(line);
        nr[k] = w.next_int(); dr[k] = w.next_double(); cut[k] = w.next_double();
      }
      TextFileReader rd(fp_, "AEAM");
      frho.assign(nel, {}); rhor.assign(nel * nel, {}); z2r.assign(nel * nel, {});
      for (int i = 0; i < nel; i++) { frho[i].assign(nrho[i] + 1, 0.0); rd.next_dvector(&frho[i][1], nrho[i]); }
      for (int k = 0; k < nel * nel; k++) { rhor[k].assign(nr[k] + 1, 0.0); rd.next_dvector(&rhor[k][1], nr[k]); }
      for (int i = 0; i < nel; i++)
        for (int j = 0; j <= i; j++) { z2r[i * nel + j].assign(nr[i * nel + j] + 1, 0.0); rd.next_dvector(&z2r[i * nel + j][1], nr[i * nel + j]); }
    } catch (std::exception &e) {
      fclose(fp_);
      error->all(FLERR, "AEAM potential file parser error: {}", e.what());
    }
    fclose(fp_);
    for (int i = 3; i < narg; i++) {
      if (strcmp(arg[i], "NULL") == 0) { map[i - 2] = -1; continue; }
      int j;
      for (j = 0; j < nel; j++)
        if (names[j] == arg[i]) break;
      if (j < nel) map[i - 2] = j;
      else error->all(FLERR, "No matching element in AEAM potential file");
    }
    for (int i = 3; i < narg; i++)
      if (i - 3 >= nel || names[i - 3] != arg[i]) error->all(FLERR, "no matching atom order of input file and potential file");
    int n = atom->ntypes, count = 0;
    for (int i = 1; i <= n; i++)
      for (int j = i; j <= n; j++) {
        setflag[i][j] = (map[i] >= 0 && map[j] >= 0) ? 1 : 0;
        if (setflag[i][j] && i == j) atom->set_mass(FLERR, i, mass[map[i]]);
        count += setflag[i][j];
      }
    if (count == 0) error->all(FLERR, "Incorrect args for pair coefficients");
  }
  void init_style() override
  {
    if (force->newton_pair == 0) error->all(FLERR, "Pair style aeam requires newton pair on");
    sf.assign(nel, {}); sr.assign(nel * nel, {}); sz.assign(nel * nel, {});
    pf.assign(nel, nullptr); pr.assign(nel * nel, nullptr); pz.assign(nel * nel, nullptr);
    for (int i = 0; i < nel; i++) {
      sf[i].assign((size_t) (nrho[i] + 1) * 7, 0.0);
      port_aeam_interpolate(nrho[i], drho[i], frho[i].data(), sf[i].data());
      pf[i] = sf[i].data();
    }
    for (int k = 0; k < nel * nel; k++) {
      sr[k].assign((size_t) (nr[k] + 1) * 7, 0.0);
      port_aeam_interpolate(nr[k], dr[k], rhor[k].data(), sr[k].data());
      pr[k] = sr[k].data();
    }
    for (int i = 0; i < nel; i++)
      for (int j = 0; j <= i; j++) {
        const int k = i * nel + j;
        sz[k].assign((size_t) (nr[k] + 1) * 7, 0.0);
        port_aeam_interpolate(nr[k], dr[k], z2r[k].data(), sz[k].data());
        pz[k] = pz[j * nel + i] = sz[k].data();
      }
    par.nel = nel; par.nnonangular = nnon; par.nrho = nrho.data(); par.drho = drho.data();
    par.nr = nr.data(); par.dr = dr.data(); par.cut = cut.data();
    par.frho_spline = pf.data(); par.rhor_spline = pr.data(); par.z2r_spline = pz.data();
    neighbor->add_request(this, NeighConst::REQ_FULL);
  }
  double init_one(int i, int j) override
  {
    if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
    return cut[(i - 1) * nel + (j - 1)];
  }
  void compute(int eflag, int vflag) override
  {
    ev_init(eflag, vflag);
    const int nlocal = atom->nlocal, nall = nlocal + atom->nghost;
    rho.resize(atom->nmax);
    fp.resize(atom->nmax);
    double *x = nall ? &atom->x[0][0] : nullptr, *f = nall ? &atom->f[0][0] : nullptr;
    port_tally t;
    begin_tally(t);
    t.vflag_global = vflag_either ? vflag_global : 0;
    port_aeam_density(&par, nlocal, nall, x, atom->type, list->numneigh, list->firstneigh, rho.data(), fp.data(), &t);
    comm->reverse_comm(this);
    comm->forward_comm(this);
    port_aeam_force(&par, nlocal, x, atom->type, list->numneigh, list->firstneigh, rho.data(), fp.data(), f, &t);
    end_tally(t);
  }
  int pack_forward_comm(int n, int *l, double *buf, int, int *) override
  {
    for (int i = 0; i < n; i++) buf[i] = fp[l[i]];
    return n;
  }
  void unpack_forward_comm(int n, int first, double *buf) override
  {
    for (int i = 0; i < n; i++) fp[first + i] = buf[i];
  }
  int pack_reverse_comm(int n, int first, double *buf) override
  {
    for (int i = 0; i < n; i++) buf[i] = rho[first + i];
    return n;
  }
  void unpack_reverse_comm(int n, int *l, double *buf) override
  {
    for (int i = 0; i < n; i++) rho[l[i]] += buf[i];
  }
};

}    // namespace LAMMPS_NS

using namespace LAMMPS_NS;
static Pair *make_rebomos(LAMMPS *l) { return new PairREBOMoSPort(l); }
static Pair *make_aeam(LAMMPS *l) { return new PairAEAMPort(l); }

extern "C" void lammpsplugin_init(void *lmp, void *handle, void *regfunc)
{
  lammpsplugin_t plugin;
  plugin.version = LAMMPS_VERSION;
  plugin.style = "pair";
  plugin.info = "oracle/port restatement (test infrastructure)";
  plugin.author = "b200md oracle";
  plugin.handle = handle;
  plugin.name = "rebomos";
  plugin.creator.v1 = (lammpsplugin_factory1 *) &make_rebomos;
  ((lammpsplugin_regfunc) regfunc)(&plugin, lmp);
  plugin.name = "aeam";
  plugin.creator.v1 = (lammpsplugin_factory1 *) &make_aeam;
  ((lammpsplugin_regfunc) regfunc)(&plugin, lmp);
}
