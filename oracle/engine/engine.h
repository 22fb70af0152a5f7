/* -*- c++ -*- ---------------------------------------------------------------
   minilmp engine -- TEST INFRASTRUCTURE (oracle side).

   A small restatement of the LAMMPS-core services a Pair plugin lives in
   (stable_2Aug2023 semantics): Domain (orthogonal +
   triclinic), Lattice, create_box/create_atoms/replicate, Atom::sort,
   CommBrick (exchange/borders/forward/reverse, incl. Pair hooks),
   NBinStandard + NStencilFull[Ghost]Bin3d + NPairFullBin[Ghost], Verlet,
   fix nve, thermo (temp/press/pe/ke), velocity create, set type/fraction,
   plugin load, and an input-line interpreter for the subset of commands
   the two shipped inputs use.  MPI ranks are threads of one process.

   It hosts (a) the reference pair styles compiled verbatim (oracle/_ref),
   (b) the plain restatement (oracle/port) and (c) the B200 plugins, which is
   how parity is asserted on identical inputs.  Only tests/, smoke() and
   bench.py's CPU-baseline legs may use it; no product path links it.
---------------------------------------------------------------------------- */
#ifndef MINILMP_ENGINE_H
#define MINILMP_ENGINE_H

#include "lmpshim.h"

#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

struct lmpshim_rankctx {
  LAMMPS_NS::Universe *universe;
  int rank;
};

namespace LAMMPS_NS {

// ------------------------------------------------------------ thread-rank "MPI"
class Universe {
 public:
  int nprocs;
  std::vector<lmpshim_rankctx> ctx;
  explicit Universe(int n);
  void barrier();
  void abort_all();
  // blocking exchange: every rank calls it once per swap
  int sendrecv(int me, int src, const double *sbuf, int nsend, std::vector<double> &rbuf);
  void allreduce_sum(int me, double *v, int n);
  void allreduce_max(int me, double *v, int n);
  bigint scan_exclusive(int me, bigint v, bigint &total);
  void bcast(int me, void *buf, size_t nbytes, int root);

 private:
  std::mutex mtx;
  std::condition_variable cv;
  int count, generation;
  bool aborted;
  std::vector<const void *> slot_ptr;
  std::vector<size_t> slot_n;
  std::vector<std::vector<double>> red;
};

// ------------------------------------------------------------ update / integrate / modify / output
struct ThermoRow {
  bigint step;
  double temp, press, pe, ke, etotal, vol;
  double virial[6];    // pair virial (global sum), energy units
};

class Update : protected Pointers {
 public:
  bigint ntimestep, firststep, laststep;
  double dt;
  std::string unit_style;
  int eflag_global, vflag_global;    // last ev_set result
  bigint nbuild, ndanger;
  double time_pair, time_neigh, time_comm, time_modify, time_loop;

  explicit Update(LAMMPS *lmp);
  void set_units(const std::string &style);
  void setup_run();           // Verlet::setup
  void run(int nsteps);       // Verlet::run
  void force_clear();
  void ev_set(bigint step, int &eflag, int &vflag);
};

class Modify : protected Pointers {
 public:
  int nve;    // 1 if "fix nve" on group all is defined
  // "fix ID all nvt temp Tstart Tstop Tdamp": Nose-Hoover chain thermostat, LAMMPS defaults (tchain 3, tloop 1,
  // drag 0), restating FixNH::setup / initial_integrate / final_integrate / nhc_temp_integrate [LAMMPS-core
  // src/fix_nh.cpp]; needed to run USER-AEAM/sample.in:23 literally
  int nvt;
  double t_start, t_stop, t_period;
  static const int MTCHAIN = 3;
  double eta[MTCHAIN], eta_dot[MTCHAIN + 1], eta_dotdot[MTCHAIN], eta_mass[MTCHAIN];
  double t_current, t_target, ke_target, tdof, t_freq;
  explicit Modify(LAMMPS *lmp) : Pointers(lmp), nve(0), nvt(0), t_start(0), t_stop(0), t_period(0)
  {
    for (int k = 0; k < MTCHAIN; k++) eta[k] = eta_dot[k] = eta_dotdot[k] = eta_mass[k] = 0.0;
    eta_dot[MTCHAIN] = 0.0;
    t_current = t_target = ke_target = tdof = t_freq = 0.0;
  }
  void setup();    // after the setup force computation (Verlet::setup -> Modify::setup)
  void initial_integrate();
  void final_integrate();
  double nh_energy() const;    // thermostat contribution to the conserved quantity (FixNH::compute_scalar)

 private:
  void nve_v();
  void nve_x();
  void compute_temp_target();
  void nhc_temp_integrate();
};

class Output : protected Pointers {
 public:
  int thermo_every;
  std::vector<ThermoRow> rows;
  explicit Output(LAMMPS *lmp) : Pointers(lmp), thermo_every(0) {}
  bigint next_thermo;
  void compute_thermo(ThermoRow &row);    // collective
  void write_thermo(bigint step);         // collective
  double compute_temp();                  // collective
};

// ------------------------------------------------------------ input
class Input : protected Pointers {
 public:
  explicit Input(LAMMPS *lmp) : Pointers(lmp) {}
  void one(const std::string &line);    // collective over ranks
  void file(const std::string &path);

 private:
  std::string substitute(const std::string &line);
  void lattice(std::vector<std::string> &a);
  void region(std::vector<std::string> &a);
  void create_box(std::vector<std::string> &a);
  void create_atoms(std::vector<std::string> &a);
  void replicate(std::vector<std::string> &a);
  void mass(std::vector<std::string> &a);
  void pair_style(std::vector<std::string> &a);
  void pair_coeff(std::vector<std::string> &a);
  void neighbor_cmd(std::vector<std::string> &a);
  void neigh_modify(std::vector<std::string> &a);
  void velocity(std::vector<std::string> &a);
  void set_cmd(std::vector<std::string> &a);
  void fix(std::vector<std::string> &a);
  void run(std::vector<std::string> &a);
  void plugin(std::vector<std::string> &a);
  void displace_atoms(std::vector<std::string> &a);
};

// Park-Miller RNG as used by LAMMPS' velocity/set commands (RanPark)
class RanPark {
 public:
  explicit RanPark(int seed_init) : seed(seed_init), save(0), second(0.0) {}
  double uniform();
  double gaussian();
  void reset(int ibase, const double *coord);

 private:
  int seed, save;
  double second;
};

// one rank's full instance
LAMMPS *create_instance(Universe *u, int rank, const int *procgrid);
void destroy_instance(LAMMPS *lmp);

}    // namespace LAMMPS_NS

#endif
