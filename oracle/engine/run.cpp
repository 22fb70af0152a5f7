// minilmp engine (test infrastructure): Update (Verlet setup/run), fix nve,
// thermo, RanPark, the input-line interpreter, plugin loading and the C API
// that Python drives through ctypes.  Restated LAMMPS-core semantics:
// src/verlet.cpp, src/fix_nve.cpp, src/compute_temp.cpp,
// src/compute_pressure.cpp, src/thermo.cpp, src/create_atoms.cpp,
// src/replicate.cpp, src/velocity.cpp, src/set.cpp, src/random_park.cpp,
// src/PLUGIN/plugin.cpp -- SURVEY.md A.1, A.5, A.6.

#include "engine.h"

#include <algorithm>
#include <chrono>
#include <dlfcn.h>
#include <fstream>
#include <functional>
#include <thread>

using namespace LAMMPS_NS;

static double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ================================================================== RanPark
#define IA 16807
#define IM 2147483647
#define AM (1.0 / IM)
#define IQ 127773
#define IR 2836

double RanPark::uniform()
{
  int k = seed / IQ;
  seed = IA * (seed - k * IQ) - IR * k;
  if (seed < 0) seed += IM;
  double ans = AM * seed;
  return ans;
}
double RanPark::gaussian()
{
  double first, v1, v2, rsq, fac;
  if (!save) {
    do {
      v1 = 2.0 * uniform() - 1.0;
      v2 = 2.0 * uniform() - 1.0;
      rsq = v1 * v1 + v2 * v2;
    } while ((rsq >= 1.0) || (rsq == 0.0));
    fac = sqrt(-2.0 * log(rsq) / rsq);
    second = v1 * fac;
    first = v2 * fac;
    save = 1;
  } else {
    first = second;
    save = 0;
  }
  return first;
}
void RanPark::reset(int ibase, const double *coord)
{
  int i;
  char *str = (char *) &ibase;
  int n = sizeof(int);
  unsigned int hash = 0;
  for (i = 0; i < n; i++) {
    hash += str[i];
    hash += (hash << 10);
    hash ^= (hash >> 6);
  }
  str = (char *) coord;
  n = 3 * sizeof(double);
  for (i = 0; i < n; i++) {
    hash += str[i];
    hash += (hash << 10);
    hash ^= (hash >> 6);
  }
  hash += (hash << 3);
  hash ^= (hash >> 11);
  hash += (hash << 15);
  // keep 31 bits of unsigned int as new seed; do not allow seed = 0
  seed = hash & 0x7ffffff;
  if (!seed) seed = 1;
  // warm up the RNG
  for (i = 0; i < 5; i++) uniform();
  save = 0;
}

// ================================================================== Update (Verlet)
Update::Update(LAMMPS *l) : Pointers(l)
{
  ntimestep = firststep = laststep = 0;
  dt = 0.001;
  eflag_global = vflag_global = -1;
  nbuild = ndanger = 0;
  time_pair = time_neigh = time_comm = time_modify = time_loop = 0.0;
  set_units("lj");
}

void Update::set_units(const std::string &style)
{
  unit_style = style;
  if (style == "lj") {
    force->boltz = 1.0;
    force->mvv2e = 1.0;
    force->ftm2v = 1.0;
    force->nktv2p = 1.0;
    dt = 0.005;
    neighbor->skin = 0.3;
  } else if (style == "metal") {
    force->boltz = 8.617343e-5;
    force->hplanck = 4.135667403e-3;
    force->mvv2e = 1.0364269e-4;
    force->ftm2v = 1.0 / 1.0364269e-4;
    force->mv2d = 1.0 / 0.602214129;
    force->nktv2p = 1.6021765e6;
    force->qqr2e = 14.399645;
    dt = 0.001;
    neighbor->skin = 2.0;
  } else
    error->all(FLERR, "Illegal units command: minilmp supports lj and metal only");
}

void Update::ev_set(bigint step, int &eflag, int &vflag)
{
  // energy + virial only on thermo output steps (thermo_style with pe and press)
  eflag = vflag = 0;
  if (output->thermo_every > 0 && step == output->next_thermo) {
    eflag = Pair::ENERGY_GLOBAL;
    vflag = (vflag_global >= 0) ? vflag_global : Pair::VIRIAL_FDOTR;
    if (eflag_global >= 0) eflag = eflag_global;
  }
  if (step == laststep || step == firststep) {
    eflag = (eflag_global >= 0) ? eflag_global : Pair::ENERGY_GLOBAL;
    vflag = (vflag_global >= 0) ? vflag_global : Pair::VIRIAL_FDOTR;
  }
}

void Update::force_clear()
{
  int nall = atom->nlocal + atom->nghost;
  if (nall) memset(&atom->f[0][0], 0, sizeof(double) * 3 * (size_t) nall);
}

void Update::setup_run()
{
  if (!domain->box_exist) error->all(FLERR, "Run command before simulation box is defined");
  if (force->pair == nullptr) error->all(FLERR, "Run command before pair style is defined");
  for (int i = 1; i <= atom->ntypes; i++)
    if (!atom->mass_setflag[i]) error->all(FLERR, "Not all per-type masses are set");

  // LAMMPS::init(): force->init (pair->init), neighbor->init, comm->init
  force->init();
  neighbor->init();
  comm->init();

  // Verlet::setup
  int triclinic = domain->triclinic;
  if (triclinic) domain->x2lamda(atom->nlocal);
  domain->pbc();
  comm->setup();
  neighbor->setup_bins();
  comm->exchange();
  if (atom->sortfreq > 0) atom->sort();
  comm->borders();
  if (triclinic) domain->lamda2x(atom->nlocal + atom->nghost);
  neighbor->build(1);
  neighbor->ncalls = 0;
  neighbor->ndanger = 0;

  int eflag, vflag;
  firststep = ntimestep;
  if (output->thermo_every > 0) output->next_thermo = ntimestep;
  ev_set(ntimestep, eflag, vflag);
  force_clear();
  force->pair->compute(eflag, vflag);
  comm->reverse_comm();
  modify->setup();
  output->write_thermo(ntimestep);
}

void Update::run(int nsteps)
{
  int eflag, vflag;
  int triclinic = domain->triclinic;
  firststep = ntimestep;
  laststep = ntimestep + nsteps;
  setup_run();

  time_pair = time_neigh = time_comm = time_modify = 0.0;
  double t0 = now_s(), t1;

  for (int i = 0; i < nsteps; i++) {
    ++ntimestep;
    ev_set(ntimestep, eflag, vflag);

    t1 = now_s();
    modify->initial_integrate();
    time_modify += now_s() - t1;

    int nflag = neighbor->decide();
    if (nflag == 0) {
      t1 = now_s();
      comm->forward_comm();
      time_comm += now_s() - t1;
    } else {
      t1 = now_s();
      if (triclinic) domain->x2lamda(atom->nlocal);
      domain->pbc();
      comm->exchange();
      if (atom->sortfreq > 0 && ntimestep >= atom->nextsort) atom->sort();
      comm->borders();
      if (triclinic) domain->lamda2x(atom->nlocal + atom->nghost);
      time_comm += now_s() - t1;
      t1 = now_s();
      neighbor->build(1);
      time_neigh += now_s() - t1;
    }

    force_clear();
    t1 = now_s();
    force->pair->compute(eflag, vflag);
    time_pair += now_s() - t1;

    t1 = now_s();
    comm->reverse_comm();
    time_comm += now_s() - t1;

    t1 = now_s();
    modify->final_integrate();
    time_modify += now_s() - t1;

    if (output->thermo_every > 0 && ntimestep == output->next_thermo) output->write_thermo(ntimestep);
    else if (ntimestep == laststep) output->write_thermo(ntimestep);
  }
  universe->barrier();
  time_loop = now_s() - t0;
  nbuild = neighbor->ncalls;
  ndanger = neighbor->ndanger;
}

// ================================================================== fix nve
void Modify::nve_v()
{
  double dtf = 0.5 * update->dt * force->ftm2v;
  double **v = atom->v, **f = atom->f;
  double *mass = atom->mass;
  int *type = atom->type;
  int nlocal = atom->nlocal;
  for (int i = 0; i < nlocal; i++) {
    double dtfm = dtf / mass[type[i]];
    v[i][0] += dtfm * f[i][0];
    v[i][1] += dtfm * f[i][1];
    v[i][2] += dtfm * f[i][2];
  }
}

void Modify::nve_x()
{
  double dtv = update->dt;
  double **x = atom->x, **v = atom->v;
  int nlocal = atom->nlocal;
  for (int i = 0; i < nlocal; i++) {
    x[i][0] += dtv * v[i][0];
    x[i][1] += dtv * v[i][1];
    x[i][2] += dtv * v[i][2];
  }
}

void Modify::initial_integrate()
{
  if (nvt) {
    // FixNH::initial_integrate: thermostat half step, then the velocity-Verlet half kick and drift
    compute_temp_target();
    nhc_temp_integrate();
    nve_v();
    nve_x();
    return;
  }
  if (!nve) return;
  nve_v();
  nve_x();
}

void Modify::final_integrate()
{
  if (nvt) {
    nve_v();
    t_current = output->compute_temp();
    nhc_temp_integrate();
    return;
  }
  if (!nve) return;
  nve_v();
}

// ---- fix nvt (FixNH with tstat only)
void Modify::compute_temp_target()
{
  double delta = (double) (update->ntimestep - update->firststep);
  if (delta != 0.0) delta /= (double) (update->laststep - update->firststep);
  t_target = t_start + delta * (t_stop - t_start);
  ke_target = tdof * force->boltz * t_target;
}

void Modify::setup()
{
  if (!nvt) return;
  t_freq = 1.0 / t_period;
  tdof = 3.0 * (double) atom->natoms - 3.0;
  if (tdof < 0.0) tdof = 0.0;
  t_current = output->compute_temp();
  compute_temp_target();
  const double boltz = force->boltz;
  eta_mass[0] = tdof * boltz * t_target / (t_freq * t_freq);
  for (int ich = 1; ich < MTCHAIN; ich++) eta_mass[ich] = boltz * t_target / (t_freq * t_freq);
  for (int ich = 1; ich < MTCHAIN; ich++)
    eta_dotdot[ich] = (eta_mass[ich - 1] * eta_dot[ich - 1] * eta_dot[ich - 1] - boltz * t_target) / eta_mass[ich];
}

void Modify::nhc_temp_integrate()
{
  const double boltz = force->boltz;
  const double dt = update->dt;
  const double dthalf = 0.5 * dt, dt4 = 0.25 * dt, dt8 = 0.125 * dt;
  double expfac;
  double kecurrent = tdof * boltz * t_current;
  // eta_mass_flag = 1: masses follow the target temperature
  eta_mass[0] = tdof * boltz * t_target / (t_freq * t_freq);
  for (int ich = 1; ich < MTCHAIN; ich++) eta_mass[ich] = boltz * t_target / (t_freq * t_freq);
  if (eta_mass[0] > 0.0) eta_dotdot[0] = (kecurrent - ke_target) / eta_mass[0];
  else eta_dotdot[0] = 0.0;
  // nc_tchain = 1, tdrag_factor = 1
  for (int ich = MTCHAIN - 1; ich > 0; ich--) {
    expfac = exp(-dt8 * eta_dot[ich + 1]);
    eta_dot[ich] *= expfac;
    eta_dot[ich] += eta_dotdot[ich] * dt4;
    eta_dot[ich] *= expfac;
  }
  expfac = exp(-dt8 * eta_dot[1]);
  eta_dot[0] *= expfac;
  eta_dot[0] += eta_dotdot[0] * dt4;
  eta_dot[0] *= expfac;
  const double factor_eta = exp(-dthalf * eta_dot[0]);
  {    // nh_v_temp
    double **v = atom->v;
    int nlocal = atom->nlocal;
    for (int i = 0; i < nlocal; i++) {
      v[i][0] *= factor_eta;
      v[i][1] *= factor_eta;
      v[i][2] *= factor_eta;
    }
  }
  t_current *= factor_eta * factor_eta;
  kecurrent = tdof * boltz * t_current;
  if (eta_mass[0] > 0.0) eta_dotdot[0] = (kecurrent - ke_target) / eta_mass[0];
  else eta_dotdot[0] = 0.0;
  for (int ich = 0; ich < MTCHAIN; ich++) eta[ich] += dthalf * eta_dot[ich];
  eta_dot[0] *= expfac;
  eta_dot[0] += eta_dotdot[0] * dt4;
  eta_dot[0] *= expfac;
  for (int ich = 1; ich < MTCHAIN; ich++) {
    expfac = exp(-dt8 * eta_dot[ich + 1]);
    eta_dot[ich] *= expfac;
    eta_dotdot[ich] = (eta_mass[ich - 1] * eta_dot[ich - 1] * eta_dot[ich - 1] - boltz * t_target) / eta_mass[ich];
    eta_dot[ich] += eta_dotdot[ich] * dt4;
    eta_dot[ich] *= expfac;
  }
}

// FixNH::compute_scalar for a pure thermostat: sum_k ( kT_k eta_k + 0.5 m_k eta_dot_k^2 ), kT_0 = ke_target
double Modify::nh_energy() const
{
  if (!nvt) return 0.0;
  const double kt = force->boltz * t_target;
  double e = ke_target * eta[0] + 0.5 * eta_mass[0] * eta_dot[0] * eta_dot[0];
  for (int ich = 1; ich < MTCHAIN; ich++) e += kt * eta[ich] + 0.5 * eta_mass[ich] * eta_dot[ich] * eta_dot[ich];
  return e;
}

// ================================================================== thermo
double Output::compute_temp()
{
  double **v = atom->v;
  double *mass = atom->mass;
  int *type = atom->type;
  int nlocal = atom->nlocal;
  double t = 0.0;
  for (int i = 0; i < nlocal; i++)
    t += (v[i][0] * v[i][0] + v[i][1] * v[i][1] + v[i][2] * v[i][2]) * mass[type[i]];
  universe->allreduce_sum(comm->me, &t, 1);
  double dof = 3.0 * (double) atom->natoms - 3.0;
  if (dof < 0.0) dof = 0.0;
  double tfactor = (dof > 0.0) ? force->mvv2e / (dof * force->boltz) : 0.0;
  return t * tfactor;
}

void Output::compute_thermo(ThermoRow &row)
{
  row.step = update->ntimestep;
  double dof = 3.0 * (double) atom->natoms - 3.0;
  if (dof < 0.0) dof = 0.0;
  row.temp = compute_temp();
  double buf[7];
  buf[0] = force->pair->eng_vdwl + force->pair->eng_coul;
  for (int k = 0; k < 6; k++) buf[1 + k] = force->pair->virial[k];
  universe->allreduce_sum(comm->me, buf, 7);
  row.pe = buf[0];
  for (int k = 0; k < 6; k++) row.virial[k] = buf[1 + k];
  row.vol = domain->volume();
  double inv_volume = 1.0 / row.vol;
  row.press = (dof * force->boltz * row.temp + row.virial[0] + row.virial[1] + row.virial[2]) / 3.0 *
      inv_volume * force->nktv2p;
  row.ke = row.temp * 0.5 * dof * force->boltz;
  row.etotal = row.pe + row.ke;
}

void Output::write_thermo(bigint step)
{
  ThermoRow row;
  compute_thermo(row);
  if (comm->me == 0) rows.push_back(row);
  if (thermo_every > 0 && step == next_thermo) next_thermo = step + thermo_every;
}

// ================================================================== Input
static std::vector<std::string> split_words(const std::string &s)
{
  std::vector<std::string> w;
  size_t i = 0, n = s.size();
  while (i < n) {
    while (i < n && isspace((unsigned char) s[i])) i++;
    if (i >= n) break;
    size_t j = i;
    while (j < n && !isspace((unsigned char) s[j])) j++;
    w.emplace_back(s.substr(i, j - i));
    i = j;
  }
  return w;
}

// tiny arithmetic evaluator for $(...) immediate variables: + - * / ^ ( ) numbers
namespace {
struct Expr {
  const char *p;
  explicit Expr(const char *s) : p(s) {}
  void ws() { while (*p && isspace((unsigned char) *p)) p++; }
  double number()
  {
    ws();
    if (*p == '(') {
      p++;
      double v = sum();
      ws();
      if (*p != ')') throw LAMMPSException("ERROR: Invalid syntax in variable formula");
      p++;
      return v;
    }
    if (*p == '-') { p++; return -power(); }
    if (*p == '+') { p++; return power(); }
    char *end;
    double v = strtod(p, &end);
    if (end == p) throw LAMMPSException("ERROR: Invalid syntax in variable formula");
    p = end;
    return v;
  }
  double power()
  {
    double b = number();
    ws();
    if (*p == '^') { p++; return pow(b, power()); }
    return b;
  }
  double prod()
  {
    double v = power();
    for (;;) {
      ws();
      if (*p == '*') { p++; v *= power(); }
      else if (*p == '/') { p++; v /= power(); }
      else return v;
    }
  }
  double sum()
  {
    double v = prod();
    for (;;) {
      ws();
      if (*p == '+') { p++; v += prod(); }
      else if (*p == '-') { p++; v -= prod(); }
      else return v;
    }
  }
};
}    // namespace

std::string Input::substitute(const std::string &line)
{
  std::string out;
  size_t i = 0, n = line.size();
  while (i < n) {
    if (line[i] == '$' && i + 1 < n && line[i + 1] == '(') {
      size_t j = i + 2;
      int depth = 1;
      while (j < n && depth) {
        if (line[j] == '(') depth++;
        else if (line[j] == ')') depth--;
        j++;
      }
      if (depth) error->all(FLERR, "Invalid immediate variable");
      std::string e = line.substr(i + 2, j - i - 3);
      Expr ex(e.c_str());
      double v = ex.sum();
      char buf[64];
      snprintf(buf, sizeof(buf), "%.20g", v);    // LAMMPS formats $() results with %.20g
      out += buf;
      i = j;
    } else
      out += line[i++];
  }
  return out;
}

void Input::file(const std::string &path)
{
  // every rank reads the file itself (no broadcast needed with thread ranks)
  std::ifstream in(path);
  if (!in) error->all(FLERR, "Cannot open input script {}", path);
  std::string line, acc;
  while (std::getline(in, line)) {
    size_t last = line.find_last_not_of(" \t\r\n");
    if (last != std::string::npos && line[last] == '&') {
      acc += line.substr(0, last);
      acc += " ";
      continue;
    }
    acc += line;
    one(acc);
    acc.clear();
  }
  if (!acc.empty()) one(acc);
}

void Input::one(const std::string &raw)
{
  std::string line = raw;
  size_t hash = line.find('#');
  if (hash != std::string::npos) line = line.substr(0, hash);
  line = substitute(line);
  std::vector<std::string> w = split_words(line);
  if (w.empty()) return;
  std::string cmd = w[0];
  std::vector<std::string> a(w.begin() + 1, w.end());

  if (cmd == "units") {
    if (a.size() != 1) error->all(FLERR, "Illegal units command");
    if (domain->box_exist) error->all(FLERR, "Units command after simulation box is defined");
    update->set_units(a[0]);
  } else if (cmd == "atom_style") {
    if (a.size() != 1 || a[0] != "atomic") error->all(FLERR, "minilmp supports atom_style atomic only");
  } else if (cmd == "dimension") {
    if (a.size() != 1 || a[0] != "3") error->all(FLERR, "minilmp supports dimension 3 only");
  } else if (cmd == "boundary") {
    if (a.size() != 3 || a[0] != "p" || a[1] != "p" || a[2] != "p")
      error->all(FLERR, "minilmp supports boundary p p p only");
  } else if (cmd == "processors") {
    if (a.size() != 3) error->all(FLERR, "Illegal processors command");
    if (domain->box_exist) error->all(FLERR, "Processors command after simulation box is defined");
    for (int d = 0; d < 3; d++) comm->user_procgrid[d] = utils::inumeric(FLERR, a[d], false, lmp);
  } else if (cmd == "lattice") lattice(a);
  else if (cmd == "region") region(a);
  else if (cmd == "create_box") create_box(a);
  else if (cmd == "create_atoms") create_atoms(a);
  else if (cmd == "replicate") replicate(a);
  else if (cmd == "mass") mass(a);
  else if (cmd == "pair_style") pair_style(a);
  else if (cmd == "pair_coeff") pair_coeff(a);
  else if (cmd == "neighbor") neighbor_cmd(a);
  else if (cmd == "neigh_modify") neigh_modify(a);
  else if (cmd == "velocity") velocity(a);
  else if (cmd == "set") set_cmd(a);
  else if (cmd == "displace_atoms") displace_atoms(a);
  else if (cmd == "fix") fix(a);
  else if (cmd == "timestep") {
    if (a.size() != 1) error->all(FLERR, "Illegal timestep command");
    update->dt = utils::numeric(FLERR, a[0], false, lmp);
  } else if (cmd == "thermo") {
    if (a.size() != 1) error->all(FLERR, "Illegal thermo command");
    output->thermo_every = utils::inumeric(FLERR, a[0], false, lmp);
  } else if (cmd == "thermo_style" || cmd == "thermo_modify" || cmd == "log" || cmd == "echo") {
    // thermo rows always carry step temp press pe ke etotal vol
  } else if (cmd == "atom_modify") {
    if (a.size() == 3 && a[0] == "sort") {
      atom->sortfreq = utils::inumeric(FLERR, a[1], false, lmp);
      atom->userbinsize = utils::numeric(FLERR, a[2], false, lmp);
    } else
      error->all(FLERR, "minilmp supports 'atom_modify sort N binsize' only");
  } else if (cmd == "comm_modify") {
    if (a.size() == 2 && a[0] == "cutoff") comm->cutghostuser = utils::numeric(FLERR, a[1], false, lmp);
    else error->all(FLERR, "minilmp supports 'comm_modify cutoff X' only");
  } else if (cmd == "reset_timestep") {
    if (a.size() != 1) error->all(FLERR, "Illegal reset_timestep command");
    update->ntimestep = utils::inumeric(FLERR, a[0], false, lmp);
  } else if (cmd == "run") run(a);
  else if (cmd == "plugin") plugin(a);
  else
    error->all(FLERR, "Unknown command: {}", raw);
}

void Input::lattice(std::vector<std::string> &a)
{
  if (a.size() < 2) error->all(FLERR, "Illegal lattice command");
  Lattice *lat = domain->lattice;
  *lat = Lattice();
  std::string style = a[0];
  lat->scale = utils::numeric(FLERR, a[1], false, lmp);
  if (update->unit_style == "lj") error->all(FLERR, "minilmp lattice command needs units metal");
  auto add = [&](double x, double y, double z) { lat->basis.push_back({x, y, z}); };
  if (style == "fcc") {
    add(0.0, 0.0, 0.0); add(0.5, 0.5, 0.0); add(0.5, 0.0, 0.5); add(0.0, 0.5, 0.5);
  } else if (style == "bcc") {
    add(0.0, 0.0, 0.0); add(0.5, 0.5, 0.5);
  } else if (style == "sc") {
    add(0.0, 0.0, 0.0);
  } else if (style == "diamond") {
    add(0.0, 0.0, 0.0); add(0.0, 0.5, 0.5); add(0.5, 0.0, 0.5); add(0.5, 0.5, 0.0);
    add(0.25, 0.25, 0.25); add(0.25, 0.75, 0.75); add(0.75, 0.25, 0.75); add(0.75, 0.75, 0.25);
  } else if (style != "custom")
    error->all(FLERR, "Illegal lattice command: unsupported style {}", style);

  size_t i = 2;
  auto num = [&](size_t k) { return utils::numeric(FLERR, a.at(k), false, lmp); };
  try {
    while (i < a.size()) {
      if (a[i] == "origin") {
        for (int d = 0; d < 3; d++) lat->origin[d] = num(i + 1 + d);
        i += 4;
      } else if (a[i] == "a1" || a[i] == "a2" || a[i] == "a3") {
        if (style != "custom") error->all(FLERR, "Invalid option in lattice command for non-custom style");
        double *v = (a[i] == "a1") ? lat->a1 : (a[i] == "a2") ? lat->a2 : lat->a3;
        for (int d = 0; d < 3; d++) v[d] = num(i + 1 + d);
        i += 4;
      } else if (a[i] == "basis") {
        if (style != "custom") error->all(FLERR, "Invalid option in lattice command for non-custom style");
        double x = num(i + 1), y = num(i + 2), z = num(i + 3);
        if (x < 0.0 || x >= 1.0 || y < 0.0 || y >= 1.0 || z < 0.0 || z >= 1.0)
          error->all(FLERR, "Illegal lattice command");
        add(x, y, z);
        i += 4;
      } else
        error->all(FLERR, "Illegal lattice command: unsupported keyword {}", a[i]);
    }
  } catch (std::out_of_range &) {
    error->all(FLERR, "Illegal lattice command");
  }
  if (lat->basis.empty()) error->all(FLERR, "No basis atoms in lattice");
  lat->setup();
}

void Input::region(std::vector<std::string> &a)
{
  if (a.size() < 8) error->all(FLERR, "Illegal region command");
  Region r;
  r.id = a[0];
  r.style = a[1];
  Lattice *lat = domain->lattice;
  auto num = [&](size_t k) { return utils::numeric(FLERR, a.at(k), false, lmp); };
  r.xy = r.xz = r.yz = 0.0;
  if (r.style == "block" || r.style == "prism") {
    r.xlo = lat->xlattice * num(2);
    r.xhi = lat->xlattice * num(3);
    r.ylo = lat->ylattice * num(4);
    r.yhi = lat->ylattice * num(5);
    r.zlo = lat->zlattice * num(6);
    r.zhi = lat->zlattice * num(7);
    if (r.style == "prism") {
      if (a.size() < 11) error->all(FLERR, "Illegal region prism command");
      r.xy = lat->xlattice * num(8);
      r.xz = lat->xlattice * num(9);
      r.yz = lat->ylattice * num(10);
    }
  } else
    error->all(FLERR, "minilmp supports region block and prism only");
  for (auto &q : domain->regions)
    if (q.id == r.id) error->all(FLERR, "Reuse of region ID {}", r.id);
  domain->regions.push_back(r);
}

void Input::create_box(std::vector<std::string> &a)
{
  if (a.size() != 2) error->all(FLERR, "Illegal create_box command");
  if (domain->box_exist) error->all(FLERR, "Cannot create_box after simulation box is defined");
  int ntypes = utils::inumeric(FLERR, a[0], false, lmp);
  const Region *r = nullptr;
  for (auto &q : domain->regions)
    if (q.id == a[1]) r = &q;
  if (!r) error->all(FLERR, "Create_box region ID {} does not exist", a[1]);
  domain->triclinic = (r->style == "prism") ? 1 : 0;
  domain->boxlo[0] = r->xlo; domain->boxhi[0] = r->xhi;
  domain->boxlo[1] = r->ylo; domain->boxhi[1] = r->yhi;
  domain->boxlo[2] = r->zlo; domain->boxhi[2] = r->zhi;
  domain->xy = r->xy; domain->xz = r->xz; domain->yz = r->yz;
  if (domain->triclinic) {
    // LAMMPS rejects skew > 0.5 (+ tolerance) of the periodic length
    double xprd = r->xhi - r->xlo, yprd = r->yhi - r->ylo;
    if (fabs(r->xy / xprd) > 0.5 + 1.0e-6 || fabs(r->xz / xprd) > 0.5 + 1.0e-6 ||
        fabs(r->yz / yprd) > 0.5 + 1.0e-6)
      error->all(FLERR, "Triclinic box skew is too large");
  }
  domain->box_exist = 1;
  atom->allocate_type_arrays(ntypes);
  if (comm->user_procgrid[0] == 0)
    error->all(FLERR, "processor grid not set");
  comm->set_proc_grid();
  domain->set_global_box();
  domain->set_local_box();
}

void Input::create_atoms(std::vector<std::string> &a)
{
  if (!domain->box_exist) error->all(FLERR, "Create_atoms command before simulation box is defined");
  if (a.size() < 2) error->all(FLERR, "Illegal create_atoms command");
  int ntype = utils::inumeric(FLERR, a[0], false, lmp);
  const Region *reg = nullptr;
  size_t iarg;
  if (a[1] == "box") iarg = 2;
  else if (a[1] == "region") {
    if (a.size() < 3) error->all(FLERR, "Illegal create_atoms command");
    for (auto &q : domain->regions)
      if (q.id == a[2]) reg = &q;
    if (!reg) error->all(FLERR, "Create_atoms region ID {} does not exist", a[2]);
    if (reg->style != "block") error->all(FLERR, "minilmp create_atoms region needs a block region");
    iarg = 3;
  } else
    error->all(FLERR, "minilmp supports create_atoms box|region only");

  Lattice *lat = domain->lattice;
  int nbasis = (int) lat->basis.size();
  std::vector<int> basistype(nbasis, ntype);
  while (iarg < a.size()) {
    if (a[iarg] == "basis") {
      if (iarg + 3 > a.size()) error->all(FLERR, "Illegal create_atoms command");
      int ib = utils::inumeric(FLERR, a[iarg + 1], false, lmp);
      int it = utils::inumeric(FLERR, a[iarg + 2], false, lmp);
      if (ib <= 0 || ib > nbasis || it <= 0 || it > atom->ntypes)
        error->all(FLERR, "Invalid basis setting in create_atoms command");
      basistype[ib - 1] = it;
      iarg += 3;
    } else
      error->all(FLERR, "Illegal create_atoms command: unsupported keyword {}", a[iarg]);
  }
  if (ntype <= 0 || ntype > atom->ntypes) error->all(FLERR, "Invalid atom type in create_atoms command");

  // sub-domain bounds; shrink upper bound at the periodic edge (CreateAtoms::command)
  const double EPSILON = 1.0e-6;
  int triclinic = domain->triclinic;
  double sublo[3], subhi[3];
  for (int d = 0; d < 3; d++) {
    sublo[d] = triclinic ? domain->sublo_lamda[d] : domain->sublo[d];
    subhi[d] = triclinic ? domain->subhi_lamda[d] : domain->subhi[d];
    if (comm->myloc[d] == comm->procgrid[d] - 1) {
      if (triclinic) subhi[d] -= 2.0 * EPSILON;
      else subhi[d] -= 2.0 * EPSILON * domain->prd[d];
    }
  }

  // CreateAtoms::add_lattice: loop bounds from the bounding box of my sub-box in lattice space
  double bboxlo[3], bboxhi[3];
  if (triclinic == 0)
    for (int d = 0; d < 3; d++) { bboxlo[d] = domain->sublo[d]; bboxhi[d] = domain->subhi[d]; }
  else
    domain->bbox(domain->sublo_lamda, domain->subhi_lamda, bboxlo, bboxhi);
  if (reg) {
    bboxlo[0] = MAX(bboxlo[0], reg->xlo); bboxhi[0] = MIN(bboxhi[0], reg->xhi);
    bboxlo[1] = MAX(bboxlo[1], reg->ylo); bboxhi[1] = MIN(bboxhi[1], reg->yhi);
    bboxlo[2] = MAX(bboxlo[2], reg->zlo); bboxhi[2] = MIN(bboxhi[2], reg->zhi);
  }
  double xmin, ymin, zmin, xmax, ymax, zmax;
  xmin = ymin = zmin = 1.0e30;
  xmax = ymax = zmax = -1.0e30;
  for (int c = 0; c < 8; c++)
    lat->bbox(1, (c & 1) ? bboxhi[0] : bboxlo[0], (c & 2) ? bboxhi[1] : bboxlo[1],
              (c & 4) ? bboxhi[2] : bboxlo[2], xmin, ymin, zmin, xmax, ymax, zmax);
  int ilo = static_cast<int>(xmin) - 1, jlo = static_cast<int>(ymin) - 1, klo = static_cast<int>(zmin) - 1;
  int ihi = static_cast<int>(xmax) + 1, jhi = static_cast<int>(ymax) + 1, khi = static_cast<int>(zmax) + 1;
  if (xmin < 0.0) ilo--;
  if (ymin < 0.0) jlo--;
  if (zmin < 0.0) klo--;

  int nprev = atom->nlocal;
  for (int k = klo; k <= khi; k++)
    for (int j = jlo; j <= jhi; j++)
      for (int i = ilo; i <= ihi; i++)
        for (int m = 0; m < nbasis; m++) {
          double x[3], lamda[3], *coord;
          x[0] = i + lat->basis[m][0];
          x[1] = j + lat->basis[m][1];
          x[2] = k + lat->basis[m][2];
          lat->lattice2box(x[0], x[1], x[2]);
          if (reg) {
            if (x[0] < reg->xlo || x[0] > reg->xhi || x[1] < reg->ylo || x[1] > reg->yhi ||
                x[2] < reg->zlo || x[2] > reg->zhi)
              continue;
          }
          if (triclinic) { domain->x2lamda(x, lamda); coord = lamda; }
          else coord = x;
          if (coord[0] < sublo[0] || coord[0] >= subhi[0] || coord[1] < sublo[1] ||
              coord[1] >= subhi[1] || coord[2] < sublo[2] || coord[2] >= subhi[2])
            continue;
          atom->create_atom(basistype[m], x, 0);
        }

  // Atom::tag_extend: new atoms numbered rank by rank after the current max tag
  bigint nnew = atom->nlocal - nprev, total;
  double maxtag = 0.0;
  for (int i = 0; i < nprev; i++) maxtag = MAX(maxtag, (double) atom->tag[i]);
  universe->allreduce_max(comm->me, &maxtag, 1);
  bigint before = universe->scan_exclusive(comm->me, nnew, total);
  tagint itag = (tagint) maxtag + (tagint) before + 1;
  for (int i = nprev; i < atom->nlocal; i++) atom->tag[i] = itag++;
  atom->natoms += total;
}

// Replicate::command (non-bbox algorithm): images ordered ix, iy, iz (iz fastest), tags offset
void Input::replicate(std::vector<std::string> &a)
{
  if (a.size() != 3) error->all(FLERR, "Illegal replicate command");
  int nx = utils::inumeric(FLERR, a[0], false, lmp);
  int ny = utils::inumeric(FLERR, a[1], false, lmp);
  int nz = utils::inumeric(FLERR, a[2], false, lmp);
  if (nx <= 0 || ny <= 0 || nz <= 0) error->all(FLERR, "Illegal replicate command");

  // old system: my atoms as 8 doubles each (x, v, type, tag); every rank sees every rank's buffer in turn
  int nold = atom->nlocal;
  std::vector<double> mine(8 * (size_t) nold + 1);
  double maxtag_d = 0.0;
  for (int i = 0; i < nold; i++) {
    double *b = &mine[8 * (size_t) i];
    for (int d = 0; d < 3; d++) { b[d] = atom->x[i][d]; b[3 + d] = atom->v[i][d]; }
    b[6] = atom->type[i];
    b[7] = atom->tag[i];
    maxtag_d = MAX(maxtag_d, (double) atom->tag[i]);
  }
  universe->allreduce_max(comm->me, &maxtag_d, 1);
  const tagint maxtag = (tagint) maxtag_d;
  double old_xprd = domain->xprd, old_yprd = domain->yprd, old_zprd = domain->zprd;
  double old_xy = domain->xy, old_xz = domain->xz, old_yz = domain->yz;

  domain->boxhi[0] = domain->boxlo[0] + nx * old_xprd;
  domain->boxhi[1] = domain->boxlo[1] + ny * old_yprd;
  domain->boxhi[2] = domain->boxlo[2] + nz * old_zprd;
  if (domain->triclinic) {
    domain->xy = ny * old_xy;
    domain->xz = nz * old_xz;
    domain->yz = nz * old_yz;
  }
  domain->set_global_box();
  domain->set_local_box();
  double sublo[3], subhi[3];
  for (int d = 0; d < 3; d++) {
    sublo[d] = domain->triclinic ? domain->sublo_lamda[d] : domain->sublo[d];
    subhi[d] = domain->triclinic ? domain->subhi_lamda[d] : domain->subhi[d];
  }

  atom->nlocal = 0;
  atom->natoms = 0;
  std::vector<double> buf;
  for (int iproc = 0; iproc < comm->nprocs; iproc++) {
    int n8 = universe->sendrecv(comm->me, iproc, mine.data(), 8 * nold, buf);
    int nb = n8 / 8;
    for (int ix = 0; ix < nx; ix++)
      for (int iy = 0; iy < ny; iy++)
        for (int iz = 0; iz < nz; iz++)
          for (int m = 0; m < nb; m++) {
            const double *b = &buf[8 * (size_t) m];
            double x[3], lamda[3], *coord;
            if (domain->triclinic == 0) {
              x[0] = b[0] + ix * old_xprd;
              x[1] = b[1] + iy * old_yprd;
              x[2] = b[2] + iz * old_zprd;
            } else {
              x[0] = b[0] + ix * old_xprd + iy * old_xy + iz * old_xz;
              x[1] = b[1] + iy * old_yprd + iz * old_yz;
              x[2] = b[2] + iz * old_zprd;
            }
            domain->remap(x);
            if (domain->triclinic) { domain->x2lamda(x, lamda); coord = lamda; }
            else coord = x;
            if (coord[0] < sublo[0] || coord[0] >= subhi[0] || coord[1] < sublo[1] || coord[1] >= subhi[1] ||
                coord[2] < sublo[2] || coord[2] >= subhi[2])
              continue;
            tagint offset = iz * ny * nx * maxtag + iy * nx * maxtag + ix * maxtag;
            atom->create_atom((int) b[6], x, (tagint) b[7] + offset);
            int i = atom->nlocal - 1;
            for (int d = 0; d < 3; d++) atom->v[i][d] = b[3 + d];
          }
  }
  double tot = atom->nlocal;
  universe->allreduce_sum(comm->me, &tot, 1);
  atom->natoms = (bigint) tot;
}

void Input::mass(std::vector<std::string> &a)
{
  if (a.size() != 2) error->all(FLERR, "Illegal mass command");
  if (!domain->box_exist) error->all(FLERR, "Mass command before simulation box is defined");
  if (a[0] == "*") {
    for (int i = 1; i <= atom->ntypes; i++)
      atom->set_mass(FLERR, i, utils::numeric(FLERR, a[1], false, lmp));
  } else
    atom->set_mass(FLERR, utils::inumeric(FLERR, a[0], false, lmp), utils::numeric(FLERR, a[1], false, lmp));
}

void Input::pair_style(std::vector<std::string> &a)
{
  if (a.empty()) error->all(FLERR, "Illegal pair_style command");
  force->create_pair(a[0], 0);
  if (force->pair) {
    std::vector<char *> args;
    for (size_t i = 1; i < a.size(); i++) args.push_back(const_cast<char *>(a[i].c_str()));
    force->pair->settings((int) args.size(), args.data());
  }
}

void Input::pair_coeff(std::vector<std::string> &a)
{
  if (!domain->box_exist) error->all(FLERR, "Pair_coeff command before simulation box is defined");
  if (force->pair == nullptr) error->all(FLERR, "Pair_coeff command without a pair style");
  std::vector<char *> args;
  for (auto &s : a) args.push_back(const_cast<char *>(s.c_str()));
  force->pair->coeff((int) args.size(), args.data());
}

void Input::neighbor_cmd(std::vector<std::string> &a)
{
  if (a.size() != 2) error->all(FLERR, "Illegal neighbor command");
  neighbor->skin = utils::numeric(FLERR, a[0], false, lmp);
  if (neighbor->skin < 0.0) error->all(FLERR, "Illegal neighbor command");
  if (a[1] != "bin") error->all(FLERR, "minilmp supports neighbor style bin only");
}

void Input::neigh_modify(std::vector<std::string> &a)
{
  size_t i = 0;
  while (i < a.size()) {
    if (i + 1 >= a.size()) error->all(FLERR, "Illegal neigh_modify command");
    if (a[i] == "every") neighbor->every = utils::inumeric(FLERR, a[i + 1], false, lmp);
    else if (a[i] == "delay") neighbor->delay = utils::inumeric(FLERR, a[i + 1], false, lmp);
    else if (a[i] == "check") neighbor->dist_check = (a[i + 1] == "yes") ? 1 : 0;
    else if (a[i] == "one") neighbor->oneatom = utils::inumeric(FLERR, a[i + 1], false, lmp);
    else if (a[i] == "page") neighbor->pgsize = utils::inumeric(FLERR, a[i + 1], false, lmp);
    else error->all(FLERR, "Illegal neigh_modify command: unsupported keyword {}", a[i]);
    i += 2;
  }
}

// Velocity::create with defaults: dist uniform, mom yes, rot no, loop all, sum no
void Input::velocity(std::vector<std::string> &a)
{
  if (a.size() < 4 || a[0] != "all" || a[1] != "create")
    error->all(FLERR, "minilmp supports 'velocity all create T seed [dist uniform|gaussian]' only");
  double t_desired = utils::numeric(FLERR, a[2], false, lmp);
  int seed = utils::inumeric(FLERR, a[3], false, lmp);
  if (seed <= 0) error->all(FLERR, "Illegal velocity create command");
  int dist_flag = 0;
  for (size_t i = 4; i + 1 < a.size(); i += 2) {
    if (a[i] == "dist") dist_flag = (a[i + 1] == "gaussian") ? 1 : 0;
    else if (a[i] == "mom" && a[i + 1] == "yes") {}
    else if (a[i] == "rot" && a[i + 1] == "no") {}
    else if (a[i] == "loop" && a[i + 1] == "all") {}
    else error->all(FLERR, "Illegal velocity create option {}", a[i]);
  }
  for (int i = 1; i <= atom->ntypes; i++)
    if (!atom->mass_setflag[i]) error->all(FLERR, "Cannot use velocity create before setting masses");

  int nlocal = atom->nlocal;
  bigint natoms = atom->natoms;
  std::vector<int> map((size_t) natoms + 1, -1);
  for (int i = 0; i < nlocal; i++) {
    if (atom->tag[i] < 1 || atom->tag[i] > natoms) error->all(FLERR, "Atom IDs are not consecutive");
    map[atom->tag[i]] = i;
  }
  double **v = atom->v;
  RanPark random(seed);
  for (bigint i = 1; i <= natoms; i++) {
    double vx, vy, vz;
    if (dist_flag == 0) {
      vx = random.uniform() - 0.5;
      vy = random.uniform() - 0.5;
      vz = random.uniform() - 0.5;
    } else {
      vx = random.gaussian();
      vy = random.gaussian();
      vz = random.gaussian();
    }
    int m = map[i];
    if (m >= 0) {
      double factor = 1.0 / sqrt(atom->mass[atom->type[m]]);
      v[m][0] = vx * factor;
      v[m][1] = vy * factor;
      v[m][2] = vz * factor;
    }
  }

  // zero linear momentum
  double p[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < nlocal; i++) {
    double mm = atom->mass[atom->type[i]];
    p[0] += v[i][0] * mm;
    p[1] += v[i][1] * mm;
    p[2] += v[i][2] * mm;
    p[3] += mm;
  }
  universe->allreduce_sum(comm->me, p, 4);
  if (p[3] > 0.0)
    for (int i = 0; i < nlocal; i++) {
      v[i][0] -= p[0] / p[3];
      v[i][1] -= p[1] / p[3];
      v[i][2] -= p[2] / p[3];
    }

  // scale to the desired temperature
  double t = output->compute_temp();
  if (t == 0.0) error->all(FLERR, "Attempting to rescale a 0.0 temperature");
  double factor = sqrt(t_desired / t);
  for (int i = 0; i < nlocal; i++) {
    v[i][0] *= factor;
    v[i][1] *= factor;
    v[i][2] *= factor;
  }
}

// Set::command: "set region R|group all type/fraction T frac seed" and "set ... type T"
void Input::set_cmd(std::vector<std::string> &a)
{
  if (a.size() < 4) error->all(FLERR, "Illegal set command");
  if (!((a[0] == "group" && a[1] == "all") || a[0] == "region"))
    error->all(FLERR, "minilmp supports 'set group all' and 'set region ID' only");
  const Region *reg = nullptr;
  if (a[0] == "region") {
    for (auto &q : domain->regions)
      if (q.id == a[1]) reg = &q;
    if (!reg) error->all(FLERR, "Set region ID {} does not exist", a[1]);
  }
  auto selected = [&](int i) {
    if (!reg) return true;
    double *x = atom->x[i];
    return x[0] >= reg->xlo && x[0] <= reg->xhi && x[1] >= reg->ylo && x[1] <= reg->yhi &&
        x[2] >= reg->zlo && x[2] <= reg->zhi;
  };
  if (a[2] == "type/fraction") {
    if (a.size() != 6) error->all(FLERR, "Illegal set type/fraction command");
    int newtype = utils::inumeric(FLERR, a[3], false, lmp);
    double fraction = utils::numeric(FLERR, a[4], false, lmp);
    int seed = utils::inumeric(FLERR, a[5], false, lmp);
    if (newtype <= 0 || newtype > atom->ntypes) error->all(FLERR, "Invalid value in set command");
    if (fraction < 0.0 || fraction > 1.0 || seed <= 0) error->all(FLERR, "Invalid value in set command");
    RanPark ranatom(1);
    for (int i = 0; i < atom->nlocal; i++) {
      if (!selected(i)) continue;
      ranatom.reset(seed, atom->x[i]);
      if (ranatom.uniform() > fraction) continue;
      atom->type[i] = newtype;
    }
  } else if (a[2] == "type") {
    int newtype = utils::inumeric(FLERR, a[3], false, lmp);
    if (newtype <= 0 || newtype > atom->ntypes) error->all(FLERR, "Invalid value in set command");
    for (int i = 0; i < atom->nlocal; i++)
      if (selected(i)) atom->type[i] = newtype;
  } else
    error->all(FLERR, "minilmp supports set type and type/fraction only");
}

// DisplaceAtoms "all random dx dy dz seed"
void Input::displace_atoms(std::vector<std::string> &a)
{
  if (a.size() < 6 || a[0] != "all" || a[1] != "random")
    error->all(FLERR, "minilmp supports 'displace_atoms all random dx dy dz seed' only");
  double dx = utils::numeric(FLERR, a[2], false, lmp);
  double dy = utils::numeric(FLERR, a[3], false, lmp);
  double dz = utils::numeric(FLERR, a[4], false, lmp);
  int seed = utils::inumeric(FLERR, a[5], false, lmp);
  if (seed <= 0) error->all(FLERR, "Illegal displace_atoms random command");
  RanPark random(1);
  double **x = atom->x;
  for (int i = 0; i < atom->nlocal; i++) {
    random.reset(seed, x[i]);
    x[i][0] += dx * 2.0 * (random.uniform() - 0.5);
    x[i][1] += dy * 2.0 * (random.uniform() - 0.5);
    x[i][2] += dz * 2.0 * (random.uniform() - 0.5);
  }
}

void Input::fix(std::vector<std::string> &a)
{
  if (a.size() < 3) error->all(FLERR, "Illegal fix command");
  if (a[1] == "all" && a[2] == "nvt") {
    // fix ID all nvt temp Tstart Tstop Tdamp   (USER-AEAM/sample.in:23)
    if (a.size() != 7 || a[3] != "temp") error->all(FLERR, "minilmp supports 'fix ID all nvt temp Tstart Tstop Tdamp' only");
    modify->t_start = utils::numeric(FLERR, a[4], false, lmp);
    modify->t_stop = utils::numeric(FLERR, a[5], false, lmp);
    modify->t_period = utils::numeric(FLERR, a[6], false, lmp);
    if (modify->t_start <= 0.0 || modify->t_stop <= 0.0) error->all(FLERR, "Target temperature for fix nvt cannot be 0.0");
    if (modify->t_period <= 0.0) error->all(FLERR, "Fix nvt Tdamp must be > 0.0");
    modify->nvt = 1;
    modify->nve = 0;
    return;
  }
  if (a[1] != "all" || a[2] != "nve")
    error->all(FLERR, "minilmp supports 'fix ID all nve|nvt' only (got fix {} {})", a[1], a[2]);
  modify->nve = 1;
  modify->nvt = 0;
}

void Input::run(std::vector<std::string> &a)
{
  if (a.size() != 1) error->all(FLERR, "Illegal run command");
  int n = utils::inumeric(FLERR, a[0], false, lmp);
  if (n < 0) error->all(FLERR, "Invalid run command N value");
  update->run(n);
}

// plugin_register (src/PLUGIN/plugin.cpp): pair styles go into force->pair_map
static void plugin_register(lammpsplugin_t *plugin, void *ptr)
{
  LAMMPS *lmp = (LAMMPS *) ptr;
  if (plugin == nullptr) return;
  if ((plugin->version == nullptr) || (plugin->style == nullptr) || (plugin->name == nullptr) ||
      (plugin->info == nullptr) || (plugin->handle == nullptr))
    return;
  if (strcmp(plugin->version, LAMMPS_VERSION) != 0)
    lmp->error->warning(FLERR, std::string("Plugin was compiled for LAMMPS version ") + plugin->version);
  std::string pstyle = plugin->style;
  if (pstyle == "pair") {
    (*lmp->force->pair_map)[plugin->name] = (Force::PairCreator) plugin->creator.v1;
  } else
    lmp->error->all(FLERR, "Loading plugins for {} styles not yet implemented", pstyle);
}

void Input::plugin(std::vector<std::string> &a)
{
  if (a.size() != 2 || a[0] != "load") error->all(FLERR, "minilmp supports 'plugin load <file>' only");
  dlerror();
  void *dso = dlopen(a[1].c_str(), RTLD_NOW | RTLD_LOCAL);
  if (dso == nullptr) error->all(FLERR, "Open of file {} failed: {}", a[1], dlerror());
  dlerror();
  void *initfunc = dlsym(dso, "lammpsplugin_init");
  if (initfunc == nullptr) {
    dlclose(dso);
    error->all(FLERR, "Plugin symbol lookup failure in file {}: {}", a[1], dlerror());
  }
  (*(lammpsplugin_initfunc) (initfunc))((void *) lmp, dso, (void *) &plugin_register);
}

// ================================================================== instance
LAMMPS *LAMMPS_NS::create_instance(Universe *u, int rank, const int *procgrid)
{
  LAMMPS *lmp = new LAMMPS();
  lmp->universe = u;
  lmp->world = &u->ctx[rank];
  lmp->memory = new Memory(lmp);
  lmp->error = new Error(lmp);
  lmp->input = new Input(lmp);
  lmp->atom = new Atom(lmp);
  lmp->neighbor = new Neighbor(lmp);
  lmp->comm = new Comm(lmp);
  lmp->domain = new Domain(lmp);
  lmp->force = new Force(lmp);
  lmp->modify = new Modify(lmp);
  lmp->output = new Output(lmp);
  lmp->update = new Update(lmp);
  for (int d = 0; d < 3; d++) lmp->comm->user_procgrid[d] = procgrid[d];
  return lmp;
}

void LAMMPS_NS::destroy_instance(LAMMPS *lmp)
{
  delete lmp->update;
  delete lmp->output;
  delete lmp->modify;
  delete lmp->force;
  delete lmp->domain;
  delete lmp->comm;
  delete lmp->neighbor;
  delete lmp->atom;
  delete lmp->input;
  delete lmp->error;
  delete lmp->memory;
  delete lmp;
}

// ================================================================== C API (ctypes)
namespace {
struct Handle {
  Universe *universe;
  std::vector<LAMMPS *> ranks;
  std::string last_error;
};

int collective(Handle *h, const std::function<void(LAMMPS *)> &fn)
{
  int n = h->universe->nprocs;
  std::vector<std::string> errs(n);
  auto body = [&](int r) {
    try {
      fn(h->ranks[r]);
    } catch (std::exception &e) {
      errs[r] = e.what();
      if (errs[r].empty()) errs[r] = "unknown error";
      h->universe->abort_all();
    }
  };
  if (n == 1) {
    try {
      fn(h->ranks[0]);
    } catch (std::exception &e) {
      errs[0] = e.what();
    }
  } else {
    std::vector<std::thread> th;
    for (int r = 0; r < n; r++) th.emplace_back(body, r);
    for (auto &t : th) t.join();
  }
  h->last_error.clear();
  for (int r = 0; r < n; r++)
    if (!errs[r].empty() && errs[r].find("rank aborted") == std::string::npos) {
      h->last_error = errs[r];
      break;
    }
  if (h->last_error.empty())
    for (int r = 0; r < n; r++)
      if (!errs[r].empty()) h->last_error = errs[r];
  return h->last_error.empty() ? 0 : 1;
}
}    // namespace

extern "C" {

void *minilmp_open(int px, int py, int pz)
{
  if (px < 1 || py < 1 || pz < 1) return nullptr;
  Handle *h = new Handle();
  h->universe = new Universe(px * py * pz);
  int grid[3] = {px, py, pz};
  for (int r = 0; r < px * py * pz; r++) h->ranks.push_back(create_instance(h->universe, r, grid));
  return h;
}

void minilmp_close(void *ptr)
{
  Handle *h = (Handle *) ptr;
  if (!h) return;
  for (LAMMPS *l : h->ranks) destroy_instance(l);
  delete h->universe;
  delete h;
}

const char *minilmp_last_error(void *ptr) { return ((Handle *) ptr)->last_error.c_str(); }

int minilmp_nprocs(void *ptr) { return ((Handle *) ptr)->universe->nprocs; }

int minilmp_command(void *ptr, const char *line)
{
  Handle *h = (Handle *) ptr;
  std::string s(line);
  return collective(h, [&](LAMMPS *l) { l->input->one(s); });
}

int minilmp_file(void *ptr, const char *path)
{
  Handle *h = (Handle *) ptr;
  std::string s(path);
  return collective(h, [&](LAMMPS *l) { l->input->file(s); });
}

// init + Verlet::setup (ghosts, neighbor list, forces at current positions)
int minilmp_setup(void *ptr, int eflag, int vflag)
{
  Handle *h = (Handle *) ptr;
  return collective(h, [&](LAMMPS *l) {
    l->update->eflag_global = eflag;
    l->update->vflag_global = vflag;
    l->update->firststep = l->update->laststep = l->update->ntimestep;
    l->update->setup_run();
  });
}

// force_clear + pair->compute on the current positions/lists (+ optional reverse comm)
int minilmp_compute(void *ptr, int eflag, int vflag, int reverse)
{
  Handle *h = (Handle *) ptr;
  return collective(h, [&](LAMMPS *l) {
    l->update->force_clear();
    l->force->pair->compute(eflag, vflag);
    if (reverse) l->comm->reverse_comm();
  });
}

// re-send ghost positions from owners (after owned x was edited from outside)
int minilmp_forward_comm(void *ptr)
{
  Handle *h = (Handle *) ptr;
  return collective(h, [&](LAMMPS *l) { l->comm->forward_comm(); });
}

void minilmp_set_flags(void *ptr, int eflag, int vflag)
{
  Handle *h = (Handle *) ptr;
  for (LAMMPS *l : h->ranks) {
    l->update->eflag_global = eflag;
    l->update->vflag_global = vflag;
  }
}

long long minilmp_get_int(void *ptr, int rank, const char *name)
{
  Handle *h = (Handle *) ptr;
  LAMMPS *l = h->ranks.at(rank);
  std::string n(name);
  if (n == "nlocal") return l->atom->nlocal;
  if (n == "nghost") return l->atom->nghost;
  if (n == "nmax") return l->atom->nmax;
  if (n == "natoms") return l->atom->natoms;
  if (n == "ntypes") return l->atom->ntypes;
  if (n == "inum") return l->neighbor->list ? l->neighbor->list->inum : 0;
  if (n == "gnum") return l->neighbor->list ? l->neighbor->list->gnum : 0;
  if (n == "nbuild") return l->update->nbuild;
  if (n == "ndanger") return l->update->ndanger;
  if (n == "ntimestep") return l->update->ntimestep;
  if (n == "triclinic") return l->domain->triclinic;
  if (n == "nswap") return l->comm->nswap;
  if (n == "nstencil") return l->neighbor->nstencil;
  if (n == "mbinx") return l->neighbor->mbinx;
  if (n == "mbiny") return l->neighbor->mbiny;
  if (n == "mbinz") return l->neighbor->mbinz;
  if (n == "mbinxlo") return l->neighbor->mbinxlo;
  if (n == "mbinylo") return l->neighbor->mbinylo;
  if (n == "mbinzlo") return l->neighbor->mbinzlo;
  if (n == "nbinx") return l->neighbor->nbinx;
  if (n == "nbiny") return l->neighbor->nbiny;
  if (n == "nbinz") return l->neighbor->nbinz;
  if (n == "ghostneigh") return l->force->pair ? l->force->pair->ghostneigh : 0;
  if (n == "bytes_forward") return l->comm->bytes_forward;
  if (n == "bytes_reverse") return l->comm->bytes_reverse;
  return -999999;
}

double minilmp_get_double(void *ptr, int rank, const char *name)
{
  Handle *h = (Handle *) ptr;
  LAMMPS *l = h->ranks.at(rank);
  std::string n(name);
  if (n == "eng_vdwl") return l->force->pair->eng_vdwl;
  if (n == "dt") return l->update->dt;
  if (n == "skin") return l->neighbor->skin;
  if (n == "cutneighmax") return l->neighbor->cutneighmax;
  if (n == "xy") return l->domain->xy;
  if (n == "xz") return l->domain->xz;
  if (n == "yz") return l->domain->yz;
  if (n == "time_loop") return l->update->time_loop;
  if (n == "time_pair") return l->update->time_pair;
  if (n == "time_neigh") return l->update->time_neigh;
  if (n == "time_comm") return l->update->time_comm;
  if (n == "time_modify") return l->update->time_modify;
  if (n == "boltz") return l->force->boltz;
  if (n == "mvv2e") return l->force->mvv2e;
  if (n == "ftm2v") return l->force->ftm2v;
  if (n == "nktv2p") return l->force->nktv2p;
  if (n == "nh_energy") return l->modify->nh_energy();
  if (n == "nh_t_current") return l->modify->t_current;
  if (n == "binsizex") return l->neighbor->binsizex;
  if (n == "binsizey") return l->neighbor->binsizey;
  if (n == "binsizez") return l->neighbor->binsizez;
  if (n.size() == 7 && n.compare(0, 6, "virial") == 0) return l->force->pair->virial[n[6] - '0'];
  if (n.size() == 6 && n.compare(0, 5, "boxlo") == 0) return l->domain->boxlo[n[5] - '0'];
  if (n.size() == 6 && n.compare(0, 5, "boxhi") == 0) return l->domain->boxhi[n[5] - '0'];
  if (n.size() == 6 && n.compare(0, 5, "sublo") == 0)
    return l->domain->triclinic ? l->domain->sublo_lamda[n[5] - '0'] : l->domain->sublo[n[5] - '0'];
  if (n.size() == 6 && n.compare(0, 5, "subhi") == 0)
    return l->domain->triclinic ? l->domain->subhi_lamda[n[5] - '0'] : l->domain->subhi[n[5] - '0'];
  if (n.size() == 9 && n.compare(0, 8, "cutghost") == 0) return l->comm->cutghost[n[8] - '0'];
  if (n.size() == 5 && n.compare(0, 4, "mass") == 0) return l->atom->mass[n[4] - '0'];
  return NAN;
}

void *minilmp_get_ptr(void *ptr, int rank, const char *name)
{
  Handle *h = (Handle *) ptr;
  LAMMPS *l = h->ranks.at(rank);
  std::string n(name);
  if (n == "x") return l->atom->x ? (void *) &l->atom->x[0][0] : nullptr;
  if (n == "v") return l->atom->v ? (void *) &l->atom->v[0][0] : nullptr;
  if (n == "f") return l->atom->f ? (void *) &l->atom->f[0][0] : nullptr;
  if (n == "type") return l->atom->type;
  if (n == "tag") return l->atom->tag;
  if (n == "mass") return l->atom->mass;
  if (n == "numneigh") return l->neighbor->list ? l->neighbor->list->numneigh : nullptr;
  if (n == "ilist") return l->neighbor->list ? l->neighbor->list->ilist : nullptr;
  if (n == "pair") return l->force->pair;
  if (n == "eatom") return l->force->pair ? (void *) l->force->pair->eatom : nullptr;
  if (n == "vatom") return (l->force->pair && l->force->pair->vatom) ? (void *) &l->force->pair->vatom[0][0] : nullptr;
  if (n == "lammps") return l;
  if (n == "stencil") return l->neighbor->stencil;
  if (n == "cutneighsq") return l->neighbor->cutneighsq ? (void *) &l->neighbor->cutneighsq[0][0] : nullptr;
  if (n == "cutneighghostsq")
    return l->neighbor->cutneighghostsq ? (void *) &l->neighbor->cutneighghostsq[0][0] : nullptr;
  return nullptr;
}

// flatten the neighbor list of one rank to CSR: offsets[nrows+1], then values
long long minilmp_neigh_total(void *ptr, int rank)
{
  Handle *h = (Handle *) ptr;
  NeighList *list = h->ranks.at(rank)->neighbor->list;
  if (!list) return 0;
  long long tot = 0;
  int nrows = list->inum + list->gnum;
  for (int ii = 0; ii < nrows; ii++) tot += list->numneigh[list->ilist[ii]];
  return tot;
}
void minilmp_neigh_csr(void *ptr, int rank, long long *offsets, int *values)
{
  Handle *h = (Handle *) ptr;
  NeighList *list = h->ranks.at(rank)->neighbor->list;
  int nrows = list->inum + list->gnum;
  long long o = 0;
  for (int i = 0; i < nrows; i++) {
    offsets[i] = o;
    int n = list->numneigh[i];
    memcpy(values + o, list->firstneigh[i], sizeof(int) * (size_t) n);
    o += n;
  }
  offsets[nrows] = o;
}

// swap (halo) structure of one rank, for testing device halo plans
int minilmp_swap_info(void *ptr, int rank, int iswap, int *out /*[5+6]*/)
{
  Handle *h = (Handle *) ptr;
  Comm *c = h->ranks.at(rank)->comm;
  if (iswap < 0 || iswap >= c->nswap) return 1;
  out[0] = c->sendnum[iswap];
  out[1] = c->recvnum[iswap];
  out[2] = c->firstrecv[iswap];
  out[3] = c->sendproc[iswap];
  out[4] = c->recvproc[iswap];
  out[5] = c->pbc_flag[iswap];
  for (int k = 0; k < 6; k++) out[6 + k] = c->pbc[iswap][k];
  return 0;
}
void minilmp_swap_sendlist(void *ptr, int rank, int iswap, int *out)
{
  Handle *h = (Handle *) ptr;
  Comm *c = h->ranks.at(rank)->comm;
  memcpy(out, c->sendlist[iswap].data(), sizeof(int) * (size_t) c->sendnum[iswap]);
}

int minilmp_thermo_count(void *ptr) { return (int) ((Handle *) ptr)->ranks[0]->output->rows.size(); }
void minilmp_thermo_row(void *ptr, int i, double *out /*[13]*/)
{
  const ThermoRow &r = ((Handle *) ptr)->ranks[0]->output->rows.at(i);
  out[0] = (double) r.step;
  out[1] = r.temp;
  out[2] = r.press;
  out[3] = r.pe;
  out[4] = r.ke;
  out[5] = r.etotal;
  out[6] = r.vol;
  for (int k = 0; k < 6; k++) out[7 + k] = r.virial[k];
}
void minilmp_thermo_clear(void *ptr)
{
  for (LAMMPS *l : ((Handle *) ptr)->ranks) l->output->rows.clear();
}

}    // extern "C"
