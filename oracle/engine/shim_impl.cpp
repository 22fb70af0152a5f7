// minilmp engine (test infrastructure): implementation of the lmpshim API
// surface -- Error, tokenizers/file readers, utils, thread-rank MPI, the Pair
// base class tallies (restated LAMMPS-core src/pair.cpp),
// NeighList and Force.

#include "engine.h"

#include <algorithm>
#include <cerrno>
#include <cctype>

using namespace LAMMPS_NS;

// ------------------------------------------------------------------ Error
void Error::all(const std::string &file, int line, const std::string &str)
{
  const char *base = strrchr(file.c_str(), '/');
  throw LAMMPSException("ERROR: " + str + " (" + (base ? base + 1 : file.c_str()) + ":" +
                        std::to_string(line) + ")");
}
void Error::one(const std::string &file, int line, const std::string &str)
{
  const char *base = strrchr(file.c_str(), '/');
  throw LAMMPSException("ERROR on proc: " + str + " (" + (base ? base + 1 : file.c_str()) + ":" +
                        std::to_string(line) + ")");
}
void Error::warning(const std::string &, int, const std::string &str)
{
  fprintf(stderr, "WARNING: %s\n", str.c_str());
}

// ------------------------------------------------------------------ tokenizer
TokenizerException::TokenizerException(const std::string &msg, const std::string &token)
{
  if (token.empty()) message = msg;
  else message = msg + ": '" + token + "'";
}

ValueTokenizer::ValueTokenizer(const std::string &str, const std::string &sep) : pos(0)
{
  size_t i = 0, n = str.size();
  while (i < n) {
    while (i < n && sep.find(str[i]) != std::string::npos) i++;
    if (i >= n) break;
    size_t j = i;
    while (j < n && sep.find(str[j]) == std::string::npos) j++;
    tokens.emplace_back(str.substr(i, j - i));
    i = j;
  }
}
std::string ValueTokenizer::next_string()
{
  if (pos >= tokens.size()) throw TokenizerException("Not enough tokens", "");
  return tokens[pos++];
}
bool utils::is_integer(const std::string &s)
{
  if (s.empty()) return false;
  size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
  if (i == s.size()) return false;
  for (; i < s.size(); i++)
    if (!isdigit((unsigned char) s[i])) return false;
  return true;
}
bool utils::is_double(const std::string &s)
{
  if (s.empty()) return false;
  char *end = nullptr;
  errno = 0;
  strtod(s.c_str(), &end);
  if (end == s.c_str() || *end != '\0') return false;
  // reject things strtod accepts but LAMMPS does not (hex, inf, nan)
  for (char c : s)
    if (!(isdigit((unsigned char) c) || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E'))
      return false;
  return true;
}
int ValueTokenizer::next_int()
{
  std::string t = next_string();
  if (!utils::is_integer(t)) throw InvalidIntegerException(t);
  return atoi(t.c_str());
}
bigint ValueTokenizer::next_bigint()
{
  std::string t = next_string();
  if (!utils::is_integer(t)) throw InvalidIntegerException(t);
  return atoll(t.c_str());
}
tagint ValueTokenizer::next_tagint() { return (tagint) next_bigint(); }
double ValueTokenizer::next_double()
{
  std::string t = next_string();
  if (!utils::is_double(t)) throw InvalidFloatException(t);
  return atof(t.c_str());
}

// ------------------------------------------------------------------ TextFileReader
TextFileReader::TextFileReader(const std::string &filename, const std::string &ftype) :
    filetype(ftype), closefp(true), ignore_comments(true)
{
  fp = fopen(filename.c_str(), "r");
  if (fp == nullptr)
    throw FileReaderException("cannot open " + filetype + " file " + filename + ": " +
                              utils::getsyserror());
}
TextFileReader::TextFileReader(FILE *f, std::string ftype) :
    filetype(std::move(ftype)), closefp(false), fp(f), ignore_comments(true)
{
  if (fp == nullptr) throw FileReaderException("Invalid file descriptor");
}
TextFileReader::~TextFileReader()
{
  if (closefp) fclose(fp);
}
void TextFileReader::skip_line()
{
  char *ptr = fgets(line, MAXLINE, fp);
  if (ptr == nullptr) throw EOFException("Missing line in " + filetype + " file!");
}
static int count_words(const char *s)
{
  int n = 0;
  while (*s) {
    while (*s && isspace((unsigned char) *s)) s++;
    if (!*s) break;
    n++;
    while (*s && !isspace((unsigned char) *s)) s++;
  }
  return n;
}
char *TextFileReader::next_line(int nparams)
{
  int n = 0, nwords = 0;
  char *ptr = fgets(line, MAXLINE, fp);
  if (ptr == nullptr) return nullptr;
  if (ignore_comments && (ptr = strchr(line, '#'))) *ptr = '\0';
  nwords = count_words(line);
  if (nwords > 0) n = strlen(line);
  while (nwords == 0 || nwords < nparams) {
    ptr = fgets(&line[n], MAXLINE - n, fp);
    if (ptr == nullptr) {
      if (nwords > 0 && nwords < nparams)
        throw EOFException("Incorrect format in " + filetype + " file! " + std::to_string(nwords) +
                           "/" + std::to_string(nparams) + " parameters");
      return nullptr;
    }
    if (ignore_comments && (ptr = strchr(line, '#'))) *ptr = '\0';
    nwords += count_words(&line[n]);
    if (nwords > 0) n = strlen(line);
  }
  return line;
}
void TextFileReader::next_dvector(double *list, int n)
{
  int i = 0;
  while (i < n) {
    char *ptr = next_line();
    if (ptr == nullptr) {
      if (i == 0) throw EOFException("EOF reached");
      throw FileReaderException("Incorrect format in " + filetype + " file! " + std::to_string(i) +
                                "/" + std::to_string(n) + " values");
    }
    ValueTokenizer values(line);
    while (values.has_next() && i < n) list[i++] = values.next_double();
  }
}
ValueTokenizer TextFileReader::next_values(int nparams, const std::string &separators)
{
  char *ptr = next_line(nparams);
  if (ptr == nullptr) throw EOFException("Missing line in " + filetype + " file!");
  return ValueTokenizer(line, separators);
}

// ------------------------------------------------------------------ PotentialFileReader
PotentialFileReader::PotentialFileReader(LAMMPS *l, const std::string &fname,
                                         const std::string &potential_name,
                                         const std::string &name_suffix, const int auto_convert) :
    lmp(l), reader(nullptr), filename(fname), filetype(potential_name + name_suffix),
    unit_convert(auto_convert)
{
  FILE *fp = utils::open_potential(fname, lmp, nullptr);
  if (fp == nullptr)
    lmp->error->one(FLERR, "cannot open {} potential file {}: {}", potential_name, fname,
                    utils::getsyserror());
  fclose(fp);
  try {
    reader = new TextFileReader(fname, filetype);
  } catch (FileReaderException &e) {
    lmp->error->one(FLERR, e.what());
  }
}
PotentialFileReader::PotentialFileReader(LAMMPS *l, const std::string &fname,
                                         const std::string &potential_name, const int auto_convert) :
    PotentialFileReader(l, fname, potential_name, " potential", auto_convert)
{
}
PotentialFileReader::~PotentialFileReader() { delete reader; }
double PotentialFileReader::next_double()
{
  char *line = reader->next_line(1);
  if (line == nullptr) throw FileReaderException("unexpected end of " + filetype + " file " + filename);
  return ValueTokenizer(line).next_double();
}
int PotentialFileReader::next_int()
{
  char *line = reader->next_line(1);
  if (line == nullptr) throw FileReaderException("unexpected end of " + filetype + " file " + filename);
  return ValueTokenizer(line).next_int();
}
std::string PotentialFileReader::next_string()
{
  char *line = reader->next_line(1);
  if (line == nullptr) throw FileReaderException("unexpected end of " + filetype + " file " + filename);
  return ValueTokenizer(line).next_string();
}

// ------------------------------------------------------------------ utils
FILE *utils::open_potential(const std::string &name, LAMMPS *, int *)
{
  FILE *fp = fopen(name.c_str(), "r");
  if (fp) return fp;
  const char *dir = getenv("LAMMPS_POTENTIALS");
  if (dir) {
    const char *base = strrchr(name.c_str(), '/');
    std::string p = std::string(dir) + "/" + (base ? base + 1 : name.c_str());
    fp = fopen(p.c_str(), "r");
  }
  return fp;
}
std::string utils::getsyserror() { return std::string(strerror(errno)); }
char *utils::strdup(const std::string &text)
{
  char *tmp = new char[text.size() + 1];
  strcpy(tmp, text.c_str());
  return tmp;
}
int utils::get_supported_conversions(const int property)
{
  if (property == ENERGY) return METAL2REAL | REAL2METAL;
  return NOCONVERT;
}
double utils::numeric(const char *file, int line, const std::string &str, bool, LAMMPS *lmp)
{
  if (!is_double(str) && !is_integer(str))
    lmp->error->all(file, line, "Expected floating point parameter instead of '{}' in input script or data file", str);
  return atof(str.c_str());
}
int utils::inumeric(const char *file, int line, const std::string &str, bool, LAMMPS *lmp)
{
  if (!is_integer(str))
    lmp->error->all(file, line, "Expected integer parameter instead of '{}' in input script or data file", str);
  return atoi(str.c_str());
}

// ------------------------------------------------------------------ Universe (thread ranks)
Universe::Universe(int n) :
    nprocs(n), ctx(n), count(0), generation(0), aborted(false), slot_ptr(n, nullptr), slot_n(n, 0),
    red(n)
{
  for (int i = 0; i < n; i++) {
    ctx[i].universe = this;
    ctx[i].rank = i;
  }
}
void Universe::barrier()
{
  if (nprocs == 1) return;
  std::unique_lock<std::mutex> lk(mtx);
  if (aborted) throw LAMMPSException("rank aborted");
  int gen = generation;
  if (++count == nprocs) {
    count = 0;
    generation++;
    cv.notify_all();
  } else {
    cv.wait(lk, [&] { return gen != generation || aborted; });
    if (aborted) throw LAMMPSException("another rank aborted");
  }
}
void Universe::abort_all()
{
  std::unique_lock<std::mutex> lk(mtx);
  aborted = true;
  cv.notify_all();
}
int Universe::sendrecv(int me, int src, const double *sbuf, int nsend, std::vector<double> &rbuf)
{
  if (nprocs == 1) {
    if ((int) rbuf.size() < nsend) rbuf.resize(nsend);
    if (nsend) memcpy(rbuf.data(), sbuf, sizeof(double) * nsend);
    return nsend;
  }
  // always publish: another rank may be reading from me even when I read my own buffer
  slot_ptr[me] = sbuf;
  slot_n[me] = nsend;
  barrier();
  int nrecv = (int) slot_n[src];
  if ((int) rbuf.size() < nrecv) rbuf.resize(nrecv);
  if (nrecv) memcpy(rbuf.data(), slot_ptr[src], sizeof(double) * nrecv);
  barrier();
  return nrecv;
}
void Universe::allreduce_sum(int me, double *v, int n)
{
  if (nprocs == 1) return;
  red[me].assign(v, v + n);
  barrier();
  for (int k = 0; k < n; k++) {
    double s = 0.0;
    for (int r = 0; r < nprocs; r++) s += red[r][k];
    v[k] = s;
  }
  barrier();
}
void Universe::allreduce_max(int me, double *v, int n)
{
  if (nprocs == 1) return;
  red[me].assign(v, v + n);
  barrier();
  for (int k = 0; k < n; k++) {
    double s = red[0][k];
    for (int r = 1; r < nprocs; r++) s = std::max(s, red[r][k]);
    v[k] = s;
  }
  barrier();
}
bigint Universe::scan_exclusive(int me, bigint v, bigint &total)
{
  if (nprocs == 1) { total = v; return 0; }
  red[me].assign(1, (double) v);
  barrier();
  bigint before = 0;
  total = 0;
  for (int r = 0; r < nprocs; r++) {
    if (r < me) before += (bigint) red[r][0];
    total += (bigint) red[r][0];
  }
  barrier();
  return before;
}
void Universe::bcast(int me, void *buf, size_t nbytes, int root)
{
  if (nprocs == 1) return;
  if (me == root) { slot_ptr[root] = buf; slot_n[root] = nbytes; }
  barrier();
  if (me != root && nbytes) memcpy(buf, slot_ptr[root], nbytes);
  barrier();
}

extern "C" int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
  if (comm == nullptr) return 0;
  comm->universe->bcast(comm->rank, buf, (size_t) count * (size_t) type, root);
  return 0;
}

// ------------------------------------------------------------------ Pair base (LAMMPS-core src/pair.cpp)
int Pair::instance_total = 0;

Pair::Pair(LAMMPS *l) : Pointers(l)
{
  instance_total++;
  eng_vdwl = eng_coul = 0.0;
  for (double &v : virial) v = 0.0;
  comm_forward = comm_reverse = comm_reverse_off = 0;
  single_enable = 1;
  born_matrix_enable = 0;
  single_hessian_enable = 0;
  restartinfo = 1;
  respa_enable = 0;
  one_coeff = 0;
  no_virial_fdotr_compute = 0;
  writedata = 0;
  finitecutflag = 0;
  ghostneigh = 0;
  unit_convert_flag = utils::NOCONVERT;
  did_mix = false;
  nextra = 0;
  pvector = nullptr;
  single_extra = 0;
  svector = nullptr;
  setflag = nullptr;
  cutsq = nullptr;
  cutghost = nullptr;
  ewaldflag = pppmflag = msmflag = dispersionflag = tip4pflag = dipoleflag = spinflag = 0;
  reinitflag = 1;
  centroidstressflag = CENTROID_SAME;
  tail_flag = 0;
  etail = ptail = etail_ij = ptail_ij = 0.0;
  trim_flag = 1;
  ncoultablebits = 12;
  ndisptablebits = 12;
  allocated = 0;
  compute_flag = 1;
  manybody_flag = 0;
  mixed_flag = 0;
  suffix_flag = 0;
  maxeatom = maxvatom = maxcvatom = 0;
  eatom = nullptr;
  vatom = nullptr;
  cvatom = nullptr;
  list = listhalf = listfull = nullptr;
  map = nullptr;
  copymode = 0;
  beyond_contact = nondefault_history_transfer = 0;
  cutforce = 0.0;
  ev_unset();
}

Pair::~Pair()
{
  if (copymode) return;
  memory->destroy(eatom);
  memory->destroy(vatom);
  memory->destroy(cvatom);
  delete[] map;    // LAMMPS frees map in the base class too (pair.cpp)
}

void Pair::init_style() { neighbor->add_request(this); }
void Pair::init_list(int, NeighList *ptr) { list = ptr; }
double Pair::memory_usage()
{
  double bytes = (double) maxeatom * sizeof(double);
  bytes += (double) maxvatom * 6 * sizeof(double);
  return bytes;
}
double Pair::mix_energy(double eps1, double eps2, double, double) { return sqrt(eps1 * eps2); }
double Pair::mix_distance(double s1, double s2) { return 0.5 * (s1 + s2); }

void Pair::init()
{
  int i, j;
  if (!allocated) error->all(FLERR, "All pair coeffs are not set");
  for (i = 1; i <= atom->ntypes; i++)
    if (setflag[i][i] == 0 && !manybody_flag) error->all(FLERR, "All pair coeffs are not set");

  // requests are made anew by every init (Neighbor::init_pair drops the old ones in LAMMPS)
  delete neighbor->request;
  neighbor->request = nullptr;
  init_style();

  cutforce = 0.0;
  double cut;
  for (i = 1; i <= atom->ntypes; i++)
    for (j = i; j <= atom->ntypes; j++) {
      did_mix = false;
      cut = init_one(i, j);
      cutsq[i][j] = cutsq[j][i] = cut * cut;
      cutforce = MAX(cutforce, cut);
    }
}

void Pair::ev_setup(int eflag, int vflag, int alloc)
{
  int i, n;

  eflag_either = eflag;
  eflag_global = eflag & ENERGY_GLOBAL;
  eflag_atom = eflag & ENERGY_ATOM;

  vflag_global = vflag & (VIRIAL_PAIR | VIRIAL_FDOTR);
  vflag_atom = vflag & VIRIAL_ATOM;
  if (vflag & VIRIAL_CENTROID && centroidstressflag != CENTROID_AVAIL) vflag_atom = 1;
  cvflag_atom = 0;
  if (vflag & VIRIAL_CENTROID && centroidstressflag == CENTROID_AVAIL) cvflag_atom = 1;
  vflag_either = vflag_global || vflag_atom || cvflag_atom;

  evflag = eflag_either || vflag_either;

  if (eflag_atom && atom->nmax > maxeatom) {
    maxeatom = atom->nmax;
    if (alloc) {
      memory->destroy(eatom);
      memory->create(eatom, maxeatom, "pair:eatom");
    }
  }
  if (vflag_atom && atom->nmax > maxvatom) {
    maxvatom = atom->nmax;
    if (alloc) {
      memory->destroy(vatom);
      memory->create(vatom, maxvatom, 6, "pair:vatom");
    }
  }

  if (eflag_global) eng_vdwl = eng_coul = 0.0;
  if (vflag_global)
    for (i = 0; i < 6; i++) virial[i] = 0.0;
  if (eflag_atom && alloc) {
    n = atom->nlocal;
    if (force->newton) n += atom->nghost;
    for (i = 0; i < n; i++) eatom[i] = 0.0;
  }
  if (vflag_atom && alloc) {
    n = atom->nlocal;
    if (force->newton) n += atom->nghost;
    for (i = 0; i < n; i++)
      for (int k = 0; k < 6; k++) vatom[i][k] = 0.0;
  }

  // if vflag_global = VIRIAL_FDOTR and pair::compute() calls virial_fdotr_compute()
  // compute global virial via (F dot r) instead of via pairwise summation
  if (vflag_global == VIRIAL_FDOTR && no_virial_fdotr_compute == 0) {
    vflag_fdotr = 1;
    vflag_global = 0;
    if (vflag_atom == 0 && cvflag_atom == 0) vflag_either = 0;
    if (vflag_either == 0 && eflag_either == 0) evflag = 0;
  } else
    vflag_fdotr = 0;
}

void Pair::ev_tally(int i, int j, int nlocal, int newton_pair, double evdwl, double ecoul,
                    double fpair, double delx, double dely, double delz)
{
  double evdwlhalf, ecoulhalf, epairhalf, v[6];

  if (eflag_either) {
    if (eflag_global) {
      if (newton_pair) {
        eng_vdwl += evdwl;
        eng_coul += ecoul;
      } else {
        evdwlhalf = 0.5 * evdwl;
        ecoulhalf = 0.5 * ecoul;
        if (i < nlocal) { eng_vdwl += evdwlhalf; eng_coul += ecoulhalf; }
        if (j < nlocal) { eng_vdwl += evdwlhalf; eng_coul += ecoulhalf; }
      }
    }
    if (eflag_atom) {
      epairhalf = 0.5 * (evdwl + ecoul);
      if (newton_pair || i < nlocal) eatom[i] += epairhalf;
      if (newton_pair || j < nlocal) eatom[j] += epairhalf;
    }
  }

  if (vflag_either) {
    v[0] = delx * delx * fpair;
    v[1] = dely * dely * fpair;
    v[2] = delz * delz * fpair;
    v[3] = delx * dely * fpair;
    v[4] = delx * delz * fpair;
    v[5] = dely * delz * fpair;

    if (vflag_global) {
      if (newton_pair) {
        for (int k = 0; k < 6; k++) virial[k] += v[k];
      } else {
        if (i < nlocal) for (int k = 0; k < 6; k++) virial[k] += 0.5 * v[k];
        if (j < nlocal) for (int k = 0; k < 6; k++) virial[k] += 0.5 * v[k];
      }
    }
    if (vflag_atom) {
      if (newton_pair || i < nlocal) for (int k = 0; k < 6; k++) vatom[i][k] += 0.5 * v[k];
      if (newton_pair || j < nlocal) for (int k = 0; k < 6; k++) vatom[j][k] += 0.5 * v[k];
    }
  }
}

void Pair::ev_tally3(int i, int j, int k, double evdwl, double ecoul, double *fj, double *fk,
                     double *drji, double *drki)
{
  double epairthird, v[6];

  if (eflag_either) {
    if (eflag_global) {
      eng_vdwl += evdwl;
      eng_coul += ecoul;
    }
    if (eflag_atom) {
      epairthird = MathConst::THIRD * (evdwl + ecoul);
      eatom[i] += epairthird;
      eatom[j] += epairthird;
      eatom[k] += epairthird;
    }
  }

  if (vflag_either) {
    v[0] = drji[0] * fj[0] + drki[0] * fk[0];
    v[1] = drji[1] * fj[1] + drki[1] * fk[1];
    v[2] = drji[2] * fj[2] + drki[2] * fk[2];
    v[3] = drji[0] * fj[1] + drki[0] * fk[1];
    v[4] = drji[0] * fj[2] + drki[0] * fk[2];
    v[5] = drji[1] * fj[2] + drki[1] * fk[2];

    if (vflag_global)
      for (int m = 0; m < 6; m++) virial[m] += v[m];
    if (vflag_atom)
      for (int m = 0; m < 6; m++) {
        vatom[i][m] += MathConst::THIRD * v[m];
        vatom[j][m] += MathConst::THIRD * v[m];
        vatom[k][m] += MathConst::THIRD * v[m];
      }
  }
}

void Pair::v_tally2(int i, int j, double fpair, double *drij)
{
  double v[6];
  v[0] = drij[0] * drij[0] * fpair;
  v[1] = drij[1] * drij[1] * fpair;
  v[2] = drij[2] * drij[2] * fpair;
  v[3] = drij[0] * drij[1] * fpair;
  v[4] = drij[0] * drij[2] * fpair;
  v[5] = drij[1] * drij[2] * fpair;
  if (vflag_global)
    for (int m = 0; m < 6; m++) virial[m] += v[m];
  if (vflag_atom)
    for (int m = 0; m < 6; m++) {
      vatom[i][m] += 0.5 * v[m];
      vatom[j][m] += 0.5 * v[m];
    }
}

void Pair::v_tally3(int i, int j, int k, double *fi, double *fj, double *drik, double *drjk)
{
  double v[6];
  v[0] = (drik[0] * fi[0] + drjk[0] * fj[0]);
  v[1] = (drik[1] * fi[1] + drjk[1] * fj[1]);
  v[2] = (drik[2] * fi[2] + drjk[2] * fj[2]);
  v[3] = (drik[0] * fi[1] + drjk[0] * fj[1]);
  v[4] = (drik[0] * fi[2] + drjk[0] * fj[2]);
  v[5] = (drik[1] * fi[2] + drjk[1] * fj[2]);
  if (vflag_global)
    for (int m = 0; m < 6; m++) virial[m] += v[m];
  if (vflag_atom)
    for (int m = 0; m < 6; m++) {
      vatom[i][m] += MathConst::THIRD * v[m];
      vatom[j][m] += MathConst::THIRD * v[m];
      vatom[k][m] += MathConst::THIRD * v[m];
    }
}

void Pair::v_tally4(int i, int j, int k, int m, double *fi, double *fj, double *fk, double *drim,
                    double *drjm, double *drkm)
{
  double v[6];
  v[0] = (drim[0] * fi[0] + drjm[0] * fj[0] + drkm[0] * fk[0]);
  v[1] = (drim[1] * fi[1] + drjm[1] * fj[1] + drkm[1] * fk[1]);
  v[2] = (drim[2] * fi[2] + drjm[2] * fj[2] + drkm[2] * fk[2]);
  v[3] = (drim[0] * fi[1] + drjm[0] * fj[1] + drkm[0] * fk[1]);
  v[4] = (drim[0] * fi[2] + drjm[0] * fj[2] + drkm[0] * fk[2]);
  v[5] = (drim[1] * fi[2] + drjm[1] * fj[2] + drkm[1] * fk[2]);
  if (vflag_global)
    for (int n = 0; n < 6; n++) virial[n] += v[n];
  if (vflag_atom)
    for (int n = 0; n < 6; n++) {
      vatom[i][n] += 0.25 * v[n];
      vatom[j][n] += 0.25 * v[n];
      vatom[k][n] += 0.25 * v[n];
      vatom[m][n] += 0.25 * v[n];
    }
}

void Pair::virial_fdotr_compute()
{
  double **x = atom->x;
  double **f = atom->f;
  // sum over force on all particles including ghosts
  int nall = atom->nlocal + atom->nghost;
  for (int i = 0; i < nall; i++) {
    virial[0] += f[i][0] * x[i][0];
    virial[1] += f[i][1] * x[i][1];
    virial[2] += f[i][2] * x[i][2];
    virial[3] += f[i][1] * x[i][0];
    virial[4] += f[i][2] * x[i][0];
    virial[5] += f[i][2] * x[i][1];
  }
  // prevent multiple calls to update the virial
  vflag_fdotr = 0;
}

// ------------------------------------------------------------------ NeighList
NeighList::NeighList(LAMMPS *l) :
    Pointers(l), inum(0), gnum(0), ilist(nullptr), numneigh(nullptr), firstneigh(nullptr),
    maxatom(0), ghost(0), ipage(nullptr)
{
}
NeighList::~NeighList()
{
  memory->destroy(ilist);
  memory->destroy(numneigh);
  memory->sfree(firstneigh);
  delete[] ipage;
}
void NeighList::grow(int nlocal, int nall)
{
  // skip if data structs are already big enough
  if (ghost) {
    if (nall <= maxatom) return;
  } else {
    if (nlocal <= maxatom) return;
  }
  if (ghost) maxatom = nall;
  else maxatom = nlocal;
  maxatom = atom->nmax > maxatom ? atom->nmax : maxatom;
  memory->destroy(ilist);
  memory->destroy(numneigh);
  memory->sfree(firstneigh);
  memory->create(ilist, maxatom, "neighlist:ilist");
  memory->create(numneigh, maxatom, "neighlist:numneigh");
  firstneigh = (int **) memory->smalloc((bigint) maxatom * sizeof(int *), "neighlist:firstneigh");
}

// ------------------------------------------------------------------ Force
Force::Force(LAMMPS *l) : Pointers(l)
{
  newton = newton_pair = newton_bond = 1;
  pair = nullptr;
  pair_style = utils::strdup("none");
  pair_map = new PairCreatorMap();
  dielectric = 1.0;
  qqr2e = qe2f = vxmu2f = xxt2kmu = qqrd2e = e_mass = hhmrr2e = mvh2r = angstrom = femtosecond =
      qelectron = hplanck = mv2d = 0.0;
  boltz = mvv2e = ftm2v = nktv2p = 1.0;
}
Force::~Force()
{
  delete[] pair_style;
  delete pair;
  delete pair_map;
}
void Force::init()
{
  if (pair) pair->init();
}
Pair *Force::new_pair(const std::string &style, int, int &sflag)
{
  sflag = 0;
  if (style == "none") return nullptr;
  if (pair_map->find(style) != pair_map->end()) {
    PairCreator &pair_creator = (*pair_map)[style];
    return pair_creator(lmp);
  }
  error->all(FLERR, "Unrecognized pair style '{}' (load its plugin first)", style);
  return nullptr;
}
void Force::create_pair(const std::string &style, int trysuffix)
{
  delete[] pair_style;
  delete pair;
  pair_style = nullptr;
  pair = nullptr;
  int sflag;
  pair = new_pair(style, trysuffix, sflag);
  pair_style = utils::strdup(style);
}
