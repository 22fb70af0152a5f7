/* -*- c++ -*- ---------------------------------------------------------------
   lmpshim -- a minimal stand-in for the LAMMPS-core C++ API surface that the
   pair-style plugins `rebomos` and `aeam` are written against.

   LAMMPS itself (lammps/lammps, stable_2Aug2023) is NOT part of the
   reference repository and is not installed in this image, so the class
   declarations a `Pair` plugin needs (Pointers, Pair, Atom, NeighList,
   Neighbor, Comm, Force, Memory, Error, MyPage, file readers, MathConst,
   MathSpecial, plugin registration structs) are restated here from the
   public LAMMPS API, member for member as far as the two styles touch them
   (census: SURVEY.md section 8b).  With a real LAMMPS checkout the include
   path is switched to ${LAMMPS_SOURCE_DIR} and this directory is unused.

   This is API surface only.  The services behind it (neighbor build, halo
   exchange, integrator, thermo, input parsing) are implemented by the
   mini engine in oracle/engine/ (test infrastructure, stand-in for the
   LAMMPS executable) -- or by LAMMPS itself in production.
---------------------------------------------------------------------------- */
#ifndef LMPSHIM_H
#define LMPSHIM_H

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <sstream>
#include <string>
#include <vector>

// ---------------------------------------------------------------- mpi stub
// Thread-rank "MPI": a communicator is a pointer to a rank context owned by
// the engine's Universe.  Only the calls the two pair styles make exist.
struct lmpshim_rankctx;
typedef struct lmpshim_rankctx *MPI_Comm;
typedef int MPI_Datatype;
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_DOUBLE 8
extern "C" int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);

namespace LAMMPS_NS {

// ---------------------------------------------------------------- lmptype.h
typedef int tagint;
typedef int imageint;
typedef int64_t bigint;
#define NEIGHMASK 0x1FFFFFFF
#define MAXSMALLINT 0x7FFFFFFF
#define FLERR __FILE__, __LINE__
#ifndef MIN
#define MIN(A, B) ((A) < (B) ? (A) : (B))
#endif
#ifndef MAX
#define MAX(A, B) ((A) > (B) ? (A) : (B))
#endif

#define LAMMPS_VERSION "2 Aug 2023"

class LAMMPS;
class Memory;
class Error;
class Universe;
class Input;
class Atom;
class Update;
class Neighbor;
class NeighList;
class NeighRequest;
class Comm;
class Domain;
class Force;
class Modify;
class Output;
class Pair;

// ---------------------------------------------------------------- error.h
namespace shimfmt {
  inline void put(std::ostringstream &os, const char *&f)
  {
    while (*f) os << *f++;
  }
  template <typename T, typename... R>
  void put(std::ostringstream &os, const char *&f, const T &v, const R &...rest)
  {
    while (*f) {
      if (f[0] == '{' && f[1] == '}') {
        os << v;
        f += 2;
        put(os, f, rest...);
        return;
      }
      os << *f++;
    }
  }
  template <typename... A> std::string format(const std::string &fmt, const A &...args)
  {
    std::ostringstream os;
    os.precision(17);
    const char *f = fmt.c_str();
    put(os, f, args...);
    return os.str();
  }
}    // namespace shimfmt

// thrown instead of MPI_Abort so that a host (pytest) survives input errors
class LAMMPSException : public std::exception {
 public:
  std::string message;
  explicit LAMMPSException(const std::string &m) : message(m) {}
  const char *what() const noexcept override { return message.c_str(); }
};

class Error {
 public:
  explicit Error(LAMMPS *) {}
  [[noreturn]] void all(const std::string &file, int line, const std::string &str);
  [[noreturn]] void one(const std::string &file, int line, const std::string &str);
  template <typename... A>
  [[noreturn]] void all(const std::string &file, int line, const std::string &fmt, const A &...a)
  {
    all(file, line, shimfmt::format(fmt, a...));
  }
  template <typename... A>
  [[noreturn]] void one(const std::string &file, int line, const std::string &fmt, const A &...a)
  {
    one(file, line, shimfmt::format(fmt, a...));
  }
  void warning(const std::string &file, int line, const std::string &str);
};

// ---------------------------------------------------------------- memory.h
// contiguous data block + row pointers, as LAMMPS' Memory::create does
class Memory {
 public:
  explicit Memory(LAMMPS *) {}
  void *smalloc(bigint n, const char *) { return n > 0 ? malloc((size_t) n) : nullptr; }
  void *srealloc(void *p, bigint n, const char *) { return realloc(p, (size_t) n); }
  void sfree(void *p) { free(p); }

  template <typename T> T *create(T *&a, int n, const char *name)
  {
    a = (T *) smalloc((bigint) sizeof(T) * n, name);
    return a;
  }
  template <typename T> T *grow(T *&a, int n, const char *name)
  {
    a = (T *) srealloc(a, (bigint) sizeof(T) * n, name);
    return a;
  }
  template <typename T> void destroy(T *&a)
  {
    sfree(a);
    a = nullptr;
  }
  template <typename T> T **create(T **&a, int n1, int n2, const char *name)
  {
    if (n1 <= 0 || n2 <= 0) { a = nullptr; return a; }
    T *data = (T *) smalloc((bigint) sizeof(T) * n1 * n2, name);
    a = (T **) smalloc((bigint) sizeof(T *) * n1, name);
    bigint n = 0;
    for (int i = 0; i < n1; i++) { a[i] = &data[n]; n += n2; }
    return a;
  }
  template <typename T> T **grow(T **&a, int n1, int n2, const char *name)
  {
    if (a == nullptr) return create(a, n1, n2, name);
    T *data = (T *) srealloc(a[0], (bigint) sizeof(T) * n1 * n2, name);
    a = (T **) srealloc(a, (bigint) sizeof(T *) * n1, name);
    bigint n = 0;
    for (int i = 0; i < n1; i++) { a[i] = &data[n]; n += n2; }
    return a;
  }
  template <typename T> void destroy(T **&a)
  {
    if (a == nullptr) return;
    sfree(a[0]);
    sfree(a);
    a = nullptr;
  }
  template <typename T> T ***create(T ***&a, int n1, int n2, int n3, const char *name)
  {
    if (n1 <= 0 || n2 <= 0 || n3 <= 0) { a = nullptr; return a; }
    T *data = (T *) smalloc((bigint) sizeof(T) * n1 * n2 * n3, name);
    T **plane = (T **) smalloc((bigint) sizeof(T *) * n1 * n2, name);
    a = (T ***) smalloc((bigint) sizeof(T **) * n1, name);
    bigint n = 0, m;
    for (int i = 0; i < n1; i++) {
      m = (bigint) i * n2;
      a[i] = &plane[m];
      for (int j = 0; j < n2; j++) { plane[m + j] = &data[n]; n += n3; }
    }
    return a;
  }
  template <typename T> void destroy(T ***&a)
  {
    if (a == nullptr) return;
    sfree(a[0][0]);
    sfree(a[0]);
    sfree(a);
    a = nullptr;
  }
};

// ---------------------------------------------------------------- my_page.h
template <class T> class MyPage {
 public:
  int ndatum, nchunk;
  MyPage() : ndatum(0), nchunk(0), page(nullptr), pages(nullptr), npage(0), ipage(0), index(0),
             maxchunk(1), pagesize(1024), pagedelta(1), errorflag(0) {}
  ~MyPage() { deallocate(); }
  int init(int user_maxchunk = 1, int user_pagesize = 1024, int user_pagedelta = 1)
  {
    maxchunk = user_maxchunk;
    pagesize = user_pagesize;
    pagedelta = user_pagedelta;
    if (maxchunk <= 0 || pagesize <= 0 || pagedelta <= 0) return 1;
    if (maxchunk > pagesize) return 1;
    deallocate();
    allocate();
    if (errorflag) return 2;
    reset();
    return 0;
  }
  T *vget()
  {
    if (index + maxchunk <= pagesize) return &page[index];
    ipage++;
    if (ipage == npage) {
      allocate();
      if (errorflag) return nullptr;
    }
    page = pages[ipage];
    index = 0;
    return &page[index];
  }
  void vgot(int n)
  {
    if (n > maxchunk) errorflag = 1;
    ndatum += n;
    nchunk++;
    index += n;
  }
  void reset()
  {
    ndatum = nchunk = 0;
    index = ipage = 0;
    page = (pages != nullptr) ? pages[ipage] : nullptr;
    errorflag = 0;
  }
  double size() const { return (double) npage * pagesize * sizeof(T); }
  int status() const { return errorflag; }

 private:
  T *page;
  T **pages;
  int npage, ipage, index;
  int maxchunk, pagesize, pagedelta, errorflag;
  void allocate()
  {
    npage += pagedelta;
    pages = (T **) realloc(pages, npage * sizeof(T *));
    if (!pages) { errorflag = 2; return; }
    for (int i = npage - pagedelta; i < npage; i++) {
      pages[i] = (T *) malloc((size_t) pagesize * sizeof(T));
      if (!pages[i]) errorflag = 2;
    }
  }
  void deallocate()
  {
    for (int i = 0; i < npage; i++) free(pages[i]);
    free(pages);
    pages = nullptr;
    npage = 0;
  }
};

// ---------------------------------------------------------------- math_const.h / math_special.h
namespace MathConst {
  static constexpr double THIRD = 1.0 / 3.0;
  static constexpr double MY_PI = 3.14159265358979323846;
  static constexpr double MY_2PI = 6.28318530717958647692;
  static constexpr double MY_PI2 = 1.57079632679489661923;
}    // namespace MathConst
namespace MathSpecial {
  static inline double square(const double &x) { return x * x; }
  static inline double cube(const double &x) { return x * x * x; }
  static inline double powint(const double &x, const int n)
  {
    double yy, ww;
    if (n == 0) return 1.0;
    if (x == 0.0) return 0.0;
    int nn = (n > 0) ? n : -n;
    ww = x;
    for (yy = 1.0; nn != 0; nn >>= 1, ww *= ww)
      if (nn & 1) yy *= ww;
    return (n > 0) ? yy : 1.0 / yy;
  }
}    // namespace MathSpecial

// ---------------------------------------------------------------- tokenizer.h / text_file_reader.h
class TokenizerException : public std::exception {
  std::string message;
 public:
  TokenizerException(const std::string &msg, const std::string &token);
  const char *what() const noexcept override { return message.c_str(); }
};
class InvalidIntegerException : public TokenizerException {
 public:
  explicit InvalidIntegerException(const std::string &token) :
      TokenizerException("Not a valid integer number", token) {}
};
class InvalidFloatException : public TokenizerException {
 public:
  explicit InvalidFloatException(const std::string &token) :
      TokenizerException("Not a valid floating-point number", token) {}
};
class FileReaderException : public std::exception {
  std::string message;
 public:
  explicit FileReaderException(const std::string &msg) : message(msg) {}
  const char *what() const noexcept override { return message.c_str(); }
};
class EOFException : public FileReaderException {
 public:
  explicit EOFException(const std::string &msg) : FileReaderException(msg) {}
};

class ValueTokenizer {
  std::vector<std::string> tokens;
  size_t pos;
 public:
  explicit ValueTokenizer(const std::string &str, const std::string &separators = " \t\r\n\f");
  bool has_next() const { return pos < tokens.size(); }
  size_t count() const { return tokens.size(); }
  std::string next_string();
  int next_int();
  bigint next_bigint();
  tagint next_tagint();
  double next_double();
  void skip(int n = 1) { pos += n; }
};

class TextFileReader {
  std::string filetype;
  bool closefp;
  static constexpr int MAXLINE = 1024;
  char line[MAXLINE];
  FILE *fp;
 public:
  bool ignore_comments;
  TextFileReader(const std::string &filename, const std::string &filetype);
  TextFileReader(FILE *fp, std::string filetype);
  ~TextFileReader();
  void skip_line();
  char *next_line(int nparams = 0);
  void next_dvector(double *list, int n);
  ValueTokenizer next_values(int nparams, const std::string &separators = " \t\r\n\f");
  void rewind() { ::rewind(fp); }
};

class PotentialFileReader {
 protected:
  LAMMPS *lmp;
  TextFileReader *reader;
  std::string filename, filetype;
  int unit_convert;
 public:
  PotentialFileReader(LAMMPS *lmp, const std::string &filename, const std::string &potential_name,
                      const int auto_convert = 0);
  PotentialFileReader(LAMMPS *lmp, const std::string &filename, const std::string &potential_name,
                      const std::string &name_suffix, const int auto_convert = 0);
  ~PotentialFileReader();
  void ignore_comments(bool value) { reader->ignore_comments = value; }
  void skip_line() { reader->skip_line(); }
  char *next_line(int nparams = 0) { return reader->next_line(nparams); }
  void next_dvector(double *list, int n) { reader->next_dvector(list, n); }
  ValueTokenizer next_values(int nparams, const std::string &separators = " \t\r\n\f")
  {
    return reader->next_values(nparams, separators);
  }
  double next_double();
  int next_int();
  std::string next_string();
  int get_unit_convert() const { return unit_convert; }
};

// ---------------------------------------------------------------- utils.h
namespace utils {
  enum { NOCONVERT = 0, METAL2REAL = 1, REAL2METAL = 1 << 1 };
  enum { UNKNOWN = 0, ENERGY };
  FILE *open_potential(const std::string &name, LAMMPS *lmp, int *auto_convert);
  std::string getsyserror();
  char *strdup(const std::string &text);
  int get_supported_conversions(const int property);
  std::string get_potential_date(const std::string &path, const std::string &potential_name);
  std::string get_potential_units(const std::string &path, const std::string &potential_name);
  bool is_integer(const std::string &str);
  bool is_double(const std::string &str);
  double numeric(const char *file, int line, const std::string &str, bool do_abort, LAMMPS *lmp);
  int inumeric(const char *file, int line, const std::string &str, bool do_abort, LAMMPS *lmp);
}    // namespace utils

// ---------------------------------------------------------------- lammps.h / pointers.h
class LAMMPS {
 public:
  Memory *memory;
  Error *error;
  Universe *universe;
  Input *input;
  Atom *atom;
  Update *update;
  Neighbor *neighbor;
  Comm *comm;
  Domain *domain;
  Force *force;
  Modify *modify;
  Output *output;
  MPI_Comm world;
  FILE *infile, *screen, *logfile;
  LAMMPS() : memory(nullptr), error(nullptr), universe(nullptr), input(nullptr), atom(nullptr),
             update(nullptr), neighbor(nullptr), comm(nullptr), domain(nullptr), force(nullptr),
             modify(nullptr), output(nullptr), world(nullptr), infile(nullptr), screen(nullptr),
             logfile(nullptr) {}
};

class Pointers {
 public:
  Pointers(LAMMPS *ptr) :
      lmp(ptr), memory(ptr->memory), error(ptr->error), universe(ptr->universe), input(ptr->input),
      atom(ptr->atom), update(ptr->update), neighbor(ptr->neighbor), comm(ptr->comm),
      domain(ptr->domain), force(ptr->force), modify(ptr->modify), output(ptr->output),
      world(ptr->world), infile(ptr->infile), screen(ptr->screen), logfile(ptr->logfile) {}
  virtual ~Pointers() = default;
  Pointers() = delete;
  Pointers(const Pointers &) = default;

 protected:
  LAMMPS *lmp;
  Memory *&memory;
  Error *&error;
  Universe *&universe;
  Input *&input;
  Atom *&atom;
  Update *&update;
  Neighbor *&neighbor;
  Comm *&comm;
  Domain *&domain;
  Force *&force;
  Modify *&modify;
  Output *&output;
  MPI_Comm &world;
  FILE *&infile;
  FILE *&screen;
  FILE *&logfile;
};

// ---------------------------------------------------------------- lattice.h / region (engine side)
class Lattice {
 public:
  double xlattice, ylattice, zlattice;
  double a1[3], a2[3], a3[3], origin[3], scale;
  std::vector<std::vector<double>> basis;
  Lattice();
  void setup();
  void lattice2box(double &x, double &y, double &z) const;
  void box2lattice(double &x, double &y, double &z) const;
  void bbox(int flag, double x, double y, double z, double &xmin, double &ymin, double &zmin,
            double &xmax, double &ymax, double &zmax) const;

 private:
  double primitive[3][3], priminv[3][3];
};

struct Region {
  std::string id, style;
  double xlo, xhi, ylo, yhi, zlo, zhi, xy, xz, yz;
};

// ---------------------------------------------------------------- domain.h
class Domain : protected Pointers {
 public:
  int box_exist, box_change, dimension, triclinic;
  int periodicity[3], xperiodic, yperiodic, zperiodic;
  double boxlo[3], boxhi[3], xy, xz, yz;
  double prd[3], prd_half[3], xprd, yprd, zprd;
  double h[6], h_inv[6];
  double boxlo_lamda[3], boxhi_lamda[3], prd_lamda[3];
  double boxlo_bound[3], boxhi_bound[3];
  double sublo[3], subhi[3], sublo_lamda[3], subhi_lamda[3];
  Lattice *lattice;
  std::vector<Region> regions;

  explicit Domain(LAMMPS *lmp);
  ~Domain() override;
  void set_global_box();
  void set_local_box();
  void x2lamda(int n);
  void lamda2x(int n);
  void x2lamda(const double *x, double *lamda) const;
  void lamda2x(const double *lamda, double *x) const;
  void bbox(const double *lo, const double *hi, double *bboxlo, double *bboxhi) const;
  void pbc();
  void remap(double *x) const;
  double volume() const { return xprd * yprd * zprd; }
};

// ---------------------------------------------------------------- atom.h (fields the styles touch)
class Atom : protected Pointers {
 public:
  bigint natoms;
  int nlocal, nghost, nmax;
  int ntypes;
  int tag_enable;
  tagint *tag;
  int *type, *mask;
  double **x, **v, **f;
  double *mass;
  int *mass_setflag;
  // spatial sorting (atom_modify sort)
  int sortfreq;
  bigint nextsort;
  double userbinsize;

  explicit Atom(LAMMPS *lmp);
  ~Atom() override;
  virtual void set_mass(const char *file, int line, int itype, double value);
  // engine side
  void allocate_type_arrays(int n);
  void grow(int n);
  void copy(int i, int j);
  void create_atom(int itype, const double *coord, tagint t);
  void setup_sort_bins();
  void sort();

  // sort bins
  int nbins, nbinx, nbiny, nbinz, maxbin, maxnext;
  int *binhead, *next, *permute;
  double bininvx, bininvy, bininvz, bboxlo[3], bboxhi[3];
};

// ---------------------------------------------------------------- neigh_list.h / neighbor.h
namespace NeighConst {
  enum {
    REQ_DEFAULT = 0,
    REQ_FULL = 1 << 0,
    REQ_GHOST = 1 << 1,
    REQ_SIZE = 1 << 2,
    REQ_HISTORY = 1 << 3,
    REQ_OCCASIONAL = 1 << 4
  };
}

class NeighList : protected Pointers {
 public:
  int inum, gnum;
  int *ilist, *numneigh;
  int **firstneigh;
  int maxatom;
  int ghost;    // 1 if list stores neighbors of ghosts
  MyPage<int> *ipage;
  explicit NeighList(LAMMPS *lmp);
  ~NeighList() override;
  void grow(int nlocal, int nall);
};

class NeighRequest : protected Pointers {
 public:
  void *requestor;
  int full, ghost;
  NeighRequest(LAMMPS *lmp, void *req, int flags) :
      Pointers(lmp), requestor(req), full(flags & NeighConst::REQ_FULL ? 1 : 0),
      ghost(flags & NeighConst::REQ_GHOST ? 1 : 0) {}
};

class Neighbor : protected Pointers {
 public:
  int style;
  int every, delay, dist_check, ago;
  int pgsize, oneatom;
  double skin, cutneighmin, cutneighmax;
  double **cutneighsq, **cutneighghostsq;
  bigint ncalls, ndanger, lastcall;
  int includegroup;

  explicit Neighbor(LAMMPS *lmp);
  ~Neighbor() override;
  NeighRequest *add_request(Pair *pair, int flags = 0);

  // engine side (restated LAMMPS-core: NBinStandard, NStencilFull[Ghost]Bin3d, NPairFullBin[Ghost])
  void init();
  void setup_bins();
  int decide();
  int check_distance();
  void build(int topoflag = 1);
  bigint memory_usage();

  NeighList *list;
  NeighRequest *request;
  // bins
  int nbinx, nbiny, nbinz, mbins, mbinx, mbiny, mbinz, mbinxlo, mbinylo, mbinzlo;
  double binsizex, binsizey, binsizez, bininvx, bininvy, bininvz;
  double bboxlo[3], bboxhi[3];
  int *binhead, *bins, *atom2bin;
  int maxbin, maxatombin;
  // stencil
  int nstencil, maxstencil, sx, sy, sz;
  int *stencil;
  int (*stencilxyz)[3];
  // displacement check
  double **xhold;
  int maxhold;
  double triggersq;
  int coord2bin(const double *x) const;
  int coord2bin(const double *x, int &ix, int &iy, int &iz) const;
  void bin_atoms();
  void create_stencil();
  double bin_distance(int i, int j, int k) const;
};

// ---------------------------------------------------------------- comm.h
class Comm : protected Pointers {
 public:
  int me, nprocs;
  int nthreads;
  int procgrid[3], user_procgrid[3], myloc[3], procneigh[3][2];
  double cutghost[3];
  double cutghostuser;
  int ghost_velocity;
  int maxexchange_atom;

  explicit Comm(LAMMPS *lmp);
  ~Comm() override;
  // called by pair styles
  virtual void forward_comm(Pair *pair);
  virtual void reverse_comm(Pair *pair);
  // engine side (restated CommBrick)
  void set_proc_grid();
  void init();
  void setup();
  void forward_comm();
  void reverse_comm();
  void exchange();
  void borders();
  double get_comm_cutoff();

  int nswap, maxswap;
  int maxneed[3];
  int *sendnum, *recvnum, *sendproc, *recvproc, *firstrecv, *pbc_flag;
  int (*pbc)[6];
  double *slablo, *slabhi;
  std::vector<std::vector<int>> sendlist;
  std::vector<double> buf_send, buf_recv;
  bigint bytes_forward, bytes_reverse;    // per-call traffic accounting (all swaps)
};

// ---------------------------------------------------------------- force.h
class Force : protected Pointers {
 public:
  double boltz, hplanck, mvv2e, ftm2v, mv2d, nktv2p, qqr2e, qe2f, vxmu2f, xxt2kmu, dielectric, qqrd2e,
      e_mass, hhmrr2e, mvh2r, angstrom, femtosecond, qelectron;
  int newton, newton_pair, newton_bond;
  Pair *pair;
  char *pair_style;
  typedef Pair *(*PairCreator)(LAMMPS *);
  typedef std::map<std::string, PairCreator> PairCreatorMap;
  PairCreatorMap *pair_map;

  explicit Force(LAMMPS *lmp);
  ~Force() override;
  void init();
  void create_pair(const std::string &style, int trysuffix);
  Pair *new_pair(const std::string &style, int trysuffix, int &sflag);
};

// ---------------------------------------------------------------- pair.h
class Pair : protected Pointers {
 public:
  static int instance_total;

  double eng_vdwl, eng_coul;
  double virial[6];
  double *eatom, **vatom, **cvatom;

  double cutforce;
  double **cutsq;
  int **setflag;

  int comm_forward, comm_reverse, comm_reverse_off;
  int single_enable, born_matrix_enable, single_hessian_enable, restartinfo, respa_enable, one_coeff,
      manybody_flag, unit_convert_flag, no_virial_fdotr_compute, writedata, finitecutflag, ghostneigh;
  double **cutghost;
  int ewaldflag, pppmflag, msmflag, dispersionflag, tip4pflag, dipoleflag, spinflag, reinitflag;
  int centroidstressflag;
  int tail_flag;
  double etail, ptail, etail_ij, ptail_ij;
  int trim_flag;
  int evflag, eflag_either, eflag_global, eflag_atom, vflag_either, vflag_global, vflag_atom,
      cvflag_atom;
  int ncoultablebits, ndisptablebits;
  int nextra;
  double *pvector;
  int single_extra;
  double *svector;
  class NeighList *list, *listhalf, *listfull;
  int allocated;
  int compute_flag;
  int mixed_flag;
  bool did_mix;

  enum { GEOMETRIC, ARITHMETIC, SIXTHPOWER };
  int beyond_contact, nondefault_history_transfer;

  Pair(LAMMPS *);
  ~Pair() override;

  // top-level Pair methods
  void init();
  virtual void reinit() {}
  virtual void setup() {}
  double mix_energy(double, double, double, double);
  double mix_distance(double, double);
  void ev_tally(int, int, int, int, double, double, double, double, double, double);
  void ev_tally3(int, int, int, double, double, double *, double *, double *, double *);
  void v_tally2(int, int, double, double *);
  void v_tally3(int, int, int, double *, double *, double *, double *);
  void v_tally4(int, int, int, int, double *, double *, double *, double *, double *, double *);

  // general child-class methods
  virtual void compute(int, int) = 0;
  virtual void compute_inner() {}
  virtual void compute_middle() {}
  virtual void compute_outer(int, int) {}
  virtual double single(int, int, int, int, double, double, double, double &fforce)
  {
    fforce = 0.0;
    return 0.0;
  }
  virtual void settings(int, char **) = 0;
  virtual void coeff(int, char **) = 0;
  virtual void init_style();
  virtual void init_list(int, class NeighList *);
  virtual double init_one(int, int) { return 0.0; }

  virtual int pack_forward_comm(int, int *, double *, int, int *) { return 0; }
  virtual void unpack_forward_comm(int, int, double *) {}
  virtual int pack_reverse_comm(int, int, double *) { return 0; }
  virtual void unpack_reverse_comm(int, int *, double *) {}
  virtual double memory_usage();

  // specific child-class methods for certain Pair styles
  virtual void *extract(const char *, int &) { return nullptr; }

  enum { ENERGY_NONE = 0, ENERGY_GLOBAL = 1, ENERGY_ATOM = 2 };
  enum { VIRIAL_NONE = 0, VIRIAL_PAIR = 1, VIRIAL_FDOTR = 2, VIRIAL_ATOM = 4, VIRIAL_CENTROID = 8 };
  enum { CENTROID_SAME = 0, CENTROID_AVAIL = 1, CENTROID_NOTAVAIL = 2 };

 protected:
  int vflag_fdotr;
  int maxeatom, maxvatom, maxcvatom;
  int copymode;
  int *map;    // used by many-body styles to map atom types to elements
  int suffix_flag;

  virtual void ev_setup(int, int, int alloc = 1);
  void ev_init(int eflag, int vflag, int alloc = 1)
  {
    if (eflag || vflag)
      ev_setup(eflag, vflag, alloc);
    else
      ev_unset();
  }
  void ev_unset()
  {
    evflag = vflag_fdotr = 0;
    eflag_either = eflag_global = eflag_atom = 0;
    vflag_either = vflag_global = vflag_atom = cvflag_atom = 0;
  }
  void virial_fdotr_compute();
};

}    // namespace LAMMPS_NS

// ---------------------------------------------------------------- lammpsplugin.h
extern "C" {
typedef void *(lammpsplugin_factory1) (void *);
typedef void *(lammpsplugin_factory2) (void *, int, char **);
typedef struct {
  const char *version;
  const char *style;
  const char *name;
  const char *info;
  const char *author;
  union {
    lammpsplugin_factory1 *v1;
    lammpsplugin_factory2 *v2;
  } creator;
  void *handle;
} lammpsplugin_t;
typedef void (*lammpsplugin_regfunc)(lammpsplugin_t *, void *);
typedef void (*lammpsplugin_initfunc)(void *, void *, void *);
// the one symbol a plugin exports
void lammpsplugin_init(void *, void *, void *);
}

#endif
