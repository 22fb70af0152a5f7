// minilmp engine (test infrastructure): Atom (incl. spatial sort), Domain,
// Lattice.  Restated LAMMPS-core semantics: src/atom.cpp (sort,
// setup_sort_bins), src/domain.cpp (set_global_box, x2lamda, lamda2x, bbox,
// pbc, remap), src/lattice.cpp -- see SURVEY.md A.4, A.6.

#include "engine.h"

#include <algorithm>

using namespace LAMMPS_NS;

static constexpr double BIG = 1.0e30;

// ================================================================== Atom
Atom::Atom(LAMMPS *l) : Pointers(l)
{
  natoms = 0;
  nlocal = nghost = nmax = 0;
  ntypes = 0;
  tag_enable = 1;
  tag = nullptr;
  type = mask = nullptr;
  x = v = f = nullptr;
  mass = nullptr;
  mass_setflag = nullptr;
  sortfreq = 1000;
  nextsort = 0;
  userbinsize = 0.0;
  nbins = nbinx = nbiny = nbinz = 0;
  maxbin = maxnext = 0;
  binhead = next = permute = nullptr;
}

Atom::~Atom()
{
  memory->destroy(tag);
  memory->destroy(type);
  memory->destroy(mask);
  memory->destroy(x);
  memory->destroy(v);
  memory->destroy(f);
  delete[] mass;
  delete[] mass_setflag;
  memory->destroy(binhead);
  memory->destroy(next);
  memory->destroy(permute);
}

void Atom::allocate_type_arrays(int n)
{
  ntypes = n;
  delete[] mass;
  delete[] mass_setflag;
  mass = new double[n + 1];
  mass_setflag = new int[n + 1];
  for (int i = 0; i <= n; i++) { mass[i] = 0.0; mass_setflag[i] = 0; }
}

void Atom::set_mass(const char *file, int line, int itype, double value)
{
  if (mass == nullptr) error->all(file, line, "Cannot set mass for atom style atomic before box");
  if (itype < 1 || itype > ntypes) error->all(file, line, "Invalid type {} for atom mass {}", itype, value);
  if (value <= 0.0) error->all(file, line, "Invalid atom mass value {}", value);
  mass[itype] = value;
  mass_setflag[itype] = 1;
}

void Atom::grow(int n)
{
  // DELTA growth like AtomVec::grow (nmax multiple of 16384)
  const int DELTA = 16384;
  if (n == 0) n = nmax + DELTA;
  if (n <= nmax) return;
  nmax = (n / DELTA + 1) * DELTA;
  tag = memory->grow(tag, nmax, "atom:tag");
  type = memory->grow(type, nmax, "atom:type");
  mask = memory->grow(mask, nmax, "atom:mask");
  memory->grow(x, nmax, 3, "atom:x");
  memory->grow(v, nmax, 3, "atom:v");
  memory->grow(f, nmax, 3, "atom:f");
}

void Atom::copy(int i, int j)
{
  tag[j] = tag[i];
  type[j] = type[i];
  mask[j] = mask[i];
  for (int d = 0; d < 3; d++) {
    x[j][d] = x[i][d];
    v[j][d] = v[i][d];
  }
}

void Atom::create_atom(int itype, const double *coord, tagint t)
{
  if (nlocal == nmax) grow(0);
  tag[nlocal] = t;
  type[nlocal] = itype;
  mask[nlocal] = 1;
  for (int d = 0; d < 3; d++) {
    x[nlocal][d] = coord[d];
    v[nlocal][d] = 0.0;
  }
  nlocal++;
}

// Atom::setup_sort_bins (src/atom.cpp): bins of 1/2 neighbor cutoff over my sub-domain bbox
void Atom::setup_sort_bins()
{
  double binsize = 0.0;
  if (userbinsize > 0.0) binsize = userbinsize;
  else if (neighbor->cutneighmax > 0.0) binsize = 0.5 * neighbor->cutneighmax;
  if (binsize == 0.0 && sortfreq > 0) { sortfreq = 0; return; }
  double bininv = 1.0 / binsize;

  if (domain->triclinic)
    domain->bbox(domain->sublo_lamda, domain->subhi_lamda, bboxlo, bboxhi);
  else
    for (int d = 0; d < 3; d++) { bboxlo[d] = domain->sublo[d]; bboxhi[d] = domain->subhi[d]; }

  nbinx = static_cast<int>((bboxhi[0] - bboxlo[0]) * bininv);
  nbiny = static_cast<int>((bboxhi[1] - bboxlo[1]) * bininv);
  nbinz = static_cast<int>((bboxhi[2] - bboxlo[2]) * bininv);
  if (nbinx == 0) nbinx = 1;
  if (nbiny == 0) nbiny = 1;
  if (nbinz == 0) nbinz = 1;
  bininvx = nbinx / (bboxhi[0] - bboxlo[0]);
  bininvy = nbiny / (bboxhi[1] - bboxlo[1]);
  bininvz = nbinz / (bboxhi[2] - bboxlo[2]);

  bigint nb = (bigint) nbinx * nbiny * nbinz;
  if (nb > MAXSMALLINT) error->one(FLERR, "Too many atom sorting bins");
  nbins = (int) nb;
  if (nbins > maxbin) {
    memory->destroy(binhead);
    maxbin = nbins;
    memory->create(binhead, maxbin, "atom:binhead");
  }
}

// Atom::sort (src/atom.cpp): reorder owned atoms bin by bin, ascending index within a bin
void Atom::sort()
{
  int i, m, n, ix, iy, iz, ibin;

  nextsort = (update->ntimestep / sortfreq) * sortfreq + sortfreq;
  setup_sort_bins();
  if (sortfreq == 0 || nbins == 1) return;

  if (nlocal > maxnext) {
    memory->destroy(next);
    memory->destroy(permute);
    maxnext = nmax;
    memory->create(next, maxnext, "atom:next");
    memory->create(permute, maxnext, "atom:permute");
  }
  for (i = 0; i < nbins; i++) binhead[i] = -1;

  // for triclinic, atoms must be in box coords (not lamda) to match bbox
  if (domain->triclinic) domain->lamda2x(nlocal);

  for (i = nlocal - 1; i >= 0; i--) {
    ix = static_cast<int>((x[i][0] - bboxlo[0]) * bininvx);
    iy = static_cast<int>((x[i][1] - bboxlo[1]) * bininvy);
    iz = static_cast<int>((x[i][2] - bboxlo[2]) * bininvz);
    ix = MAX(ix, 0);
    iy = MAX(iy, 0);
    iz = MAX(iz, 0);
    ix = MIN(ix, nbinx - 1);
    iy = MIN(iy, nbiny - 1);
    iz = MIN(iz, nbinz - 1);
    ibin = iz * nbiny * nbinx + iy * nbinx + ix;
    next[i] = binhead[ibin];
    binhead[ibin] = i;
  }

  if (domain->triclinic) domain->x2lamda(nlocal);

  // permute[I] = J means Ith new atom will be Jth old atom
  n = 0;
  for (m = 0; m < nbins; m++) {
    i = binhead[m];
    while (i >= 0) {
      permute[n++] = i;
      i = next[i];
    }
  }

  // apply the permutation (out-of-place; same result as LAMMPS' in-place cycle walk)
  std::vector<double> xs(3 * (size_t) nlocal), vs(3 * (size_t) nlocal);
  std::vector<int> ts(nlocal), gs(nlocal), ms(nlocal);
  for (i = 0; i < nlocal; i++) {
    int j = permute[i];
    for (int d = 0; d < 3; d++) { xs[3 * i + d] = x[j][d]; vs[3 * i + d] = v[j][d]; }
    ts[i] = type[j];
    gs[i] = tag[j];
    ms[i] = mask[j];
  }
  for (i = 0; i < nlocal; i++) {
    for (int d = 0; d < 3; d++) { x[i][d] = xs[3 * i + d]; v[i][d] = vs[3 * i + d]; }
    type[i] = ts[i];
    tag[i] = gs[i];
    mask[i] = ms[i];
  }
}

// ================================================================== Lattice
Lattice::Lattice() : xlattice(1.0), ylattice(1.0), zlattice(1.0), scale(1.0)
{
  for (int d = 0; d < 3; d++) a1[d] = a2[d] = a3[d] = origin[d] = 0.0;
  a1[0] = a2[1] = a3[2] = 1.0;
}

void Lattice::setup()
{
  // primitive = columns a1 a2 a3 (orient = identity, spacing not user-set)
  for (int r = 0; r < 3; r++) {
    primitive[r][0] = a1[r];
    primitive[r][1] = a2[r];
    primitive[r][2] = a3[r];
  }
  double det = primitive[0][0] * primitive[1][1] * primitive[2][2] +
      primitive[0][1] * primitive[1][2] * primitive[2][0] +
      primitive[0][2] * primitive[1][0] * primitive[2][1] -
      primitive[0][0] * primitive[1][2] * primitive[2][1] -
      primitive[0][1] * primitive[1][0] * primitive[2][2] -
      primitive[0][2] * primitive[1][1] * primitive[2][0];
  priminv[0][0] = (primitive[1][1] * primitive[2][2] - primitive[1][2] * primitive[2][1]) / det;
  priminv[1][0] = (primitive[1][2] * primitive[2][0] - primitive[1][0] * primitive[2][2]) / det;
  priminv[2][0] = (primitive[1][0] * primitive[2][1] - primitive[1][1] * primitive[2][0]) / det;
  priminv[0][1] = (primitive[0][2] * primitive[2][1] - primitive[0][1] * primitive[2][2]) / det;
  priminv[1][1] = (primitive[0][0] * primitive[2][2] - primitive[0][2] * primitive[2][0]) / det;
  priminv[2][1] = (primitive[0][1] * primitive[2][0] - primitive[0][0] * primitive[2][1]) / det;
  priminv[0][2] = (primitive[0][1] * primitive[1][2] - primitive[0][2] * primitive[1][1]) / det;
  priminv[1][2] = (primitive[0][2] * primitive[1][0] - primitive[0][0] * primitive[1][2]) / det;
  priminv[2][2] = (primitive[0][0] * primitive[1][1] - primitive[0][1] * primitive[1][0]) / det;

  // lattice spacings = bounding box of the unit cell (origin not applied while measuring)
  double xmin, ymin, zmin, xmax, ymax, zmax;
  xmin = ymin = zmin = BIG;
  xmax = ymax = zmax = -BIG;
  xlattice = ylattice = zlattice = 0.0;
  for (int k = 0; k <= 1; k++)
    for (int j = 0; j <= 1; j++)
      for (int i = 0; i <= 1; i++) bbox(0, i, j, k, xmin, ymin, zmin, xmax, ymax, zmax);
  xlattice = xmax - xmin;
  ylattice = ymax - ymin;
  zlattice = zmax - zmin;
}

void Lattice::lattice2box(double &x, double &y, double &z) const
{
  double x1 = primitive[0][0] * x + primitive[0][1] * y + primitive[0][2] * z;
  double y1 = primitive[1][0] * x + primitive[1][1] * y + primitive[1][2] * z;
  double z1 = primitive[2][0] * x + primitive[2][1] * y + primitive[2][2] * z;
  x1 *= scale;
  y1 *= scale;
  z1 *= scale;
  x = x1 + xlattice * origin[0];
  y = y1 + ylattice * origin[1];
  z = z1 + zlattice * origin[2];
}

void Lattice::box2lattice(double &x, double &y, double &z) const
{
  x -= xlattice * origin[0];
  y -= ylattice * origin[1];
  z -= zlattice * origin[2];
  x /= scale;
  y /= scale;
  z /= scale;
  double x1 = priminv[0][0] * x + priminv[0][1] * y + priminv[0][2] * z;
  double y1 = priminv[1][0] * x + priminv[1][1] * y + priminv[1][2] * z;
  double z1 = priminv[2][0] * x + priminv[2][1] * y + priminv[2][2] * z;
  x = x1;
  y = y1;
  z = z1;
}

void Lattice::bbox(int flag, double x, double y, double z, double &xmin, double &ymin, double &zmin,
                   double &xmax, double &ymax, double &zmax) const
{
  if (flag == 0) lattice2box(x, y, z);
  else box2lattice(x, y, z);
  xmin = MIN(x, xmin);
  ymin = MIN(y, ymin);
  zmin = MIN(z, zmin);
  xmax = MAX(x, xmax);
  ymax = MAX(y, ymax);
  zmax = MAX(z, zmax);
}

// ================================================================== Domain
Domain::Domain(LAMMPS *l) : Pointers(l)
{
  box_exist = 0;
  box_change = 0;
  dimension = 3;
  triclinic = 0;
  xperiodic = yperiodic = zperiodic = 1;
  for (int d = 0; d < 3; d++) {
    periodicity[d] = 1;
    boxlo[d] = -0.5;
    boxhi[d] = 0.5;
    boxlo_lamda[d] = 0.0;
    boxhi_lamda[d] = 1.0;
    prd_lamda[d] = 1.0;
  }
  xy = xz = yz = 0.0;
  for (double &hh : h) hh = 0.0;
  for (double &hh : h_inv) hh = 0.0;
  lattice = new Lattice();
}
Domain::~Domain() { delete lattice; }

void Domain::set_global_box()
{
  prd[0] = xprd = boxhi[0] - boxlo[0];
  prd[1] = yprd = boxhi[1] - boxlo[1];
  prd[2] = zprd = boxhi[2] - boxlo[2];
  h[0] = xprd;
  h[1] = yprd;
  h[2] = zprd;
  h_inv[0] = 1.0 / h[0];
  h_inv[1] = 1.0 / h[1];
  h_inv[2] = 1.0 / h[2];
  for (int d = 0; d < 3; d++) prd_half[d] = 0.5 * prd[d];

  if (triclinic) {
    h[3] = yz;
    h[4] = xz;
    h[5] = xy;
    h_inv[3] = -h[3] / (h[1] * h[2]);
    h_inv[4] = (h[3] * h[5] - h[1] * h[4]) / (h[0] * h[1] * h[2]);
    h_inv[5] = -h[5] / (h[0] * h[1]);

    boxlo_bound[0] = MIN(boxlo[0], boxlo[0] + xy);
    boxlo_bound[0] = MIN(boxlo_bound[0], boxlo_bound[0] + xz);
    boxlo_bound[1] = MIN(boxlo[1], boxlo[1] + yz);
    boxlo_bound[2] = boxlo[2];
    boxhi_bound[0] = MAX(boxhi[0], boxhi[0] + xy);
    boxhi_bound[0] = MAX(boxhi_bound[0], boxhi_bound[0] + xz);
    boxhi_bound[1] = MAX(boxhi[1], boxhi[1] + yz);
    boxhi_bound[2] = boxhi[2];
  }
}

void Domain::set_local_box()
{
  const int *myloc = comm->myloc;
  const int *procgrid = comm->procgrid;
  if (triclinic == 0) {
    for (int d = 0; d < 3; d++) {
      sublo[d] = boxlo[d] + prd[d] * (myloc[d] * 1.0 / procgrid[d]);
      if (myloc[d] < procgrid[d] - 1) subhi[d] = boxlo[d] + prd[d] * ((myloc[d] + 1) * 1.0 / procgrid[d]);
      else subhi[d] = boxhi[d];
    }
  } else {
    for (int d = 0; d < 3; d++) {
      sublo_lamda[d] = myloc[d] * 1.0 / procgrid[d];
      subhi_lamda[d] = (myloc[d] < procgrid[d] - 1) ? (myloc[d] + 1) * 1.0 / procgrid[d] : 1.0;
    }
  }
}

void Domain::x2lamda(int n)
{
  double delta[3];
  double **x = atom->x;
  for (int i = 0; i < n; i++) {
    delta[0] = x[i][0] - boxlo[0];
    delta[1] = x[i][1] - boxlo[1];
    delta[2] = x[i][2] - boxlo[2];
    x[i][0] = h_inv[0] * delta[0] + h_inv[5] * delta[1] + h_inv[4] * delta[2];
    x[i][1] = h_inv[1] * delta[1] + h_inv[3] * delta[2];
    x[i][2] = h_inv[2] * delta[2];
  }
}
void Domain::lamda2x(int n)
{
  double **x = atom->x;
  for (int i = 0; i < n; i++) {
    x[i][0] = h[0] * x[i][0] + h[5] * x[i][1] + h[4] * x[i][2] + boxlo[0];
    x[i][1] = h[1] * x[i][1] + h[3] * x[i][2] + boxlo[1];
    x[i][2] = h[2] * x[i][2] + boxlo[2];
  }
}
void Domain::x2lamda(const double *x, double *lamda) const
{
  double delta[3];
  delta[0] = x[0] - boxlo[0];
  delta[1] = x[1] - boxlo[1];
  delta[2] = x[2] - boxlo[2];
  lamda[0] = h_inv[0] * delta[0] + h_inv[5] * delta[1] + h_inv[4] * delta[2];
  lamda[1] = h_inv[1] * delta[1] + h_inv[3] * delta[2];
  lamda[2] = h_inv[2] * delta[2];
}
void Domain::lamda2x(const double *lamda, double *x) const
{
  x[0] = h[0] * lamda[0] + h[5] * lamda[1] + h[4] * lamda[2] + boxlo[0];
  x[1] = h[1] * lamda[1] + h[3] * lamda[2] + boxlo[1];
  x[2] = h[2] * lamda[2] + boxlo[2];
}

// bounding box (in box coords) of a lamda-space brick
void Domain::bbox(const double *lo, const double *hi, double *bboxlo, double *bboxhi) const
{
  double xx[3], lam[3];
  bboxlo[0] = bboxlo[1] = bboxlo[2] = BIG;
  bboxhi[0] = bboxhi[1] = bboxhi[2] = -BIG;
  for (int c = 0; c < 8; c++) {
    lam[0] = (c & 1) ? hi[0] : lo[0];
    lam[1] = (c & 2) ? hi[1] : lo[1];
    lam[2] = (c & 4) ? hi[2] : lo[2];
    lamda2x(lam, xx);
    for (int d = 0; d < 3; d++) {
      bboxlo[d] = MIN(bboxlo[d], xx[d]);
      bboxhi[d] = MAX(bboxhi[d], xx[d]);
    }
  }
}

// enforce PBC on owned atoms (called with lamda coords if triclinic)
void Domain::pbc()
{
  double *lo, *hi, *period;
  int nlocal = atom->nlocal;
  double **x = atom->x;
  if (triclinic == 0) { lo = boxlo; hi = boxhi; period = prd; }
  else { lo = boxlo_lamda; hi = boxhi_lamda; period = prd_lamda; }
  for (int i = 0; i < nlocal; i++)
    for (int d = 0; d < 3; d++) {
      if (x[i][d] < lo[d]) x[i][d] += period[d];
      if (x[i][d] >= hi[d]) {
        x[i][d] -= period[d];
        x[i][d] = MAX(x[i][d], lo[d]);
      }
    }
}

// remap a point (box coords) into the periodic box
void Domain::remap(double *x) const
{
  double lamda[3];
  const double *lo, *hi, *period;
  double *coord;
  if (triclinic == 0) { lo = boxlo; hi = boxhi; period = prd; coord = x; }
  else { lo = boxlo_lamda; hi = boxhi_lamda; period = prd_lamda; x2lamda(x, lamda); coord = lamda; }
  for (int d = 0; d < 3; d++) {
    while (coord[d] < lo[d]) coord[d] += period[d];
    while (coord[d] >= hi[d]) coord[d] -= period[d];
    coord[d] = MAX(coord[d], lo[d]);
  }
  if (triclinic) lamda2x(coord, x);
}
