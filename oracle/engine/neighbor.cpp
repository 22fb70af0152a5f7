// minilmp engine (test infrastructure): neighbor lists, restating LAMMPS-core
// NBinStandard (setup_bins, coord2bin, bin_atoms), NStencilFullBin3d /
// NStencilFullGhostBin3d, NPairFullBin / NPairFullBinGhost and
// Neighbor::decide/check_distance -- SURVEY.md A.3.  Row order is the
// contract the GPU neighbor build is tested against bit-for-bit.

#include "engine.h"

using namespace LAMMPS_NS;

static constexpr double SMALL = 1.0e-6;
static constexpr double CUT2BIN_RATIO = 100.0;
static constexpr double BIGD = 1.0e20;

Neighbor::Neighbor(LAMMPS *l) : Pointers(l)
{
  style = 1;    // BIN
  every = 1;
  delay = 0;
  dist_check = 1;
  ago = -1;
  pgsize = 100000;
  oneatom = 2000;
  skin = 2.0;
  cutneighmin = cutneighmax = 0.0;
  cutneighsq = cutneighghostsq = nullptr;
  ncalls = ndanger = 0;
  lastcall = -1;
  includegroup = 0;
  list = nullptr;
  request = nullptr;
  binhead = bins = atom2bin = nullptr;
  maxbin = maxatombin = 0;
  nstencil = maxstencil = 0;
  stencil = nullptr;
  stencilxyz = nullptr;
  xhold = nullptr;
  maxhold = 0;
  triggersq = 0.0;
  mbins = 0;
}

Neighbor::~Neighbor()
{
  memory->destroy(cutneighsq);
  memory->destroy(cutneighghostsq);
  delete list;
  delete request;
  memory->destroy(binhead);
  memory->destroy(bins);
  memory->destroy(atom2bin);
  memory->destroy(stencil);
  delete[] stencilxyz;
  memory->destroy(xhold);
}

NeighRequest *Neighbor::add_request(Pair *pair, int flags)
{
  delete request;
  request = new NeighRequest(lmp, (void *) pair, flags);
  return request;
}

void Neighbor::init()
{
  int n = atom->ntypes;
  triggersq = 0.25 * skin * skin;

  memory->destroy(cutneighsq);
  memory->destroy(cutneighghostsq);
  memory->create(cutneighsq, n + 1, n + 1, "neigh:cutneighsq");
  memory->create(cutneighghostsq, n + 1, n + 1, "neigh:cutneighghostsq");

  cutneighmin = BIGD;
  cutneighmax = 0.0;
  double cutoff, delta, cut;
  for (int i = 1; i <= n; i++)
    for (int j = 1; j <= n; j++) {
      if (force->pair) cutoff = sqrt(force->pair->cutsq[i][j]);
      else cutoff = 0.0;
      if (cutoff > 0.0) delta = skin;
      else delta = 0.0;
      cut = cutoff + delta;
      cutneighsq[i][j] = cut * cut;
      cutneighmin = MIN(cutneighmin, cut);
      cutneighmax = MAX(cutneighmax, cut);
      if (force->pair && force->pair->ghostneigh) {
        cut = force->pair->cutghost[i][j] + skin;
        cutneighghostsq[i][j] = cut * cut;
      } else
        cutneighghostsq[i][j] = cut * cut;
    }

  // (re)create the one perpetual list from the pair style's request.  A pair style that builds its own list on the
  // device makes none (like the GPU package's styles with neighbor builds on the device): bins, xhold and the rebuild
  // decision still run, no host list is built
  if (request == nullptr) {
    delete list;
    list = nullptr;
    ago = -1;
    return;
  }
  if (!request->full) error->all(FLERR, "minilmp only builds full neighbor lists");
  delete list;
  list = new NeighList(lmp);
  list->ghost = request->ghost;
  list->ipage = new MyPage<int>[1];
  list->ipage[0].init(oneatom, pgsize, 1);
  force->pair->init_list(0, list);
  ago = -1;
}

// NBinStandard::setup_bins + NStencil::create_setup/create
void Neighbor::setup_bins()
{
  double bbox[3], bsubboxlo[3], bsubboxhi[3];
  double *cutghost = comm->cutghost;

  if (domain->triclinic == 0) {
    for (int d = 0; d < 3; d++) {
      bboxlo[d] = domain->boxlo[d];
      bboxhi[d] = domain->boxhi[d];
      bsubboxlo[d] = domain->sublo[d] - cutghost[d];
      bsubboxhi[d] = domain->subhi[d] + cutghost[d];
    }
  } else {
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      bboxlo[d] = domain->boxlo_bound[d];
      bboxhi[d] = domain->boxhi_bound[d];
      lo[d] = domain->sublo_lamda[d] - cutghost[d];
      hi[d] = domain->subhi_lamda[d] + cutghost[d];
    }
    domain->bbox(lo, hi, bsubboxlo, bsubboxhi);
  }

  bbox[0] = bboxhi[0] - bboxlo[0];
  bbox[1] = bboxhi[1] - bboxlo[1];
  bbox[2] = bboxhi[2] - bboxlo[2];

  // optimal bin size is roughly 1/2 the cutoff
  double binsize_optimal = 0.5 * cutneighmax;
  if (binsize_optimal == 0.0) binsize_optimal = bbox[0];
  double binsizeinv = 1.0 / binsize_optimal;

  if (bbox[0] * binsizeinv > MAXSMALLINT || bbox[1] * binsizeinv > MAXSMALLINT ||
      bbox[2] * binsizeinv > MAXSMALLINT)
    error->all(FLERR, "Domain too large for neighbor bins");

  // create actual bins; always have one bin even if cutoff > bbox
  nbinx = static_cast<int>(bbox[0] * binsizeinv);
  nbiny = static_cast<int>(bbox[1] * binsizeinv);
  nbinz = static_cast<int>(bbox[2] * binsizeinv);
  if (nbinx == 0) nbinx = 1;
  if (nbiny == 0) nbiny = 1;
  if (nbinz == 0) nbinz = 1;

  // compute actual bin size for nbins to fit into box exactly
  binsizex = bbox[0] / nbinx;
  binsizey = bbox[1] / nbiny;
  binsizez = bbox[2] / nbinz;
  bininvx = 1.0 / binsizex;
  bininvy = 1.0 / binsizey;
  bininvz = 1.0 / binsizez;

  if (binsize_optimal * bininvx > CUT2BIN_RATIO || binsize_optimal * bininvy > CUT2BIN_RATIO ||
      binsize_optimal * bininvz > CUT2BIN_RATIO)
    error->all(FLERR, "Cannot use neighbor bins - box size << cutoff");

  // mbinlo/hi = lowest and highest global bins my ghost atoms could be in
  // static_cast(-1.5) = -1, so subract additional -1; add in SMALL for round-off safety
  int mbinxhi, mbinyhi, mbinzhi;
  double coord;

  coord = bsubboxlo[0] - SMALL * bbox[0];
  mbinxlo = static_cast<int>((coord - bboxlo[0]) * bininvx);
  if (coord < bboxlo[0]) mbinxlo = mbinxlo - 1;
  coord = bsubboxhi[0] + SMALL * bbox[0];
  mbinxhi = static_cast<int>((coord - bboxlo[0]) * bininvx);

  coord = bsubboxlo[1] - SMALL * bbox[1];
  mbinylo = static_cast<int>((coord - bboxlo[1]) * bininvy);
  if (coord < bboxlo[1]) mbinylo = mbinylo - 1;
  coord = bsubboxhi[1] + SMALL * bbox[1];
  mbinyhi = static_cast<int>((coord - bboxlo[1]) * bininvy);

  coord = bsubboxlo[2] - SMALL * bbox[2];
  mbinzlo = static_cast<int>((coord - bboxlo[2]) * bininvz);
  if (coord < bboxlo[2]) mbinzlo = mbinzlo - 1;
  coord = bsubboxhi[2] + SMALL * bbox[2];
  mbinzhi = static_cast<int>((coord - bboxlo[2]) * bininvz);

  // extend bins by 1 to ensure stencil extent is included
  mbinxlo = mbinxlo - 1;
  mbinxhi = mbinxhi + 1;
  mbinx = mbinxhi - mbinxlo + 1;
  mbinylo = mbinylo - 1;
  mbinyhi = mbinyhi + 1;
  mbiny = mbinyhi - mbinylo + 1;
  mbinzlo = mbinzlo - 1;
  mbinzhi = mbinzhi + 1;
  mbinz = mbinzhi - mbinzlo + 1;

  bigint bbin = ((bigint) mbinx) * ((bigint) mbiny) * ((bigint) mbinz) + 1;
  if (bbin > MAXSMALLINT) error->one(FLERR, "Too many neighbor bins");
  mbins = (int) bbin;

  if (mbins > maxbin) {
    maxbin = mbins;
    memory->destroy(binhead);
    memory->create(binhead, maxbin, "neigh:binhead");
  }

  create_stencil();
}

double Neighbor::bin_distance(int i, int j, int k) const
{
  double delx, dely, delz;
  if (i > 0) delx = (i - 1) * binsizex;
  else if (i == 0) delx = 0.0;
  else delx = (i + 1) * binsizex;
  if (j > 0) dely = (j - 1) * binsizey;
  else if (j == 0) dely = 0.0;
  else dely = (j + 1) * binsizey;
  if (k > 0) delz = (k - 1) * binsizez;
  else if (k == 0) delz = 0.0;
  else delz = (k + 1) * binsizez;
  return (delx * delx + dely * dely + delz * delz);
}

void Neighbor::create_stencil()
{
  // sx,sy,sz = max range of stencil in each dim (NStencil::create_setup)
  sx = static_cast<int>(cutneighmax * bininvx);
  if (sx * binsizex < cutneighmax) sx++;
  sy = static_cast<int>(cutneighmax * bininvy);
  if (sy * binsizey < cutneighmax) sy++;
  sz = static_cast<int>(cutneighmax * bininvz);
  if (sz * binsizez < cutneighmax) sz++;

  int smax = (2 * sx + 1) * (2 * sy + 1) * (2 * sz + 1);
  if (smax > maxstencil) {
    maxstencil = smax;
    memory->destroy(stencil);
    memory->create(stencil, maxstencil, "neighstencil:stencil");
    delete[] stencilxyz;
    stencilxyz = new int[maxstencil][3];
  }

  double cutneighmaxsq = cutneighmax * cutneighmax;
  nstencil = 0;
  for (int k = -sz; k <= sz; k++)
    for (int j = -sy; j <= sy; j++)
      for (int i = -sx; i <= sx; i++)
        if (bin_distance(i, j, k) < cutneighmaxsq) {
          stencilxyz[nstencil][0] = i;
          stencilxyz[nstencil][1] = j;
          stencilxyz[nstencil][2] = k;
          stencil[nstencil++] = k * mbiny * mbinx + j * mbinx + i;
        }
}

int Neighbor::coord2bin(const double *x, int &ix, int &iy, int &iz) const
{
  if (!std::isfinite(x[0]) || !std::isfinite(x[1]) || !std::isfinite(x[2]))
    lmp->error->one(FLERR, "Non-numeric positions - simulation unstable");

  if (x[0] >= bboxhi[0]) ix = static_cast<int>((x[0] - bboxhi[0]) * bininvx) + nbinx;
  else if (x[0] >= bboxlo[0]) {
    ix = static_cast<int>((x[0] - bboxlo[0]) * bininvx);
    ix = MIN(ix, nbinx - 1);
  } else
    ix = static_cast<int>((x[0] - bboxlo[0]) * bininvx) - 1;

  if (x[1] >= bboxhi[1]) iy = static_cast<int>((x[1] - bboxhi[1]) * bininvy) + nbiny;
  else if (x[1] >= bboxlo[1]) {
    iy = static_cast<int>((x[1] - bboxlo[1]) * bininvy);
    iy = MIN(iy, nbiny - 1);
  } else
    iy = static_cast<int>((x[1] - bboxlo[1]) * bininvy) - 1;

  if (x[2] >= bboxhi[2]) iz = static_cast<int>((x[2] - bboxhi[2]) * bininvz) + nbinz;
  else if (x[2] >= bboxlo[2]) {
    iz = static_cast<int>((x[2] - bboxlo[2]) * bininvz);
    iz = MIN(iz, nbinz - 1);
  } else
    iz = static_cast<int>((x[2] - bboxlo[2]) * bininvz) - 1;

  ix -= mbinxlo;
  iy -= mbinylo;
  iz -= mbinzlo;
  return iz * mbiny * mbinx + iy * mbinx + ix;
}

int Neighbor::coord2bin(const double *x) const
{
  int ix, iy, iz;
  return coord2bin(x, ix, iy, iz);
}

void Neighbor::bin_atoms()
{
  int i, ibin;
  int nall = atom->nlocal + atom->nghost;
  if (nall > maxatombin) {
    maxatombin = atom->nmax;
    memory->destroy(bins);
    memory->destroy(atom2bin);
    memory->create(bins, maxatombin, "neigh:bins");
    memory->create(atom2bin, maxatombin, "neigh:atom2bin");
  }
  for (i = 0; i < mbins; i++) binhead[i] = -1;

  // bin in reverse order so linked list will be in forward order
  // also puts ghost atoms at end of list
  double **x = atom->x;
  for (i = nall - 1; i >= 0; i--) {
    ibin = coord2bin(x[i]);
    if (ibin < 0 || ibin >= mbins) error->one(FLERR, "Atom {} outside neighbor bins", i);
    atom2bin[i] = ibin;
    bins[i] = binhead[ibin];
    binhead[ibin] = i;
  }
}

int Neighbor::decide()
{
  ago++;
  if (ago >= delay && ago % every == 0) {
    if (dist_check == 0) return 1;
    return check_distance();
  }
  return 0;
}

int Neighbor::check_distance()
{
  double delx, dely, delz, rsq;
  double **x = atom->x;
  int nlocal = atom->nlocal;
  int flag = 0;
  for (int i = 0; i < nlocal; i++) {
    delx = x[i][0] - xhold[i][0];
    dely = x[i][1] - xhold[i][1];
    delz = x[i][2] - xhold[i][2];
    rsq = delx * delx + dely * dely + delz * delz;
    if (rsq > triggersq) flag = 1;
  }
  double fl = flag;
  universe->allreduce_max(comm->me, &fl, 1);
  int flagall = (int) fl;
  if (flagall && ago == MAX(every, delay)) ndanger++;
  return flagall;
}

// Neighbor::build + NPairFullBin[Ghost]::build
void Neighbor::build(int)
{
  int i, j, k, n, itype, jtype, ibin;
  double xtmp, ytmp, ztmp, delx, dely, delz, rsq;
  int *neighptr;

  ago = 0;
  ncalls++;
  lastcall = update->ntimestep;

  int nlocal = atom->nlocal;
  int nall = nlocal + atom->nghost;

  // store current atom positions for the displacement check
  if (dist_check) {
    double **x = atom->x;
    if (atom->nmax > maxhold) {
      maxhold = atom->nmax;
      memory->destroy(xhold);
      memory->create(xhold, maxhold, 3, "neigh:xhold");
    }
    for (i = 0; i < nlocal; i++) {
      xhold[i][0] = x[i][0];
      xhold[i][1] = x[i][1];
      xhold[i][2] = x[i][2];
    }
  }

  if (!list) return;    // no request: nothing to build on the host

  bin_atoms();

  list->grow(nlocal, nall);
  double **x = atom->x;
  int *type = atom->type;
  int *ilist = list->ilist;
  int *numneigh = list->numneigh;
  int **firstneigh = list->firstneigh;
  MyPage<int> *ipage = list->ipage;
  int inum = 0;
  ipage->reset();

  int nrows = list->ghost ? nall : nlocal;
  for (i = 0; i < nrows; i++) {
    n = 0;
    neighptr = ipage->vget();
    itype = type[i];
    xtmp = x[i][0];
    ytmp = x[i][1];
    ztmp = x[i][2];

    // loop over all atoms in surrounding bins in stencil including self, skip i = j
    // when i is a ghost atom, must check if stencil bin is out of bounds
    if (i < nlocal) {
      ibin = atom2bin[i];
      for (k = 0; k < nstencil; k++) {
        for (j = binhead[ibin + stencil[k]]; j >= 0; j = bins[j]) {
          if (i == j) continue;
          jtype = type[j];
          delx = xtmp - x[j][0];
          dely = ytmp - x[j][1];
          delz = ztmp - x[j][2];
          rsq = delx * delx + dely * dely + delz * delz;
          if (rsq <= cutneighsq[itype][jtype]) neighptr[n++] = j;
        }
      }
    } else {
      int xbin, ybin, zbin, xbin2, ybin2, zbin2;
      ibin = coord2bin(x[i], xbin, ybin, zbin);
      for (k = 0; k < nstencil; k++) {
        xbin2 = xbin + stencilxyz[k][0];
        ybin2 = ybin + stencilxyz[k][1];
        zbin2 = zbin + stencilxyz[k][2];
        if (xbin2 < 0 || xbin2 >= mbinx || ybin2 < 0 || ybin2 >= mbiny || zbin2 < 0 ||
            zbin2 >= mbinz)
          continue;
        for (j = binhead[ibin + stencil[k]]; j >= 0; j = bins[j]) {
          if (i == j) continue;
          jtype = type[j];
          delx = xtmp - x[j][0];
          dely = ytmp - x[j][1];
          delz = ztmp - x[j][2];
          rsq = delx * delx + dely * dely + delz * delz;
          if (rsq <= cutneighghostsq[itype][jtype]) neighptr[n++] = j;
        }
      }
    }

    ilist[inum++] = i;
    firstneigh[i] = neighptr;
    numneigh[i] = n;
    ipage->vgot(n);
    if (ipage->status()) error->one(FLERR, "Neighbor list overflow, boost neigh_modify one");
  }

  list->inum = nlocal;
  list->gnum = inum - nlocal;
}

bigint Neighbor::memory_usage()
{
  bigint bytes = 0;
  if (list && list->ipage) bytes += (bigint) list->ipage[0].size();
  return bytes;
}
