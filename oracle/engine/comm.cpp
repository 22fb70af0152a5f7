// minilmp engine (test infrastructure): brick domain decomposition and halo
// exchange, restating LAMMPS-core src/comm.cpp + src/comm_brick.cpp
// (setup, exchange, borders, forward_comm, reverse_comm and the Pair
// variants) -- SURVEY.md A.4.  Ranks are threads; messages go through
// Universe::sendrecv.

#include "engine.h"

using namespace LAMMPS_NS;

static constexpr double BIG = 1.0e20;
static constexpr int EXCHANGE_SIZE = 10;    // n, x[3], v[3], tag, type, mask
static constexpr int BORDER_SIZE = 6;       // x[3], tag, type, mask

Comm::Comm(LAMMPS *l) : Pointers(l)
{
  me = world ? world->rank : 0;
  nprocs = universe ? universe->nprocs : 1;
  nthreads = 1;
  for (int d = 0; d < 3; d++) {
    procgrid[d] = user_procgrid[d] = 0;
    myloc[d] = 0;
    procneigh[d][0] = procneigh[d][1] = me;
    cutghost[d] = 0.0;
    maxneed[d] = 0;
  }
  cutghostuser = 0.0;
  ghost_velocity = 0;
  maxexchange_atom = 0;
  nswap = maxswap = 0;
  sendnum = recvnum = sendproc = recvproc = firstrecv = pbc_flag = nullptr;
  pbc = nullptr;
  slablo = slabhi = nullptr;
  bytes_forward = bytes_reverse = 0;
}

Comm::~Comm()
{
  delete[] sendnum;
  delete[] recvnum;
  delete[] sendproc;
  delete[] recvproc;
  delete[] firstrecv;
  delete[] pbc_flag;
  delete[] pbc;
  delete[] slablo;
  delete[] slabhi;
}

// rank -> brick location: x fastest (LAMMPS' default "xyz" cart mapping w/o reordering
// gives ranks ordered with z fastest for MPI_Cart; here the grid is fixed by the user
// ("processors" command) and ranks are numbered x fastest, then y, then z)
void Comm::set_proc_grid()
{
  if (user_procgrid[0] * user_procgrid[1] * user_procgrid[2] != nprocs)
    error->all(FLERR, "Processor grid {}x{}x{} does not match {} ranks", user_procgrid[0],
               user_procgrid[1], user_procgrid[2], nprocs);
  for (int d = 0; d < 3; d++) procgrid[d] = user_procgrid[d];
  myloc[0] = me % procgrid[0];
  myloc[1] = (me / procgrid[0]) % procgrid[1];
  myloc[2] = me / (procgrid[0] * procgrid[1]);
  auto rankof = [&](int ix, int iy, int iz) {
    ix = (ix + procgrid[0]) % procgrid[0];
    iy = (iy + procgrid[1]) % procgrid[1];
    iz = (iz + procgrid[2]) % procgrid[2];
    return iz * procgrid[1] * procgrid[0] + iy * procgrid[0] + ix;
  };
  procneigh[0][0] = rankof(myloc[0] - 1, myloc[1], myloc[2]);
  procneigh[0][1] = rankof(myloc[0] + 1, myloc[1], myloc[2]);
  procneigh[1][0] = rankof(myloc[0], myloc[1] - 1, myloc[2]);
  procneigh[1][1] = rankof(myloc[0], myloc[1] + 1, myloc[2]);
  procneigh[2][0] = rankof(myloc[0], myloc[1], myloc[2] - 1);
  procneigh[2][1] = rankof(myloc[0], myloc[1], myloc[2] + 1);
}

void Comm::init() {}

double Comm::get_comm_cutoff() { return MAX(cutghostuser, neighbor->cutneighmax); }

// CommBrick::setup: slab boundaries, swap partners and PBC shifts for every swap
void Comm::setup()
{
  double *prd, *sublo, *subhi;
  double cut = get_comm_cutoff();
  int triclinic = domain->triclinic;

  if (triclinic == 0) {
    prd = domain->prd;
    sublo = domain->sublo;
    subhi = domain->subhi;
    cutghost[0] = cutghost[1] = cutghost[2] = cut;
  } else {
    prd = domain->prd_lamda;
    sublo = domain->sublo_lamda;
    subhi = domain->subhi_lamda;
    double *h_inv = domain->h_inv;
    double length0, length1, length2;
    length0 = sqrt(h_inv[0] * h_inv[0] + h_inv[5] * h_inv[5] + h_inv[4] * h_inv[4]);
    cutghost[0] = cut * length0;
    length1 = sqrt(h_inv[1] * h_inv[1] + h_inv[3] * h_inv[3]);
    cutghost[1] = cut * length1;
    length2 = h_inv[2];
    cutghost[2] = cut * length2;
  }

  // uniform layout, fully periodic
  for (int d = 0; d < 3; d++) maxneed[d] = static_cast<int>(cutghost[d] * procgrid[d] / prd[d]) + 1;

  nswap = 2 * (maxneed[0] + maxneed[1] + maxneed[2]);
  if (nswap > maxswap) {
    delete[] sendnum; delete[] recvnum; delete[] sendproc; delete[] recvproc;
    delete[] firstrecv; delete[] pbc_flag; delete[] pbc; delete[] slablo; delete[] slabhi;
    maxswap = nswap;
    sendnum = new int[maxswap];
    recvnum = new int[maxswap];
    sendproc = new int[maxswap];
    recvproc = new int[maxswap];
    firstrecv = new int[maxswap];
    pbc_flag = new int[maxswap];
    pbc = new int[maxswap][6];
    slablo = new double[maxswap];
    slabhi = new double[maxswap];
    sendlist.resize(maxswap);
  }

  int iswap = 0;
  for (int dim = 0; dim < 3; dim++) {
    for (int ineed = 0; ineed < 2 * maxneed[dim]; ineed++) {
      pbc_flag[iswap] = 0;
      for (int k = 0; k < 6; k++) pbc[iswap][k] = 0;

      if (ineed % 2 == 0) {
        sendproc[iswap] = procneigh[dim][0];
        recvproc[iswap] = procneigh[dim][1];
        if (ineed < 2) slablo[iswap] = -BIG;
        else slablo[iswap] = 0.5 * (sublo[dim] + subhi[dim]);
        slabhi[iswap] = sublo[dim] + cutghost[dim];
        if (myloc[dim] == 0) {
          pbc_flag[iswap] = 1;
          pbc[iswap][dim] = 1;
          if (triclinic) {
            if (dim == 1) pbc[iswap][5] = 1;
            else if (dim == 2) pbc[iswap][4] = pbc[iswap][3] = 1;
          }
        }
      } else {
        sendproc[iswap] = procneigh[dim][1];
        recvproc[iswap] = procneigh[dim][0];
        slablo[iswap] = subhi[dim] - cutghost[dim];
        if (ineed < 2) slabhi[iswap] = BIG;
        else slabhi[iswap] = 0.5 * (sublo[dim] + subhi[dim]);
        if (myloc[dim] == procgrid[dim] - 1) {
          pbc_flag[iswap] = 1;
          pbc[iswap][dim] = -1;
          if (triclinic) {
            if (dim == 1) pbc[iswap][5] = -1;
            else if (dim == 2) pbc[iswap][4] = pbc[iswap][3] = -1;
          }
        }
      }
      iswap++;
    }
  }
}

// CommBrick::exchange: migrate owned atoms that left my sub-box, dim by dim
void Comm::exchange()
{
  double *sublo, *subhi;
  double **x;
  atom->nghost = 0;

  if (domain->triclinic == 0) { sublo = domain->sublo; subhi = domain->subhi; }
  else { sublo = domain->sublo_lamda; subhi = domain->subhi_lamda; }

  for (int dim = 0; dim < 3; dim++) {
    x = atom->x;
    double lo = sublo[dim], hi = subhi[dim];
    int nlocal = atom->nlocal;
    int i = 0, nsend = 0;
    buf_send.clear();

    // fill buffer with atoms leaving my box, using < and >=
    // when atom is deleted, fill it in with last atom
    while (i < nlocal) {
      if (x[i][dim] < lo || x[i][dim] >= hi) {
        buf_send.push_back(EXCHANGE_SIZE);
        for (int d = 0; d < 3; d++) buf_send.push_back(atom->x[i][d]);
        for (int d = 0; d < 3; d++) buf_send.push_back(atom->v[i][d]);
        buf_send.push_back(atom->tag[i]);
        buf_send.push_back(atom->type[i]);
        buf_send.push_back(atom->mask[i]);
        nsend += EXCHANGE_SIZE;
        atom->copy(nlocal - 1, i);
        nlocal--;
      } else
        i++;
    }
    atom->nlocal = nlocal;

    // if 1 proc in dimension, no send/recv (leavers would be lost: pbc() ran before)
    int nrecv = 0;
    std::vector<double> recv_all;
    if (procgrid[dim] > 1) {
      int n1 = universe->sendrecv(me, procneigh[dim][1], buf_send.data(), nsend, buf_recv);
      recv_all.assign(buf_recv.begin(), buf_recv.begin() + n1);
      nrecv = n1;
      if (procgrid[dim] > 2) {
        int n2 = universe->sendrecv(me, procneigh[dim][0], buf_send.data(), nsend, buf_recv);
        recv_all.insert(recv_all.end(), buf_recv.begin(), buf_recv.begin() + n2);
        nrecv += n2;
      }
    }

    // check incoming atoms to see if they are in my box in this dim
    int m = 0;
    while (m < nrecv) {
      double value = recv_all[m + dim + 1];
      if (value >= lo && value < hi) {
        if (atom->nlocal == atom->nmax) atom->grow(0);
        int j = atom->nlocal;
        for (int d = 0; d < 3; d++) atom->x[j][d] = recv_all[m + 1 + d];
        for (int d = 0; d < 3; d++) atom->v[j][d] = recv_all[m + 4 + d];
        atom->tag[j] = (tagint) recv_all[m + 7];
        atom->type[j] = (int) recv_all[m + 8];
        atom->mask[j] = (int) recv_all[m + 9];
        atom->nlocal++;
      }
      m += static_cast<int>(recv_all[m]);
    }
  }
}

// CommBrick::borders: build ghost atoms + send lists (coords are lamda if triclinic)
void Comm::borders()
{
  int iswap = 0;
  int nfirst = 0, nlast = 0;
  int triclinic = domain->triclinic;

  for (int dim = 0; dim < 3; dim++) {
    nlast = 0;
    int twoneed = 2 * maxneed[dim];
    for (int ineed = 0; ineed < twoneed; ineed++) {
      double **x = atom->x;
      double lo = slablo[iswap], hi = slabhi[iswap];
      if (ineed % 2 == 0) {
        nfirst = nlast;
        nlast = atom->nlocal + atom->nghost;
      }

      // find atoms within slab boundaries lo/hi using <= and >=
      std::vector<int> &sl = sendlist[iswap];
      sl.clear();
      for (int i = nfirst; i < nlast; i++)
        if (x[i][dim] >= lo && x[i][dim] <= hi) sl.push_back(i);
      int nsend = (int) sl.size();

      // pack (AtomVec::pack_border)
      double dx = 0.0, dy = 0.0, dz = 0.0;
      if (pbc_flag[iswap]) {
        if (triclinic == 0) {
          dx = pbc[iswap][0] * domain->xprd;
          dy = pbc[iswap][1] * domain->yprd;
          dz = pbc[iswap][2] * domain->zprd;
        } else {
          dx = pbc[iswap][0];
          dy = pbc[iswap][1];
          dz = pbc[iswap][2];
        }
      }
      buf_send.resize((size_t) nsend * BORDER_SIZE + 1);
      int m = 0;
      for (int k = 0; k < nsend; k++) {
        int j = sl[k];
        if (pbc_flag[iswap]) {
          buf_send[m++] = x[j][0] + dx;
          buf_send[m++] = x[j][1] + dy;
          buf_send[m++] = x[j][2] + dz;
        } else {
          buf_send[m++] = x[j][0];
          buf_send[m++] = x[j][1];
          buf_send[m++] = x[j][2];
        }
        buf_send[m++] = atom->tag[j];
        buf_send[m++] = atom->type[j];
        buf_send[m++] = atom->mask[j];
      }

      // swap atoms with other proc, put incoming ghosts at end of my atom arrays
      int nr = universe->sendrecv(me, recvproc[iswap], buf_send.data(), m, buf_recv);
      int nrecv = nr / BORDER_SIZE;

      int first = atom->nlocal + atom->nghost;
      if (first + nrecv > atom->nmax) atom->grow(first + nrecv);
      m = 0;
      for (int k = first; k < first + nrecv; k++) {
        atom->x[k][0] = buf_recv[m++];
        atom->x[k][1] = buf_recv[m++];
        atom->x[k][2] = buf_recv[m++];
        atom->tag[k] = (tagint) buf_recv[m++];
        atom->type[k] = (int) buf_recv[m++];
        atom->mask[k] = (int) buf_recv[m++];
        atom->v[k][0] = atom->v[k][1] = atom->v[k][2] = 0.0;
      }

      sendnum[iswap] = nsend;
      recvnum[iswap] = nrecv;
      firstrecv[iswap] = first;
      atom->nghost += nrecv;
      iswap++;
    }
  }
}

// CommBrick::forward_comm: ghost positions from owners (box coords)
void Comm::forward_comm()
{
  double **x = atom->x;
  bytes_forward = 0;
  for (int iswap = 0; iswap < nswap; iswap++) {
    const std::vector<int> &sl = sendlist[iswap];
    int n = sendnum[iswap];
    double dx = 0.0, dy = 0.0, dz = 0.0;
    if (pbc_flag[iswap]) {
      const int *p = pbc[iswap];
      if (domain->triclinic == 0) {
        dx = p[0] * domain->xprd;
        dy = p[1] * domain->yprd;
        dz = p[2] * domain->zprd;
      } else {
        dx = p[0] * domain->xprd + p[5] * domain->xy + p[4] * domain->xz;
        dy = p[1] * domain->yprd + p[3] * domain->yz;
        dz = p[2] * domain->zprd;
      }
    }
    buf_send.resize((size_t) 3 * n + 1);
    int m = 0;
    if (pbc_flag[iswap] == 0) {
      for (int k = 0; k < n; k++) {
        int j = sl[k];
        buf_send[m++] = x[j][0];
        buf_send[m++] = x[j][1];
        buf_send[m++] = x[j][2];
      }
    } else {
      for (int k = 0; k < n; k++) {
        int j = sl[k];
        buf_send[m++] = x[j][0] + dx;
        buf_send[m++] = x[j][1] + dy;
        buf_send[m++] = x[j][2] + dz;
      }
    }
    universe->sendrecv(me, recvproc[iswap], buf_send.data(), m, buf_recv);
    int first = firstrecv[iswap];
    m = 0;
    for (int k = first; k < first + recvnum[iswap]; k++) {
      x[k][0] = buf_recv[m++];
      x[k][1] = buf_recv[m++];
      x[k][2] = buf_recv[m++];
    }
    bytes_forward += (bigint) sizeof(double) * 3 * n;
  }
}

// CommBrick::reverse_comm: ghost forces summed into owners, swaps in reverse order
void Comm::reverse_comm()
{
  double **f = atom->f;
  bytes_reverse = 0;
  for (int iswap = nswap - 1; iswap >= 0; iswap--) {
    int first = firstrecv[iswap];
    int nr = recvnum[iswap];
    buf_send.resize((size_t) 3 * nr + 1);
    int m = 0;
    for (int k = first; k < first + nr; k++) {
      buf_send[m++] = f[k][0];
      buf_send[m++] = f[k][1];
      buf_send[m++] = f[k][2];
    }
    // reverse direction: I send to the rank I received from
    universe->sendrecv(me, sendproc[iswap], buf_send.data(), m, buf_recv);
    const std::vector<int> &sl = sendlist[iswap];
    m = 0;
    for (int k = 0; k < sendnum[iswap]; k++) {
      int j = sl[k];
      f[j][0] += buf_recv[m++];
      f[j][1] += buf_recv[m++];
      f[j][2] += buf_recv[m++];
    }
    bytes_reverse += (bigint) sizeof(double) * 3 * nr;
  }
}

// CommBrick::forward_comm(Pair *): per-atom pair quantity, owners -> ghosts
void Comm::forward_comm(Pair *pair)
{
  int nsize = pair->comm_forward;
  for (int iswap = 0; iswap < nswap; iswap++) {
    int n = sendnum[iswap];
    buf_send.resize((size_t) nsize * n + 1);
    int m = pair->pack_forward_comm(n, sendlist[iswap].data(), buf_send.data(), pbc_flag[iswap],
                                    pbc[iswap]);
    universe->sendrecv(me, recvproc[iswap], buf_send.data(), m, buf_recv);
    pair->unpack_forward_comm(recvnum[iswap], firstrecv[iswap], buf_recv.data());
  }
}

// CommBrick::reverse_comm(Pair *): per-atom pair quantity, ghosts summed into owners
void Comm::reverse_comm(Pair *pair)
{
  int nsize = MAX(pair->comm_reverse, pair->comm_reverse_off);
  for (int iswap = nswap - 1; iswap >= 0; iswap--) {
    int nr = recvnum[iswap];
    buf_send.resize((size_t) nsize * nr + 1);
    int m = pair->pack_reverse_comm(nr, firstrecv[iswap], buf_send.data());
    universe->sendrecv(me, sendproc[iswap], buf_send.data(), m, buf_recv);
    pair->unpack_reverse_comm(sendnum[iswap], sendlist[iswap].data(), buf_recv.data());
  }
}
