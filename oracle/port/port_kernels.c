/* oracle/port/port_kernels.c -- TEST INFRASTRUCTURE, see port_kernels.h.  C99, scalar, no dependencies. */
#include "port_kernels.h"

#include <math.h>
#include <string.h>

#define PI 3.14159265358979323846
#define TWOPI 6.28318530717958647692
#define NEIGHMASK 0x1FFFFFFF
#define TOL 1.0e-9

/* ------------------------------------------------------------------ tallies (LAMMPS-core pair.cpp) */
static void tally_pair(port_tally *t, double evdwl, double fpair, double dx, double dy, double dz)
{ /* Pair::ev_tally with newton_pair = 1 */
  if (t->eflag_global) t->eng_vdwl += evdwl;
  if (t->vflag_global) {
    t->virial[0] += dx * dx * fpair;
    t->virial[1] += dy * dy * fpair;
    t->virial[2] += dz * dz * fpair;
    t->virial[3] += dx * dy * fpair;
    t->virial[4] += dx * dz * fpair;
    t->virial[5] += dy * dz * fpair;
  }
}
static void tally_v2(port_tally *t, double fpair, const double *d)
{ /* Pair::v_tally2 */
  if (!t->vflag_global) return;
  t->virial[0] += d[0] * d[0] * fpair;
  t->virial[1] += d[1] * d[1] * fpair;
  t->virial[2] += d[2] * d[2] * fpair;
  t->virial[3] += d[0] * d[1] * fpair;
  t->virial[4] += d[0] * d[2] * fpair;
  t->virial[5] += d[1] * d[2] * fpair;
}
static void tally_v3(port_tally *t, const double *fa, const double *fb, const double *da, const double *db)
{ /* Pair::v_tally3 and the virial part of Pair::ev_tally3: v = da (x) fa + db (x) fb */
  if (!t->vflag_global) return;
  t->virial[0] += da[0] * fa[0] + db[0] * fb[0];
  t->virial[1] += da[1] * fa[1] + db[1] * fb[1];
  t->virial[2] += da[2] * fa[2] + db[2] * fb[2];
  t->virial[3] += da[0] * fa[1] + db[0] * fb[1];
  t->virial[4] += da[0] * fa[2] + db[0] * fb[2];
  t->virial[5] += da[1] * fa[2] + db[1] * fb[2];
}

void port_virial_fdotr(int nall, const double *x, const double *f, double virial[6])
{
  for (int i = 0; i < nall; i++) {
    virial[0] += f[3 * i] * x[3 * i];
    virial[1] += f[3 * i + 1] * x[3 * i + 1];
    virial[2] += f[3 * i + 2] * x[3 * i + 2];
    virial[3] += f[3 * i + 1] * x[3 * i];
    virial[4] += f[3 * i + 2] * x[3 * i];
    virial[5] += f[3 * i + 2] * x[3 * i + 1];
  }
}

/* ------------------------------------------------------------------ REBOMoS scalar functions */
static double powint(double x, int n)
{ /* MathSpecial::powint */
  double yy, ww;
  if (n == 0) return 1.0;
  if (x == 0.0) return 0.0;
  int nn = n > 0 ? n : -n;
  ww = x;
  for (yy = 1.0; nn != 0; nn >>= 1, ww *= ww)
    if (nn & 1) yy *= ww;
  return n > 0 ? yy : 1.0 / yy;
}

/* cutoff switch Sp (pair_rebomos.h:195-211) */
static double Sp(double X, double Xmin, double Xmax, double *dX)
{
  const double t = (X - Xmin) / (Xmax - Xmin);
  if (t <= 0.0) { *dX = 0.0; return 1.0; }
  if (t >= 1.0) { *dX = 0.0; return 0.0; }
  *dX = (-0.5 * PI * sin(t * PI)) / (Xmax - Xmin);
  return 0.5 * (1.0 + cos(t * PI));
}

/* degree-6 polynomial and its derivative, evaluated in the reference's interleaved Horner order */
static double poly6(const double *c, double x, double *d)
{
  double g = c[6] * x, dg = 6.0 * c[6] * x;
  for (int k = 5; k >= 2; k--) {
    g += c[k];
    dg += k * c[k];
    g *= x;
    dg *= x;
  }
  g += c[1];
  dg += c[1];
  g *= x;
  g += c[0];
  *d = dg;
  return g;
}

/* angular function G(cos) (pair_rebomos.h:68-167) */
static double gSpline(const port_rebomos_par *p, double costh, int ti, double *dgdc)
{
  if (costh >= -1.0 && costh < 0.5) return poly6(p->b[ti], costh, dgdc);
  if (costh >= 0.5 && costh <= 1.0) {
    double dgcos, dgamma;
    const double gcos = poly6(p->b[ti], costh, &dgcos);
    const double gamma = poly6(p->bg[ti], costh, &dgamma);
    const double tmp = TWOPI * (costh - 0.5);
    const double psi = 0.5 * (1 - cos(tmp));
    const double dpsi = PI * sin(tmp);
    *dgdc = dgcos + dpsi * (gamma - gcos) + psi * (dgamma - dgcos);
    return gcos + psi * (gamma - gcos);
  }
  *dgdc = 0.0;
  return 0.0;
}

/* coordination function P(N) (pair_rebomos.h:173-179) */
static double PijSpline(const port_rebomos_par *p, double NM, double NS, int ti, double *dp)
{
  const double N = NM + NS;
  const double *a = p->a[ti];
  *dp = -a[0] + a[1] * a[2] * exp(-a[2] * N);
  return -a[0] * (N - 1) - a[1] * exp(-a[2] * N) + a[3];
}

void port_rebomos_setup(port_rebomos_par *p, const double v[61])
{
  int k = 0;
  double *tabs[7] = {&p->rcmin[0][0], &p->rcmax[0][0], &p->Q[0][0], &p->alpha[0][0], &p->A[0][0], &p->BIJc[0][0], &p->Beta[0][0]};
  for (int t = 0; t < 7; t++) {
    double *a = tabs[t];
    a[0] = v[k];
    a[1] = a[2] = v[k + 1];
    a[3] = v[k + 2];
    k += 3;
  }
  for (int e = 0; e < 2; e++) {
    for (int o = 0; o < 7; o++) p->b[e][o] = v[k++];
    for (int o = 0; o < 7; o++) p->bg[e][o] = v[k++];
  }
  for (int e = 0; e < 2; e++)
    for (int o = 0; o < 4; o++) p->a[e][o] = v[k++];
  const double eps_mm = v[k], eps_ss = v[k + 1], sig_mm = v[k + 2], sig_ss = v[k + 3];
  p->sigma[0][0] = sig_mm;
  p->sigma[0][1] = p->sigma[1][0] = (sig_mm + sig_ss) / 2;
  p->sigma[1][1] = sig_ss;
  p->epsilon[0][0] = eps_mm;
  p->epsilon[0][1] = p->epsilon[1][0] = sqrt(eps_mm * eps_ss);
  p->epsilon[1][1] = eps_ss;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) {
      p->rcmaxsq[i][j] = p->rcmax[i][j] * p->rcmax[i][j];
      p->rcLJmin[i][j] = p->rcmin[i][j];
      p->rcLJmax[i][j] = 2.5 * p->sigma[i][j];
      p->lj1[i][j] = 48.0 * p->epsilon[i][j] * powint(p->sigma[i][j], 12);
      p->lj2[i][j] = 24.0 * p->epsilon[i][j] * powint(p->sigma[i][j], 6);
      p->lj3[i][j] = 4.0 * p->epsilon[i][j] * powint(p->sigma[i][j], 12);
      p->lj4[i][j] = 4.0 * p->epsilon[i][j] * powint(p->sigma[i][j], 6);
    }
}

int port_rebo_neigh(const port_rebomos_par *p, int nrows, const double *x, const int *elem, const int *numneigh,
                    int *const *firstneigh, int *rebo_num, long *rebo_first, int *store, long cap, double *nM,
                    double *nS)
{
  long used = 0;
  for (int i = 0; i < nrows; i++) {
    const int ti = elem[i];
    int n = 0;
    rebo_first[i] = used;
    nM[i] = nS[i] = 0.0;
    for (int jj = 0; jj < numneigh[i]; jj++) {
      const int j = firstneigh[i][jj] & NEIGHMASK;
      const int tj = elem[j];
      const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
      const double rsq = dx * dx + dy * dy + dz * dz;
      if (rsq < p->rcmaxsq[ti][tj]) {
        if (used + n >= cap) return -1;
        store[used + n++] = j;
        double dS;
        const double w = Sp(sqrt(rsq), p->rcmin[ti][tj], p->rcmax[ti][tj], &dS);
        if (tj == 0) nM[i] += w;
        else nS[i] += w;
      }
    }
    rebo_num[i] = n;
    used += n;
  }
  return 0;
}

/* half-pair selection shared by FREBO and FLJ (pair_rebomos.cpp:394-402, 498-506): 1 = skip */
static int skip_pair(int itag, int jtag, const double *xi, const double *xj)
{
  if (itag > jtag) return (itag + jtag) % 2 == 0;
  if (itag < jtag) return (itag + jtag) % 2 == 1;
  if (xj[2] < xi[2]) return 1;
  if (xj[2] == xi[2] && xj[1] < xi[1]) return 1;
  if (xj[2] == xi[2] && xj[1] == xi[1] && xj[0] < xi[0]) return 1;
  return 0;
}

/* one side of the bond order (pair_rebomos.cpp:606-725 for the i side, :731-843 for the j side).
 * a = the atom whose neighbors are summed, b = its partner, rab = x_a - x_b "as seen from i":
 * sgn = +1 for the i side (cos from rij.rik), -1 for the j side (cos from -rij.rjl). */
static double bond_side(const port_rebomos_par *p, int a, int b, const double *rij, double rijmag, double VA, int sgn,
                        const double *x, const int *elem, const int *rebo_num, const long *rebo_first,
                        const int *store, const double *nM, const double *nS, double dwij, double *f, port_tally *t)
{
  const int ta = elem[a];
  const int *nb = store + rebo_first[a];
  const int nn = rebo_num[a];
  double Etmp = 0.0, dgdc, dS;
  for (int q = 0; q < nn; q++) {
    const int k = nb[q];
    if (k == b) continue;
    const int tk = elem[k];
    const double rak[3] = {x[3 * a] - x[3 * k], x[3 * a + 1] - x[3 * k + 1], x[3 * a + 2] - x[3 * k + 2]};
    const double rakmag = sqrt(rak[0] * rak[0] + rak[1] * rak[1] + rak[2] * rak[2]);
    const double wak = Sp(rakmag, p->rcmin[ta][tk], p->rcmax[ta][tk], &dS);
    double c = sgn * (rij[0] * rak[0] + rij[1] * rak[1] + rij[2] * rak[2]) / (rijmag * rakmag);
    c = c < 1.0 ? c : 1.0;
    c = c > -1.0 ? c : -1.0;
    Etmp += wak * gSpline(p, c, ta, &dgdc);
  }
  double dp;
  const double P = PijSpline(p, nM[a], nS[a], ta, &dp);
  const double pab = 1.0 / sqrt(1.0 + Etmp + P);
  const double tmp = -0.5 * pab * pab * pab;

  /* i = first atom of the bond, j = second, as in the reference's force expressions */
  const int atomi = sgn > 0 ? a : b, atomj = sgn > 0 ? b : a;
  for (int q = 0; q < nn; q++) {
    const int k = nb[q];
    if (k == b) continue;
    const int tk = elem[k];
    const double rak[3] = {x[3 * a] - x[3 * k], x[3 * a + 1] - x[3 * k + 1], x[3 * a + 2] - x[3 * k + 2]};
    const double rakmag = sqrt(rak[0] * rak[0] + rak[1] * rak[1] + rak[2] * rak[2]);
    double dwak;
    const double wak = Sp(rakmag, p->rcmin[ta][tk], p->rcmax[ta][tk], &dwak);
    double c = sgn * (rij[0] * rak[0] + rij[1] * rak[1] + rij[2] * rak[2]) / (rijmag * rakmag);
    c = c < 1.0 ? c : 1.0;
    c = c > -1.0 ? c : -1.0;
    double dci[3], dcj[3], dck[3];
    const double rr = rijmag * rakmag, r2ij = rijmag * rijmag, r2ak = rakmag * rakmag;
    for (int d = 0; d < 3; d++) {
      if (sgn > 0) {    /* pair_rebomos.cpp:648-665 */
        dci[d] = ((rij[d] + rak[d]) / rr) - (c * ((rij[d] / r2ij) + (rak[d] / r2ak)));
        dck[d] = (-rij[d] / rr) + (c * (rak[d] / r2ak));
        dcj[d] = (-rak[d] / rr) + (c * (rij[d] / r2ij));
      } else {          /* pair_rebomos.cpp:771-782, a = j, k = l */
        dci[d] = (-rak[d] / rr) - (c * rij[d] / r2ij);
        dcj[d] = ((-rij[d] + rak[d]) / rr) + (c * ((rij[d] / r2ij) - (rak[d] / r2ak)));
        dck[d] = (rij[d] / rr) + (c * rak[d] / r2ak);
      }
    }
    const double g = gSpline(p, c, ta, &dgdc);
    double fi[3], fj[3], fk[3];
    double tmp2 = VA * 0.5 * (tmp * wak * dgdc);
    for (int d = 0; d < 3; d++) {
      fi[d] = -tmp2 * dci[d];
      fj[d] = -tmp2 * dcj[d];
      fk[d] = -tmp2 * dck[d];
    }
    /* coordination forces: dw (from g) and dw (from P(N)); both along r_ak on a and k */
    double *fa = sgn > 0 ? fi : fj;
    tmp2 = VA * 0.5 * (tmp * dwak * g) / rakmag;
    for (int d = 0; d < 3; d++) { fa[d] -= tmp2 * rak[d]; fk[d] += tmp2 * rak[d]; }
    tmp2 = VA * 0.5 * (tmp * dp * dwak) / rakmag;
    for (int d = 0; d < 3; d++) { fa[d] -= tmp2 * rak[d]; fk[d] += tmp2 * rak[d]; }
    for (int d = 0; d < 3; d++) {
      f[3 * atomi + d] += fi[d];
      f[3 * atomj + d] += fj[d];
      f[3 * k + d] += fk[d];
    }
    if (t->vflag_global) {
      if (sgn > 0) {    /* v_tally3(i,j,k, fj, fk, rji, rki)   :707-711 */
        const double rji[3] = {-rij[0], -rij[1], -rij[2]}, rki[3] = {-rak[0], -rak[1], -rak[2]};
        tally_v3(t, fj, fk, rji, rki);
      } else {          /* v_tally3(i,j,l, fi, fl, rij, rlj)   :826-829 */
        const double rlj[3] = {-rak[0], -rak[1], -rak[2]};
        tally_v3(t, fi, fk, rij, rlj);
      }
    }
  }
  /* P(N) term through w_ij itself: N includes the bond partner (:715-725, :833-843) */
  const double tmp2 = -VA * 0.5 * (tmp * dp * dwij) / rijmag;
  for (int d = 0; d < 3; d++) {
    f[3 * atomi + d] += rij[d] * tmp2;
    f[3 * atomj + d] -= rij[d] * tmp2;
  }
  tally_v2(t, tmp2, rij);
  return pab;
}

void port_frebo(const port_rebomos_par *p, int nlocal, const double *x, const int *elem, const int *tag,
                const int *rebo_num, const long *rebo_first, const int *store, const double *nM, const double *nS,
                double *f, port_tally *t)
{
  for (int i = 0; i < nlocal; i++) {
    const int ti = elem[i];
    for (int q = 0; q < rebo_num[i]; q++) {
      const int j = store[rebo_first[i] + q];
      if (skip_pair(tag[i], tag[j], x + 3 * i, x + 3 * j)) continue;
      const int tj = elem[j];
      const double del[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
      const double rsq = del[0] * del[0] + del[1] * del[1] + del[2] * del[2];
      const double rij = sqrt(rsq);
      double dwij;
      const double wij = Sp(rij, p->rcmin[ti][tj], p->rcmax[ti][tj], &dwij);
      if (wij <= TOL) continue;
      const double Qij = p->Q[ti][tj], Aij = p->A[ti][tj], al = p->alpha[ti][tj];
      const double VR = wij * (1.0 + (Qij / rij)) * Aij * exp(-al * rij);
      const double pre = wij * Aij * exp(-al * rij);
      double dVRdi = pre * ((-al) - (Qij / rsq) - (Qij * al / rij));
      dVRdi += VR / wij * dwij;
      const double VA = -wij * p->BIJc[ti][tj] * exp(-p->Beta[ti][tj] * rij);
      double dVA = -p->Beta[ti][tj] * VA;
      dVA += VA / wij * dwij;
      const double pij = bond_side(p, i, j, del, rij, VA, +1, x, elem, rebo_num, rebo_first, store, nM, nS, dwij, f, t);
      const double pji = bond_side(p, j, i, del, rij, VA, -1, x, elem, rebo_num, rebo_first, store, nM, nS, dwij, f, t);
      const double bij = 0.5 * (pij + pji);
      const double fpair = -(dVRdi + bij * dVA) / rij;
      for (int d = 0; d < 3; d++) {
        f[3 * i + d] += del[d] * fpair;
        f[3 * j + d] -= del[d] * fpair;
      }
      tally_pair(t, VR + bij * VA, fpair, del[0], del[1], del[2]);
    }
  }
}

void port_flj(const port_rebomos_par *p, int nlocal, const double *x, const int *elem, const int *tag,
              const int *numneigh, int *const *firstneigh, double *f, port_tally *t)
{
  for (int i = 0; i < nlocal; i++) {
    const int ti = elem[i];
    for (int jj = 0; jj < numneigh[i]; jj++) {
      const int j = firstneigh[i][jj] & NEIGHMASK;
      if (skip_pair(tag[i], tag[j], x + 3 * i, x + 3 * j)) continue;
      const int tj = elem[j];
      const double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
      const double rsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      const double rij = sqrt(rsq);
      const double sig = p->sigma[ti][tj], rmin = p->rcLJmin[ti][tj], rmax = p->rcLJmax[ti][tj];
      double VLJ = 0.0, dVLJ = 0.0;
      if (rij > rmax || rij < rmin) {
        /* outside the window */
      } else if (rij <= rmax && rij >= 0.95 * sig) {
        const double r2inv = 1.0 / rsq, r6inv = r2inv * r2inv * r2inv;
        VLJ = r6inv * (p->lj3[ti][tj] * r6inv - p->lj4[ti][tj]);
        dVLJ = -r6inv * (p->lj1[ti][tj] * r6inv - p->lj2[ti][tj]) / rij;
      } else if (rij < 0.95 * sig && rij >= rmin) {
        const double eps = p->epsilon[ti][tj];
        const double dr = 0.95 * sig - rmin;
        const double r6 = powint((sig / (0.95 * sig)), 6);
        const double vdw = 4 * eps * r6 * (r6 - 1.0);
        const double dvdw = (-4 * eps / (0.95 * sig)) * r6 * (12.0 * r6 - 6.0);
        const double c2 = ((3.0 / dr) * vdw - dvdw) / dr;
        const double c3 = (vdw / (dr * dr) - c2) / dr;
        const double drp = rij - rmin;
        VLJ = drp * drp * (drp * c3 + c2);
        dVLJ = drp * (3.0 * drp * c3 + 2.0 * c2);
      }
      const double fpair = -dVLJ / rij;
      for (int k = 0; k < 3; k++) {
        f[3 * i + k] += d[k] * fpair;
        f[3 * j + k] -= d[k] * fpair;
      }
      tally_pair(t, VLJ, fpair, d[0], d[1], d[2]);
    }
  }
}

/* ------------------------------------------------------------------ AEAM */
void port_aeam_interpolate(int n, double delta, const double *f, double *s)
{
#define S(m, k) s[(long) (m) *7 + (k)]
  for (int m = 1; m <= n; m++) S(m, 6) = f[m];
  S(1, 5) = S(2, 6) - S(1, 6);
  S(2, 5) = 0.5 * (S(3, 6) - S(1, 6));
  S(n - 1, 5) = 0.5 * (S(n, 6) - S(n - 2, 6));
  S(n, 5) = S(n, 6) - S(n - 1, 6);
  for (int m = 3; m <= n - 2; m++) S(m, 5) = ((S(m - 2, 6) - S(m + 2, 6)) + 8.0 * (S(m + 1, 6) - S(m - 1, 6))) / 12.0;
  for (int m = 1; m <= n - 1; m++) {
    S(m, 4) = 3.0 * (S(m + 1, 6) - S(m, 6)) - 2.0 * S(m, 5) - S(m + 1, 5);
    S(m, 3) = S(m, 5) + S(m + 1, 5) - 2.0 * (S(m + 1, 6) - S(m, 6));
  }
  S(n, 4) = 0.0;
  S(n, 3) = 0.0;
  for (int m = 1; m <= n; m++) {
    S(m, 2) = S(m, 5) / delta;
    S(m, 1) = 2.0 * S(m, 4) / delta;
    S(m, 0) = 3.0 * S(m, 3) / delta;
  }
#undef S
}

/* table position of a radius: p = r/dr + 1, m = min(int(p), n-1), p = min(p - m, 1)  (pair_aeam.cpp:195-201) */
static const double *rrow(const port_aeam_par *P, double **tab, int pt, double r, double *pp)
{
  const double rdr = 1 / P->dr[pt];
  double p = r * rdr + 1.0;
  int m = (int) p;
  if (m > P->nr[pt] - 1) m = P->nr[pt] - 1;
  p -= m;
  if (p > 1.0) p = 1.0;
  *pp = p;
  return tab[pt] + (long) m * 7;
}
static double sval(const double *c, double p) { return ((c[3] * p + c[4]) * p + c[5]) * p + c[6]; }
static double sder(const double *c, double p) { return (c[0] * p + c[1]) * p + c[2]; }

void port_aeam_density(const port_aeam_par *P, int nlocal, int nall, const double *x, const int *type,
                       const int *numneigh, int *const *firstneigh, double *rho, double *fp, port_tally *t)
{
  const double THIRD = 1.0 / 3.0;
  const int nel = P->nel, nn = P->nnonangular;
  for (int i = 0; i < nall; i++) rho[i] = 0.0;
  for (int i = 0; i < nlocal; i++) {
    const int it = type[i];
    const int *jl = firstneigh[i];
    for (int jj = 0; jj < numneigh[i]; jj++) {
      const int j = jl[jj] & NEIGHMASK;
      const int jt = type[j];
      const double d1[3] = {x[3 * j] - x[3 * i], x[3 * j + 1] - x[3 * i + 1], x[3 * j + 2] - x[3 * i + 2]};
      const double rsq1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2];
      const double r1 = sqrt(rsq1);
      const int pij = (it - 1) * nel + (jt - 1);
      const double cutdec = (it > nn && jt > nn) ? 1.5 : 0;
      if (r1 > P->cut[pij] - cutdec) continue;
      double p1;
      const double *c = rrow(P, P->rhor_spline, pij, r1, &p1);
      const double fij = sval(c, p1);
      if (it <= nn) {
        rho[i] += fij;
      } else {
        for (int kk = jj + 1; kk < numneigh[i]; kk++) {
          const int k = jl[kk] & NEIGHMASK;
          const int kt = type[k];
          const double d2[3] = {x[3 * k] - x[3 * i], x[3 * k + 1] - x[3 * i + 1], x[3 * k + 2] - x[3 * i + 2]};
          const int pik = (it - 1) * nel + (kt - 1);
          const double cd = (kt > nn) ? 1.5 : 0;
          const double rsq2 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
          const double r2 = sqrt(rsq2);
          if (r2 > P->cut[pik] - cd) continue;
          const double d3[3] = {x[3 * k] - x[3 * j], x[3 * k + 1] - x[3 * j + 1], x[3 * k + 2] - x[3 * j + 2]};
          const double rsq3 = d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2];
          double p2;
          const double *c2 = rrow(P, P->rhor_spline, pik, r2, &p2);
          const double fik = sval(c2, p2);
          const double cs = (rsq1 + rsq2 - rsq3) / (2 * r1 * r2);
          const double delcs = cs + THIRD;
          rho[i] += 2 * fij * fik * (delcs * delcs);
        }
      }
    }
  }
  /* embedding (pair_aeam.cpp:264-303) */
  for (int i = 0; i < nlocal; i++) {
    const int it = type[i];
    const double rdrho = 1 / P->drho[it - 1];
    const double ni = (it <= nn) ? 1 : 0.5;
    double p = pow(rho[i], ni) * rdrho + 1.0;
    int m = (int) p;
    if (m > P->nrho[it - 1] - 1) m = P->nrho[it - 1] - 1;
    if (m < 1) m = 1;
    p -= m;
    if (p > 1.0) p = 1.0;
    const double *c = P->frho_spline[it - 1] + (long) m * 7;
    fp[i] = sder(c, p);
    if (t->eflag_global) t->eng_vdwl += sval(c, p);
  }
}

void port_aeam_force(const port_aeam_par *P, int nlocal, const double *x, const int *type, const int *numneigh,
                     int *const *firstneigh, const double *rho, const double *fp, double *f, port_tally *t)
{
  const double THIRD = 1.0 / 3.0, minrho = 0.0000000000001;
  const int nel = P->nel, nn = P->nnonangular;
  for (int i = 0; i < nlocal; i++) {
    const int it = type[i];
    const double ni = (it <= nn) ? 1 : 0.5, deli = (it <= nn) ? 0 : 1, ci = (it <= nn) ? 0 : 2;
    const double Fptmp = (rho[i] > minrho) ? ni * pow(rho[i], (ni - 1)) : 0;
    const int *jl = firstneigh[i];
    for (int jj = 0; jj < numneigh[i]; jj++) {
      const int j = jl[jj] & NEIGHMASK;
      const int jt = type[j];
      const double d1[3] = {x[3 * j] - x[3 * i], x[3 * j + 1] - x[3 * i + 1], x[3 * j + 2] - x[3 * i + 2]};
      const double rsq1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2];
      const double r1 = sqrt(rsq1);
      const int pij = (it - 1) * nel + (jt - 1);
      if (r1 > P->cut[pij]) continue;
      double p1;
      const double *c = rrow(P, P->rhor_spline, pij, r1, &p1);
      const double fij = sval(c, p1), dfij = sder(c, p1);
      const double *cz = rrow(P, P->z2r_spline, pij, r1, &p1);
      const double phip = sder(cz, p1), phi = sval(cz, p1);
      const double recip = 1 / r1;
      const double Feam = -(1 - deli) * Fptmp * fp[i] * (dfij * recip);
      const double fpair = Feam + 0.5 * (-phip * recip);
      for (int d = 0; d < 3; d++) {
        f[3 * i + d] -= d1[d] * fpair;
        f[3 * j + d] += d1[d] * fpair;
      }
      if (t->eflag_global) t->eng_vdwl += 0.5 * phi;
      tally_pair(t, 0.0, fpair, d1[0], d1[1], d1[2]);
      if (it <= nn) continue;
      for (int kk = jj + 1; kk < numneigh[i]; kk++) {
        const int k = jl[kk] & NEIGHMASK;
        const int kt = type[k];
        const double d2[3] = {x[3 * k] - x[3 * i], x[3 * k + 1] - x[3 * i + 1], x[3 * k + 2] - x[3 * i + 2]};
        const int pik = (it - 1) * nel + (kt - 1);
        const double cd = (kt > nn) ? 1.5 : 0;
        const double rsq2 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
        const double r2 = sqrt(rsq2);
        if (r2 > P->cut[pik] - cd) continue;
        const double d3[3] = {x[3 * k] - x[3 * j], x[3 * k + 1] - x[3 * j + 1], x[3 * k + 2] - x[3 * j + 2]};
        const double rsq3 = d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2];
        const double r3 = sqrt(rsq3);
        double p2;
        const double *c2 = rrow(P, P->rhor_spline, pik, r2, &p2);
        const double fik = sval(c2, p2), dfik = sder(c2, p2);
        const double cs = (rsq1 + rsq2 - rsq3) / (2 * r1 * r2);
        const double dcosij = 1 / r2 - cs / r1, dcosik = 1 / r1 - cs / r2, dcosjk = -r3 / (r1 * r2);
        const double delcs = cs + THIRD, ftet = delcs * delcs, delcs2 = 2 * delcs;
        const double DFij = ci * (fik * dfij * ftet + fij * fik * delcs2 * dcosij);
        const double DFik = ci * (fij * dfik * ftet + fij * fik * delcs2 * dcosik);
        const double DFjk = ci * fij * fik * delcs2 * dcosjk;
        const double FFij = -Fptmp * fp[i] * DFij / r1, FFik = -Fptmp * fp[i] * DFik / r2, FFjk = -Fptmp * fp[i] * DFjk / r3;
        double fj[3], fk[3];
        for (int d = 0; d < 3; d++) {
          fj[d] = d1[d] * FFij - d3[d] * FFjk;
          fk[d] = d2[d] * FFik + d3[d] * FFjk;
          f[3 * i + d] -= fj[d] + fk[d];
          f[3 * j + d] += fj[d];
          f[3 * k + d] += fk[d];
        }
        tally_v3(t, fj, fk, d1, d2);    /* ev_tally3(i,j,k,0,0,fj,fk,delr1,delr2) */
      }
    }
  }
}
