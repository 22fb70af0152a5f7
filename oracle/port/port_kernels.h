/* oracle/port -- TEST INFRASTRUCTURE ("parity pinned": validated against the reference sources compiled
 * verbatim in oracle/_ref by tests/test_oracle_cpu.py, and against the goldens of log.rebomos-bulk.1/.4).
 *
 * Plain restatement of the arithmetic of the two reference pair styles as free functions over flat arrays
 * (x[n][3], 0-based element per atom, LAMMPS-style neighbor rows).  Each function cites the reference lines
 * it restates.  Never linked or called by the product (lammps_plugins_b200/).                              */
#ifndef B200MD_PORT_KERNELS_H
#define B200MD_PORT_KERNELS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double rcmin[2][2], rcmax[2][2], rcmaxsq[2][2];
  double Q[2][2], alpha[2][2], A[2][2], BIJc[2][2], Beta[2][2];
  double b[2][7], bg[2][7], a[2][4];
  double rcLJmin[2][2], rcLJmax[2][2], epsilon[2][2], sigma[2][2];
  double lj1[2][2], lj2[2][2], lj3[2][2], lj4[2][2];
} port_rebomos_par;

/* energy / virial accumulators with LAMMPS' flag semantics (Pair::ev_tally & friends) */
typedef struct {
  int eflag_global, vflag_global;
  double eng_vdwl;
  double virial[6];
} port_tally;

/* fills derived members (rcmaxsq, mixed LJ terms, lj1..lj4) from the 61 file values in file order
 * (pair_rebomos.cpp:884-948, 964-1066, 262-265) */
void port_rebomos_setup(port_rebomos_par *p, const double file_values[61]);

/* REBO_neigh (pair_rebomos.cpp:281-352).  rows of owned+ghost atoms; out rows packed into `store`
 * (capacity cap ints), rebo_first[i] = offset into store. returns 0, or -1 on overflow */
int port_rebo_neigh(const port_rebomos_par *p, int nrows, const double *x, const int *elem, const int *numneigh,
                    int *const *firstneigh, int *rebo_num, long *rebo_first, int *store, long cap, double *nM,
                    double *nS);

/* FREBO + bondorder (pair_rebomos.cpp:358-447, 571-847) over owned atoms */
void port_frebo(const port_rebomos_par *p, int nlocal, const double *x, const int *elem, const int *tag,
                const int *rebo_num, const long *rebo_first, const int *store, const double *nM, const double *nS,
                double *f, port_tally *t);

/* FLJ (pair_rebomos.cpp:453-558) over the full list of owned atoms */
void port_flj(const port_rebomos_par *p, int nlocal, const double *x, const int *elem, const int *tag,
              const int *numneigh, int *const *firstneigh, double *f, port_tally *t);

/* ---- AEAM ---- */
typedef struct {
  int nel, nnonangular;
  const int *nrho;         /* [nel] */
  const double *drho;      /* [nel] */
  const int *nr;           /* [nel*nel] */
  const double *dr, *cut;  /* [nel*nel] */
  double **frho_spline;    /* [nel] -> (nrho+1)*7 */
  double **rhor_spline;    /* [nel*nel] -> (nr+1)*7 */
  double **z2r_spline;     /* [nel*nel], symmetric pointers -> (nr+1)*7 */
} port_aeam_par;

/* interpolate (pair_aeam.cpp:915-942): f is 1-based [1..n], spline (n+1)*7 */
void port_aeam_interpolate(int n, double delta, const double *f, double *spline);

/* density pass + embedding (pair_aeam.cpp:158-303): rho[nlocal], fp[nlocal]; adds F(rho) to t->eng_vdwl */
void port_aeam_density(const port_aeam_par *p, int nlocal, int nall, const double *x, const int *type,
                       const int *numneigh, int *const *firstneigh, double *rho, double *fp, port_tally *t);

/* force pass (pair_aeam.cpp:309-476) */
void port_aeam_force(const port_aeam_par *p, int nlocal, const double *x, const int *type, const int *numneigh,
                     int *const *firstneigh, const double *rho, const double *fp, double *f, port_tally *t);

/* Pair::virial_fdotr_compute (LAMMPS-core): virial += sum over nall of x (x) f */
void port_virial_fdotr(int nall, const double *x, const double *f, double virial[6]);

#ifdef __cplusplus
}
#endif
#endif
