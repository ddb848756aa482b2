/*
 * ed_oracle.c -- see ed_oracle.h.  TEST INFRASTRUCTURE ONLY (checker / CPU baseline).
 *
 * Loop-by-loop C restatement of the reference's NORMAL-mode (ed_total_ud=T, Nph=0)
 * Fortran include fragments.  Deliberately keeps the reference's algorithmic cost:
 * bdecomp per row, c/cdg sign loops, binary search per matrix element.
 */
#include "ed_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------ */
/* ED_AUX_FUNX.f90:334-357  c(pos,in,out,fsgn)                         */
int ora_c(int pos, int32_t in, int32_t *out, double *sgn) {
  if (!((in >> (pos - 1)) & 1)) return 0; /* "C error: C_i|...0_i...>" */
  double s = 1.0;
  for (int l = 1; l <= pos - 1; l++)
    if ((in >> (l - 1)) & 1) s = -s;
  *sgn = s;
  *out = in & ~((int32_t)1 << (pos - 1));
  return 1;
}

/* ED_AUX_FUNX.f90:360-384  cdg(pos,in,out,fsgn) */
int ora_cdg(int pos, int32_t in, int32_t *out, double *sgn) {
  if ((in >> (pos - 1)) & 1) return 0; /* "C^+ error" */
  double s = 1.0;
  for (int l = 1; l <= pos - 1; l++)
    if ((in >> (l - 1)) & 1) s = -s;
  *sgn = s;
  *out = in | ((int32_t)1 << (pos - 1));
  return 1;
}

/* ED_AUX_FUNX.f90:399-407  bdecomp(i,Ntot) */
static inline void bdecomp(int32_t i, int ntot, int *ivec) {
  for (int l = 0; l < ntot; l++) ivec[l] = (i >> l) & 1;
}

/* ED_AUX_FUNX.f90:463-480  recursive binary_search(a,value): 1-based index, 0 if absent.
 * mid = size/2+1; go left on a(mid)>value, right (offset mid) on a(mid)<value. */
int64_t ora_binary_search(const int32_t *a, int64_t n, int32_t value) {
  int64_t base = 0;
  while (n > 0) {
    int64_t mid = n / 2 + 1; /* 1-based inside current slice */
    int32_t am = a[base + mid - 1];
    if (am > value) {
      n = mid - 1;
    } else if (am < value) {
      base += mid;
      n = n - mid;
    } else {
      return base + mid;
    }
  }
  return 0;
}

/* ED_SECTOR.f90:1925  binomial */
int64_t ora_binomial(int n, int k) {
  if (k < 0 || k > n) return 0;
  if (k > n - k) k = n - k;
  int64_t r = 1;
  for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
  return r;
}

/* ED_SECTOR.f90:217-242: scan iup=0..2^Ns-1 ascending, keep popcnt==nel */
int64_t ora_build_map(int Ns, int nel, int32_t *map) {
  int64_t dim = 0;
  for (int64_t i = 0; i < ((int64_t)1 << Ns); i++) {
    if (__builtin_popcountll((unsigned long long)i) != nel) continue;
    if (map) map[dim] = (int32_t)i;
    dim++;
  }
  return dim;
}

/* ------------------------------------------------------------------ */
/* direct/HxV_local.f90:14-83 : diagonal energy of state (mup,mdw)     */
static double diag_energy(const ora_params *p, const int *nup, const int *ndw) {
  const int Norb = p->Norb, Nbath = p->Nbath;
  double htmp = 0.0;
  for (int a = 0; a < Norb; a++) {
    htmp += p->eloc[0][a][a] * nup[a];
    htmp += p->eloc[1][a][a] * ndw[a];
    htmp -= p->xmu * (nup[a] + ndw[a]);
  }
  {
    int any = 0;
    for (int a = 0; a < Norb; a++) any |= (p->spin_field_z[a] != 0.0);
    if (any)
      for (int a = 0; a < Norb; a++) htmp += p->spin_field_z[a] * (nup[a] - ndw[a]);
  }
  for (int a = 0; a < Norb; a++) htmp += p->Uloc[a] * nup[a] * ndw[a];
  if (Norb > 1) {
    for (int a = 0; a < Norb; a++)
      for (int b = a + 1; b < Norb; b++)
        htmp += p->Ust[a][b] * (nup[a] * ndw[b] + nup[b] * ndw[a]);
    for (int a = 0; a < Norb; a++)
      for (int b = a + 1; b < Norb; b++)
        htmp += (p->Ust[a][b] - p->Jh[a][b]) * (nup[a] * nup[b] + ndw[a] * ndw[b]);
  }
  if (p->hfmode) {
    for (int a = 0; a < Norb; a++)
      htmp = htmp - 0.5 * p->Uloc[a] * (nup[a] + ndw[a]) + 0.25 * p->Uloc[a];
    if (Norb > 1) {
      for (int a = 0; a < Norb; a++)
        for (int b = a + 1; b < Norb; b++) {
          htmp = htmp - 0.5 * p->Ust[a][b] * (nup[a] + ndw[a] + nup[b] + ndw[b]) + 0.5 * p->Ust[a][b];
          htmp = htmp - 0.5 * (p->Ust[a][b] - p->Jh[a][b]) * (nup[a] + ndw[a] + nup[b] + ndw[b]) +
                 0.5 * (p->Ust[a][b] - p->Jh[a][b]);
        }
    }
  }
  for (int a = 0; a < p->Nfoo; a++)
    for (int k = 0; k < Nbath; k++) {
      int ialfa = p->stride[a][k];
      htmp += p->bath_diag[0][a][k] * nup[ialfa - 1];
      htmp += p->bath_diag[1][a][k] * ndw[ialfa - 1];
    }
  return htmp;
}

/*
 * One-body moves of one spin species out of Fock state m (direct/HxV_up.f90:11-122 and
 * HxV_dw.f90 with spin index Nspin): for every allowed term calls
 *   emit(ctx, target_index_1based, htmp)
 * in the reference's loop order.  map/dim = that species' sector map.
 * direct = 1 follows the direct path's exc_field loop (jorb=iorb+1..Norb,
 * HxV_up.f90:103), direct = 0 the stored path's (jorb=1..Norb, stored/H_up.f90:88).
 */
typedef void (*emit_fn)(void *ctx, int64_t i, double h);

static void species_hops(const ora_params *p, int s, int32_t m, const int32_t *map, int64_t dim,
                         int direct, emit_fn emit, void *ctx) {
  const int Norb = p->Norb, Nbath = p->Nbath, Ns = p->Ns;
  int n[32];
  int32_t k1, k2;
  double sg1, sg2;
  bdecomp(m, Ns, n);
  /* H_imp off-diagonal */
  for (int io = 0; io < Norb; io++)
    for (int jo = 0; jo < Norb; jo++) {
      if (p->eloc[s][io][jo] != 0.0 && n[jo] == 1 && n[io] == 0) {
        ora_c(jo + 1, m, &k1, &sg1);
        ora_cdg(io + 1, k1, &k2, &sg2);
        int64_t i = ora_binary_search(map, dim, k2);
        emit(ctx, i, p->eloc[s][io][jo] * sg1 * sg2);
      }
    }
  /* H_bath inter-orbital (replica/general) */
  if (p->bath_type == ORA_BATH_REPLICA || p->bath_type == ORA_BATH_GENERAL) {
    for (int kp = 0; kp < Nbath; kp++)
      for (int io = 0; io < Norb; io++)
        for (int jo = 0; jo < Norb; jo++) {
          int ialfa = p->stride[io][kp], ibeta = p->stride[jo][kp];
          if (p->hbath[s][io][jo][kp] != 0.0 && n[ibeta - 1] == 1 && n[ialfa - 1] == 0) {
            ora_c(ibeta, m, &k1, &sg1);
            ora_cdg(ialfa, k1, &k2, &sg2);
            int64_t i = ora_binary_search(map, dim, k2);
            emit(ctx, i, p->hbath[s][io][jo][kp] * sg1 * sg2);
          }
        }
  }
  /* H_hyb imp <-> bath */
  for (int io = 0; io < Norb; io++)
    for (int kp = 0; kp < Nbath; kp++) {
      int ialfa = p->stride[io][kp];
      double vk = p->diag_hybr[s][io][kp];
      if (vk != 0.0 && n[io] == 1 && n[ialfa - 1] == 0) {
        ora_c(io + 1, m, &k1, &sg1);
        ora_cdg(ialfa, k1, &k2, &sg2);
        int64_t i = ora_binary_search(map, dim, k2);
        emit(ctx, i, vk * sg1 * sg2);
      }
      if (vk != 0.0 && n[io] == 0 && n[ialfa - 1] == 1) {
        ora_c(ialfa, m, &k1, &sg1);
        ora_cdg(io + 1, k1, &k2, &sg2);
        int64_t i = ora_binary_search(map, dim, k2);
        emit(ctx, i, vk * sg1 * sg2);
      }
    }
  /* exciton fields */
  if (p->exc_field[0] != 0.0 || p->exc_field[1] != 0.0 || p->exc_field[2] != 0.0 ||
      p->exc_field[3] != 0.0) {
    for (int io = 0; io < Norb; io++)
      for (int jo = (direct ? io + 1 : 0); jo < Norb; jo++) {
        if (n[jo] == 1 && n[io] == 0) {
          ora_c(jo + 1, m, &k1, &sg1);
          ora_cdg(io + 1, k1, &k2, &sg2);
          int64_t i = ora_binary_search(map, dim, k2);
          emit(ctx, i, p->exc_field[0] * sg1 * sg2);
          emit(ctx, i, (s == 0 ? 1.0 : -1.0) * p->exc_field[3] * sg1 * sg2);
        }
      }
  }
}

/*
 * Non-local S-E / P-H moves out of state (mup,mdw): direct/HxV_non_local.f90:16-72.
 * emit(ctx, iup, idw packed as iup + (idw-1)*DimUp, htmp).
 */
static int nonloc_condition(const ora_params *p) {
  if (p->Norb <= 1) return 0;
  for (int a = 0; a < p->Norb; a++)
    for (int b = 0; b < p->Norb; b++)
      if (p->Jx[a][b] != 0.0 || p->Jp[a][b] != 0.0) return 1;
  return 0;
}
static int any_mat(const double m[ORA_MAXORB][ORA_MAXORB], int n) {
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++)
      if (m[a][b] != 0.0) return 1;
  return 0;
}

static void nonlocal_moves(const ora_params *p, int32_t mup, int32_t mdw, const int32_t *mapu,
                           int64_t DimUp, const int32_t *mapd, int64_t DimDw, emit_fn emit,
                           void *ctx) {
  const int Norb = p->Norb, Ns = p->Ns;
  int nup[32], ndw[32];
  int32_t k1, k2, k3, k4;
  double sg1, sg2, sg3, sg4;
  bdecomp(mup, Ns, nup);
  bdecomp(mdw, Ns, ndw);
  if (Norb > 1 && any_mat(p->Jx, Norb)) {
    for (int io = 0; io < Norb; io++)
      for (int jo = 0; jo < Norb; jo++) {
        if (io != jo && nup[jo] == 1 && ndw[io] == 1 && ndw[jo] == 0 && nup[io] == 0) {
          ora_c(io + 1, mdw, &k1, &sg1);
          ora_cdg(jo + 1, k1, &k2, &sg2);
          int64_t idw = ora_binary_search(mapd, DimDw, k2);
          ora_c(jo + 1, mup, &k3, &sg3);
          ora_cdg(io + 1, k3, &k4, &sg4);
          int64_t iup = ora_binary_search(mapu, DimUp, k4);
          emit(ctx, iup + (idw - 1) * DimUp, p->Jx[io][jo] * sg1 * sg2 * sg3 * sg4);
        }
      }
  }
  if (Norb > 1 && any_mat(p->Jp, Norb)) {
    for (int io = 0; io < Norb; io++)
      for (int jo = 0; jo < Norb; jo++) {
        if (nup[jo] == 1 && ndw[jo] == 1 && ndw[io] == 0 && nup[io] == 0) {
          ora_c(jo + 1, mdw, &k1, &sg1);
          ora_cdg(io + 1, k1, &k2, &sg2);
          int64_t idw = ora_binary_search(mapd, DimDw, k2);
          ora_c(jo + 1, mup, &k3, &sg3);
          ora_cdg(io + 1, k3, &k4, &sg4);
          int64_t iup = ora_binary_search(mapu, DimUp, k4);
          emit(ctx, iup + (idw - 1) * DimUp, p->Jp[io][jo] * sg1 * sg2 * sg3 * sg4);
        }
      }
  }
}

/* ------------------------------------------------------------------ */
typedef struct {
  const double *vin;
  double *hv;
  int64_t off;    /* added to the emitted 1-based index */
  int64_t stride; /* multiplied into the emitted index (1 for contiguous) */
  int64_t self;   /* 0-based position of the row being processed */
} gather_ctx;

/* Hv(j) += h*vin(i) : gather form used by HxV_up and HxV_non_local */
static void emit_gather(void *c, int64_t i, double h) {
  gather_ctx *g = (gather_ctx *)c;
  g->hv[g->self] += h * g->vin[g->off + (i - 1) * g->stride];
}
/* Hv(i) += h*vin(j) : scatter form used by HxV_dw */
static void emit_scatter(void *c, int64_t i, double h) {
  gather_ctx *g = (gather_ctx *)c;
  g->hv[g->off + (i - 1) * g->stride] += h * g->vin[g->self];
}

/* electronic part of one phonon slice: HxV_local + HxV_up + HxV_dw + HxV_non_local */
static void direct_hxv_slice(const ora_params *p, const int32_t *mapu, int64_t DimUp,
                             const int32_t *mapd, int64_t DimDw, const double *v, double *Hv) {
  const int Ns = p->Ns;
  const int64_t Dim = DimUp * DimDw;
  int nu[32], nd[32];
  /* direct/HxV_local.f90 */
  for (int64_t i = 1; i <= Dim; i++) {
    int64_t iup = i % DimUp;
    if (iup == 0) iup = DimUp;          /* iup_index ED_SECTOR.f90:1705 */
    int64_t idw = (i - 1) / DimUp + 1;  /* idw_index :1713 */
    bdecomp(mapu[iup - 1], Ns, nu);
    bdecomp(mapd[idw - 1], Ns, nd);
    Hv[i - 1] += diag_energy(p, nu, nd) * v[i - 1];
  }
  /* direct/HxV_up.f90 : Hv(j) += h*vin(i), i = iup + (jdw-1)*DimUp */
  for (int64_t jdw = 1; jdw <= DimDw; jdw++)
    for (int64_t jup = 1; jup <= DimUp; jup++) {
      gather_ctx g = {v, Hv, (jdw - 1) * DimUp, 1, jup - 1 + (jdw - 1) * DimUp};
      species_hops(p, 0, mapu[jup - 1], mapu, DimUp, 1, emit_gather, &g);
    }
  /* direct/HxV_dw.f90 : Hv(i) += h*vin(j), i = jup + (idw-1)*DimUp */
  for (int64_t jup = 1; jup <= DimUp; jup++)
    for (int64_t jdw = 1; jdw <= DimDw; jdw++) {
      gather_ctx g = {v, Hv, jup - 1, DimUp, jup - 1 + (jdw - 1) * DimUp};
      species_hops(p, 1, mapd[jdw - 1], mapd, DimDw, 1, emit_scatter, &g);
    }
  /* direct/HxV_non_local.f90 */
  if (nonloc_condition(p)) {
    for (int64_t j = 1; j <= Dim; j++) {
      int64_t jup = j % DimUp;
      if (jup == 0) jup = DimUp;
      int64_t jdw = (j - 1) / DimUp + 1;
      gather_ctx g = {v, Hv, 0, 1, j - 1};
      nonlocal_moves(p, mapu[jup - 1], mapd[jdw - 1], mapu, DimUp, mapd, DimDw, emit_gather, &g);
    }
  }
}

/* ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:23-130 (serial, DimPh=1) */
int ora_direct_hxv(const ora_params *p, int nup_el, int ndw_el, const double *v, double *Hv) {
  return ora_direct_hxv_ext(p, nup_el, ndw_el, 0, NULL, NULL, v, Hv);
}

/* one fermionic operator of a sundry chain on (mup, mdw): spin 1 -> up integer, 2 -> dw */
static int sundry_op(int create, int orb, int spin, int32_t *mup, int32_t *mdw, double *sg) {
  int32_t *m = (spin == 1) ? mup : mdw, out;
  int ok = create ? ora_cdg(orb, *m, &out, sg) : ora_c(orb, *m, &out, sg);
  if (ok) *m = out;
  return ok;
}

int ora_direct_hxv_ext(const ora_params *p, int nup_el, int ndw_el, int nsundry,
                       const ora_sundry_term *terms, const ora_phonons *ph, const double *v,
                       double *Hv) {
  const int Ns = p->Ns, Norb = p->Norb;
  const int64_t DimUp = ora_binomial(Ns, nup_el), DimDw = ora_binomial(Ns, ndw_el);
  const int64_t DimEl = DimUp * DimDw;
  const int DimPh = (ph ? ph->Nph : 0) + 1;
  const int64_t Dim = DimEl * DimPh;
  /* spin balance of every sundry term (HxV_sundry.f90:24-35) */
  for (int t = 0; t < nsundry; t++) {
    int sc = 0;
    sc += (terms[t].c_l[1] == 1) ? 1 : -1;
    sc -= (terms[t].cd_j[1] == 1) ? 1 : -1;
    sc += (terms[t].c_k[1] == 1) ? 1 : -1;
    sc -= (terms[t].cd_i[1] == 1) ? 1 : -1;
    if (sc != 0) return -2;
  }
  int32_t *mapu = (int32_t *)malloc(sizeof(int32_t) * DimUp);
  int32_t *mapd = (int32_t *)malloc(sizeof(int32_t) * DimDw);
  if (!mapu || !mapd) return -1;
  ora_build_map(Ns, nup_el, mapu);
  ora_build_map(Ns, ndw_el, mapd);
  memset(Hv, 0, sizeof(double) * Dim); /* Hv=zero (:98) */
  /* HxV_local / HxV_up / HxV_dw act on every phonon slice (their loops run over the whole
   * vector with j_el = mod(j-1,DimUp*DimDw)+1, e.g. HxV_local.f90:2-4) */
  for (int iph = 0; iph < DimPh; iph++)
    direct_hxv_slice(p, mapu, DimUp, mapd, DimDw, v + iph * DimEl, Hv + iph * DimEl);
  if (DimPh > 1) {
    /* direct/HxV_ph.f90:1-6 (+ A_ph(b+b^dag) of stored/H_ph.f90:6-17) */
    for (int64_t i = 1; i <= Dim; i++) {
      int64_t iph = (i - 1) / DimEl + 1;
      Hv[i - 1] += ph->w0 * (double)(iph - 1) * v[i - 1];
      if (ph->A != 0.0) {
        if (iph < DimPh) Hv[i - 1 + DimEl] += ph->A * sqrt((double)iph) * v[i - 1];
        if (iph > 1) Hv[i - 1 - DimEl] += ph->A * sqrt((double)(iph - 1)) * v[i - 1];
      }
    }
    /* direct/HxV_eph.f90:1-81, scatter form Hv(j) += h*vin(i) */
    int nu[32], nd[32];
    for (int64_t i = 1; i <= Dim; i++) {
      int64_t i_el = (i - 1) % DimEl + 1, iph = (i - 1) / DimEl + 1;
      int64_t iup = i_el % DimUp;
      if (iup == 0) iup = DimUp;
      int64_t idw = (i_el - 1) / DimUp + 1;
      int32_t mup = mapu[iup - 1], mdw = mapd[idw - 1], k1, k2;
      double sg1, sg2;
      bdecomp(mup, Ns, nu);
      bdecomp(mdw, Ns, nd);
      double htmp = 0.0;
      for (int a = 0; a < Norb; a++) htmp += ph->g[a][a] * (nu[a] + nd[a]);
      if (iph < DimPh) Hv[i_el - 1 + iph * DimEl] += htmp * sqrt((double)iph) * v[i - 1];
      if (iph > 1) Hv[i_el - 1 + (iph - 2) * DimEl] += htmp * sqrt((double)(iph - 1)) * v[i - 1];
      for (int s = 0; s < 2; s++) {
        const int *n = s == 0 ? nu : nd;
        for (int io = 0; io < Norb; io++)
          for (int jo = 0; jo < Norb; jo++) {
            if (!(ph->g[io][jo] != 0.0 && n[jo] == 1 && n[io] == 0)) continue;
            ora_c(jo + 1, s == 0 ? mup : mdw, &k1, &sg1);
            ora_cdg(io + 1, k1, &k2, &sg2);
            int64_t jup = iup, jdw = idw;
            if (s == 0)
              jup = ora_binary_search(mapu, DimUp, k2);
            else
              jdw = ora_binary_search(mapd, DimDw, k2);
            double h = ph->g[io][jo] * sg1 * sg2;
            int64_t j_el = jup + (jdw - 1) * DimUp;
            if (iph < DimPh) Hv[j_el - 1 + iph * DimEl] += h * sqrt((double)iph) * v[i - 1];
            if (iph > 1) Hv[j_el - 1 + (iph - 2) * DimEl] += h * sqrt((double)(iph - 1)) * v[i - 1];
          }
      }
    }
  }
  /* direct/HxV_sundry.f90:1-109, gather form Hv(j) += U*sg*vin(i), operators applied right to
   * left as c_l, cd_j, c_k, cd_i on the row state j */
  if (nsundry > 0) {
    for (int64_t j = 1; j <= Dim; j++) {
      int64_t j_el = (j - 1) % DimEl + 1, iph = (j - 1) / DimEl + 1;
      int64_t jup = j_el % DimUp;
      if (jup == 0) jup = DimUp;
      int64_t jdw = (j_el - 1) / DimUp + 1;
      for (int t = 0; t < nsundry; t++) {
        const ora_sundry_term *T = &terms[t];
        int32_t mu = mapu[jup - 1], md = mapd[jdw - 1];
        double s1, s2, s3, s4;
        if (!sundry_op(0, T->c_l[0], T->c_l[1], &mu, &md, &s1)) continue;
        if (!sundry_op(1, T->cd_j[0], T->cd_j[1], &mu, &md, &s2)) continue;
        if (!sundry_op(0, T->c_k[0], T->c_k[1], &mu, &md, &s3)) continue;
        if (!sundry_op(1, T->cd_i[0], T->cd_i[1], &mu, &md, &s4)) continue;
        int64_t idw = ora_binary_search(mapd, DimDw, md), iup = ora_binary_search(mapu, DimUp, mu);
        if (idw == 0 || iup == 0) { /* "H_sundry: impossible operator" */
          free(mapu);
          free(mapd);
          return -3;
        }
        int64_t i = iup + (idw - 1) * DimUp + (iph - 1) * DimEl;
        Hv[j - 1] += T->U * s1 * s2 * s3 * s4 * v[i - 1];
      }
    }
  }
  free(mapu);
  free(mapd);
  return 0;
}

/* ------------------------------------------------------------------ */
/* dw split of ED_HAMILTONIAN_NORMAL.f90:128-142 and row split of       */
/* vector_transpose_MPI (..._COMMON.f90:104-112): first (n mod P) ranks */
/* get one extra element.                                               */
static void block_split(int64_t n, int P, int r, int64_t *q, int64_t *start) {
  int64_t base = n / P, rem = n % P;
  *q = base + (r < rem ? 1 : 0);
  *start = r * base + (r < rem ? r : rem);
}

/*
 * ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:236-375 with P ranks emulated in one address
 * space.  Rank r owns dw columns [d0,d0+Qdw) (contiguous slice of the global vector).
 * vector_transpose_MPI(DimUp,Qdw,v,DimDw,Qup,vt) gives rank r the block
 * vt(idw, iup_loc) = v_global(iup0+iup_loc, idw); the all-to-all is realised as reads of
 * the other ranks' slices.  The three phases are separated by barriers exactly where the
 * reference has its collectives.
 */
static int direct_hxv_mpi_impl(const ora_params *p, int nup_el, int ndw_el, int P, int nrun,
                               int nthreads, const double *v, double *Hv, double *seconds) {
  const int Ns = p->Ns;
  const int64_t DimUp = ora_binomial(Ns, nup_el), DimDw = ora_binomial(Ns, ndw_el);
  if (P > DimDw) P = (int)DimDw; /* sub-communicator of DimDw ranks (:98-126) */
  if (P > DimUp) P = (int)DimUp; /* keep every rank a non-empty transposed block */
  int32_t *mapu = (int32_t *)malloc(sizeof(int32_t) * DimUp);
  int32_t *mapd = (int32_t *)malloc(sizeof(int32_t) * DimDw);
  ora_build_map(Ns, nup_el, mapu);
  ora_build_map(Ns, ndw_el, mapd);
  double **vt = (double **)calloc(P, sizeof(double *));
  double **hvt = (double **)calloc(P, sizeof(double *));
  const int nonloc = nonloc_condition(p);
  if (nthreads < 1) nthreads = 1;
  if (nrun < 1 || nrun > P) nrun = P; /* nrun < P: timing sample, only ranks [0,nrun) run */
  const double t_begin = now_s();
#pragma omp parallel num_threads(nthreads)
  {
    int nu[32], nd[32];
#pragma omp for schedule(static)
    for (int r = 0; r < nrun; r++) {
      int64_t Qdw, d0;
      block_split(DimDw, P, r, &Qdw, &d0);
      const int64_t ishift = d0 * DimUp; /* mpiIshift */
      const int64_t Nloc = DimUp * Qdw;
      const double *vin = v + ishift;
      double *hv = Hv + ishift;
      memset(hv, 0, sizeof(double) * Nloc);
      /* direct_mpi/HxV_local.f90 */
      for (int64_t i = 1; i <= Nloc; i++) {
        int64_t ig = i + ishift;
        int64_t iup = ig % DimUp;
        if (iup == 0) iup = DimUp;
        int64_t idw = (ig - 1) / DimUp + 1;
        bdecomp(mapu[iup - 1], Ns, nu);
        bdecomp(mapd[idw - 1], Ns, nd);
        hv[i - 1] += diag_energy(p, nu, nd) * vin[i - 1];
      }
      /* direct_mpi/HxV_up.f90 */
      for (int64_t jdw = 1; jdw <= Qdw; jdw++)
        for (int64_t jup = 1; jup <= DimUp; jup++) {
          gather_ctx g = {vin, hv, (jdw - 1) * DimUp, 1, jup - 1 + (jdw - 1) * DimUp};
          species_hops(p, 0, mapu[jup - 1], mapu, DimUp, 1, emit_gather, &g);
        }
      /* vector_transpose_MPI: vt(DimDw,Qup) */
      int64_t Qup, u0;
      block_split(DimUp, P, r, &Qup, &u0);
      vt[r] = (double *)malloc(sizeof(double) * (size_t)(Qup * DimDw));
      hvt[r] = (double *)calloc((size_t)(Qup * DimDw), sizeof(double));
      for (int64_t iu = 0; iu < Qup; iu++)
        for (int64_t id = 0; id < DimDw; id++) vt[r][id + iu * DimDw] = v[(u0 + iu) + id * DimUp];
      /* direct_mpi/HxV_dw.f90 on the transposed block: Hvt(i) += h*vt(j) */
      for (int64_t jdw = 1; jdw <= Qup; jdw++)
        for (int64_t jup = 1; jup <= DimDw; jup++) {
          gather_ctx g = {vt[r], hvt[r], (jdw - 1) * DimDw, 1, jup - 1 + (jdw - 1) * DimDw};
          species_hops(p, 1, mapd[jup - 1], mapd, DimDw, 1, emit_scatter, &g);
        }
    }
    /* implicit barrier = second vector_transpose_MPI; Hv += transpose(Hvt) (:341-342) */
#pragma omp for schedule(static)
    for (int r = 0; r < nrun; r++) {
      int64_t Qdw, d0;
      block_split(DimDw, P, r, &Qdw, &d0);
      for (int s = 0; s < P; s++) {
        int64_t Qup, u0;
        block_split(DimUp, P, s, &Qup, &u0);
        /* in a timing sample the blocks of ranks that did not run are stood in for by a
         * block that did (same size class), so the cost of this phase is still counted */
        const double *src = hvt[s < nrun ? s : s % nrun];
        int64_t Qs, us;
        block_split(DimUp, P, s < nrun ? s : s % nrun, &Qs, &us);
        if (Qs < Qup) Qup = Qs;
        for (int64_t id = d0; id < d0 + Qdw; id++)
          for (int64_t iu = 0; iu < Qup; iu++)
            Hv[(u0 + iu) + id * DimUp] += src[id + iu * DimDw];
      }
      /* allgather_vector_MPI + direct_mpi/HxV_non_local.f90 (:355-370) */
      if (nonloc) {
        const int64_t ishift = d0 * DimUp;
        for (int64_t j = 1; j <= DimUp * Qdw; j++) {
          int64_t jg = j + ishift;
          int64_t jup = jg % DimUp;
          if (jup == 0) jup = DimUp;
          int64_t jdw = (jg - 1) / DimUp + 1;
          gather_ctx g = {v, Hv + ishift, 0, 1, j - 1};
          nonlocal_moves(p, mapu[jup - 1], mapd[jdw - 1], mapu, DimUp, mapd, DimDw, emit_gather,
                         &g);
        }
      }
    }
  }
  if (seconds) *seconds = now_s() - t_begin;
  for (int r = 0; r < P; r++) {
    free(vt[r]);
    free(hvt[r]);
  }
  free(vt);
  free(hvt);
  free(mapu);
  free(mapd);
  return 0;
}

int ora_direct_hxv_mpi(const ora_params *p, int nup_el, int ndw_el, int P, int nthreads,
                       const double *v, double *Hv) {
  return direct_hxv_mpi_impl(p, nup_el, ndw_el, P, P, nthreads, v, Hv, NULL);
}

/* Timing sample of the same algorithm: P emulated ranks of which only the first nrun are
 * executed (one per thread); *seconds covers the product only (maps are built before, as
 * build_Hv_sector_normal does once per sector).  Hv is NOT a valid product when nrun < P. */
int ora_direct_hxv_mpi_sample(const ora_params *p, int nup_el, int ndw_el, int P, int nrun,
                              int nthreads, const double *v, double *Hv, double *seconds) {
  return direct_hxv_mpi_impl(p, nup_el, ndw_el, P, nrun, nthreads, v, Hv, seconds);
}

/* ------------------------------------------------------------------ */
/* stored path: list-of-rows with duplicate accumulation               */
/* (ED_SPARSE_MATRIX.f90:328-357) flattened to CSR in insertion order. */
typedef struct {
  int64_t row;    /* current target row filter (1-based) or 0 = count all */
  int64_t n;      /* entries in current row buffer */
  int64_t *cols;  /* row buffer */
  double *vals;
} rowbuf;

static void rowbuf_insert(rowbuf *b, int64_t col, double val) {
  for (int64_t k = 0; k < b->n; k++)
    if (b->cols[k] == col) {
      b->vals[k] += val;
      return;
    }
  b->cols[b->n] = col;
  b->vals[b->n] = val;
  b->n++;
}

/*
 * stored/H_up.f90 / H_dw.f90: loops over SOURCE states jup and inserts at
 * (row = target iup, col = jup).  To emit CSR rows without a dynamic structure we use the
 * fact that the one-body operator is structurally symmetric: the set of (target,source)
 * pairs of row i is produced by scanning all sources.  For the small Dim_sigma this O(nnz)
 * scan is done once per source and bucketed by target row (two passes: count, fill), which
 * preserves the reference's per-row insertion order (ascending source jup, then term order).
 */
typedef struct {
  int64_t src;
  int64_t *count;  /* per target row counts (pass 1) */
  int64_t *rowptr; /* pass 2 */
  int64_t *fill;
  int64_t *cols;
  double *vals;
  int pass;
} hop_ctx;

static void emit_hop(void *c, int64_t i, double h) {
  hop_ctx *x = (hop_ctx *)c;
  if (x->pass == 1) {
    x->count[i - 1]++;
  } else {
    int64_t pos = x->rowptr[i - 1] + x->fill[i - 1]++;
    x->cols[pos] = x->src;
    x->vals[pos] = h;
  }
}

int64_t ora_build_hop_csr(const ora_params *p, int spin, int nel, int64_t *rowptr, int32_t *cols,
                          double *vals) {
  const int Ns = p->Ns;
  const int64_t dim = ora_binomial(Ns, nel);
  int32_t *map = (int32_t *)malloc(sizeof(int32_t) * dim);
  ora_build_map(Ns, nel, map);
  int64_t *count = (int64_t *)calloc(dim + 1, sizeof(int64_t));
  hop_ctx x = {0, count, NULL, NULL, NULL, NULL, 1};
  for (int64_t j = 1; j <= dim; j++) {
    x.src = j;
    species_hops(p, spin, map[j - 1], map, dim, 0, emit_hop, &x);
  }
  int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * (dim + 1));
  rp[0] = 0;
  for (int64_t i = 0; i < dim; i++) rp[i + 1] = rp[i] + count[i];
  const int64_t nraw = rp[dim];
  int64_t *rcols = (int64_t *)malloc(sizeof(int64_t) * (nraw ? nraw : 1));
  double *rvals = (double *)malloc(sizeof(double) * (nraw ? nraw : 1));
  int64_t *fill = (int64_t *)calloc(dim + 1, sizeof(int64_t));
  x.pass = 2;
  x.rowptr = rp;
  x.fill = fill;
  x.cols = rcols;
  x.vals = rvals;
  for (int64_t j = 1; j <= dim; j++) {
    x.src = j;
    species_hops(p, spin, map[j - 1], map, dim, 0, emit_hop, &x);
  }
  /* accumulate duplicates per row in insertion order (sp_insert_element) */
  int64_t maxrow = 0;
  for (int64_t i = 0; i < dim; i++)
    if (count[i] > maxrow) maxrow = count[i];
  rowbuf b = {0, 0, (int64_t *)malloc(sizeof(int64_t) * (maxrow + 1)),
              (double *)malloc(sizeof(double) * (maxrow + 1))};
  int64_t nnz = 0;
  for (int64_t i = 0; i < dim; i++) {
    b.n = 0;
    for (int64_t k = rp[i]; k < rp[i + 1]; k++) rowbuf_insert(&b, rcols[k], rvals[k]);
    if (rowptr) rowptr[i] = nnz;
    if (cols && vals)
      for (int64_t k = 0; k < b.n; k++) {
        cols[nnz + k] = (int32_t)b.cols[k];
        vals[nnz + k] = b.vals[k];
      }
    nnz += b.n;
  }
  if (rowptr) rowptr[dim] = nnz;
  free(b.cols);
  free(b.vals);
  free(fill);
  free(rcols);
  free(rvals);
  free(rp);
  free(count);
  free(map);
  return nnz;
}

/* stored/H_local.f90: spH0d(i,i) */
int ora_build_diag(const ora_params *p, int nup_el, int ndw_el, double *diag) {
  const int Ns = p->Ns;
  const int64_t DimUp = ora_binomial(Ns, nup_el), DimDw = ora_binomial(Ns, ndw_el);
  int32_t *mapu = (int32_t *)malloc(sizeof(int32_t) * DimUp);
  int32_t *mapd = (int32_t *)malloc(sizeof(int32_t) * DimDw);
  ora_build_map(Ns, nup_el, mapu);
  ora_build_map(Ns, ndw_el, mapd);
  int nu[32], nd[32];
  for (int64_t i = 1; i <= DimUp * DimDw; i++) {
    int64_t iup = i % DimUp;
    if (iup == 0) iup = DimUp;
    int64_t idw = (i - 1) / DimUp + 1;
    bdecomp(mapu[iup - 1], Ns, nu);
    bdecomp(mapd[idw - 1], Ns, nd);
    diag[i - 1] = diag_energy(p, nu, nd);
  }
  free(mapu);
  free(mapd);
  return 0;
}

typedef struct {
  rowbuf *b;
} nl_ctx;
static void emit_nl(void *c, int64_t j, double h) { rowbuf_insert(((nl_ctx *)c)->b, j, h); }

/* stored/H_non_local.f90: row i = source state, col j = target (the transposed, symmetric
 * matrix), rows in insertion order with duplicate accumulation. */
int64_t ora_build_nonlocal_csr(const ora_params *p, int nup_el, int ndw_el, int64_t *rowptr,
                               int64_t *cols, double *vals) {
  const int Ns = p->Ns;
  const int64_t DimUp = ora_binomial(Ns, nup_el), DimDw = ora_binomial(Ns, ndw_el);
  int32_t *mapu = (int32_t *)malloc(sizeof(int32_t) * DimUp);
  int32_t *mapd = (int32_t *)malloc(sizeof(int32_t) * DimDw);
  ora_build_map(Ns, nup_el, mapu);
  ora_build_map(Ns, ndw_el, mapd);
  const int cap = 2 * ORA_MAXORB * ORA_MAXORB + 1;
  rowbuf b = {0, 0, (int64_t *)malloc(sizeof(int64_t) * cap), (double *)malloc(sizeof(double) * cap)};
  nl_ctx x = {&b};
  int64_t nnz = 0;
  const int on = nonloc_condition(p);
  for (int64_t i = 1; i <= DimUp * DimDw; i++) {
    int64_t iup = i % DimUp;
    if (iup == 0) iup = DimUp;
    int64_t idw = (i - 1) / DimUp + 1;
    b.n = 0;
    if (on) nonlocal_moves(p, mapu[iup - 1], mapd[idw - 1], mapu, DimUp, mapd, DimDw, emit_nl, &x);
    if (rowptr) rowptr[i - 1] = nnz;
    if (cols && vals)
      for (int64_t k = 0; k < b.n; k++) {
        cols[nnz + k] = b.cols[k];
        vals[nnz + k] = b.vals[k];
      }
    nnz += b.n;
  }
  if (rowptr) rowptr[DimUp * DimDw] = nnz;
  free(b.cols);
  free(b.vals);
  free(mapu);
  free(mapd);
  return nnz;
}

/* ED_HAMILTONIAN_NORMAL_STORED_HxV.f90:517-650 spMatVec_normal_main (DimPh=1) */
int ora_stored_hxv(const ora_params *p, int nup_el, int ndw_el, const double *v, double *Hv) {
  const int Ns = p->Ns;
  const int64_t DimUp = ora_binomial(Ns, nup_el), DimDw = ora_binomial(Ns, ndw_el);
  const int64_t Dim = DimUp * DimDw;
  double *diag = (double *)malloc(sizeof(double) * Dim);
  ora_build_diag(p, nup_el, ndw_el, diag);
  int64_t nzu = ora_build_hop_csr(p, 0, nup_el, NULL, NULL, NULL);
  int64_t nzd = ora_build_hop_csr(p, 1, ndw_el, NULL, NULL, NULL);
  int64_t *rpu = (int64_t *)malloc(sizeof(int64_t) * (DimUp + 1));
  int64_t *rpd = (int64_t *)malloc(sizeof(int64_t) * (DimDw + 1));
  int32_t *cu = (int32_t *)malloc(sizeof(int32_t) * (nzu + 1));
  int32_t *cd = (int32_t *)malloc(sizeof(int32_t) * (nzd + 1));
  double *vu = (double *)malloc(sizeof(double) * (nzu + 1));
  double *vd = (double *)malloc(sizeof(double) * (nzd + 1));
  ora_build_hop_csr(p, 0, nup_el, rpu, cu, vu);
  ora_build_hop_csr(p, 1, ndw_el, rpd, cd, vd);
  memset(Hv, 0, sizeof(double) * Dim);
  for (int64_t i = 0; i < Dim; i++) Hv[i] += diag[i] * v[i];
  /* DW (:548-566) */
  for (int64_t iup = 0; iup < DimUp; iup++)
    for (int64_t idw = 0; idw < DimDw; idw++) {
      int64_t i = iup + idw * DimUp;
      for (int64_t jj = rpd[idw]; jj < rpd[idw + 1]; jj++)
        Hv[i] += vd[jj] * v[iup + (int64_t)(cd[jj] - 1) * DimUp];
    }
  /* UP (:569-586) */
  for (int64_t idw = 0; idw < DimDw; idw++)
    for (int64_t iup = 0; iup < DimUp; iup++) {
      int64_t i = iup + idw * DimUp;
      for (int64_t jj = rpu[iup]; jj < rpu[iup + 1]; jj++)
        Hv[i] += vu[jj] * v[(cu[jj] - 1) + idw * DimUp];
    }
  /* Non-local (:629-645) */
  if (nonloc_condition(p)) {
    int64_t nz = ora_build_nonlocal_csr(p, nup_el, ndw_el, NULL, NULL, NULL);
    int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * (Dim + 1));
    int64_t *cc = (int64_t *)malloc(sizeof(int64_t) * (nz + 1));
    double *vv = (double *)malloc(sizeof(double) * (nz + 1));
    ora_build_nonlocal_csr(p, nup_el, ndw_el, rp, cc, vv);
    for (int64_t i = 0; i < Dim; i++)
      for (int64_t jj = rp[i]; jj < rp[i + 1]; jj++) Hv[i] += vv[jj] * v[cc[jj] - 1];
    free(rp);
    free(cc);
    free(vv);
  }
  free(diag);
  free(rpu);
  free(rpd);
  free(cu);
  free(cd);
  free(vu);
  free(vd);
  return 0;
}


/*
 * spMatVec_mpi_normal_main (..._STORED_HxV.f90:765-929) with P emulated ranks and the
 * diagonal kept as a stored vector: the reference's ED_SPARSE_H=T algorithm.  The build
 * happens once (as in build_Hv_sector_normal: stored_build), the product (stored_apply) is
 * repeated ncalls times and the mean product time returned in *seconds_per_call.
 */
typedef struct {
  int64_t DimUp, DimDw, Dim;
  int P, nthreads;
  int64_t *rpu, *rpd;
  int32_t *cu, *cd;
  double *vu, *vd, *diag;
  double **vt, **hvt;
  int64_t *rpn, *cn;
  double *vn;
} stored_sector;

static void stored_free(stored_sector *S) {
  free(S->rpn);
  free(S->cn);
  free(S->vn);
  if (S->vt)
    for (int r = 0; r < S->P; r++) {
      free(S->vt[r]);
      free(S->hvt[r]);
    }
  free(S->vt);
  free(S->hvt);
  free(S->diag);
  free(S->rpu);
  free(S->rpd);
  free(S->cu);
  free(S->cd);
  free(S->vu);
  free(S->vd);
  memset(S, 0, sizeof(*S));
}

static void stored_build(stored_sector *S, const ora_params *p, int nup_el, int ndw_el, int P,
                         int nthreads) {
  const int Ns = p->Ns;
  memset(S, 0, sizeof(*S));
  const int64_t DimUp = ora_binomial(Ns, nup_el), DimDw = ora_binomial(Ns, ndw_el);
  const int64_t Dim = DimUp * DimDw;
  if (P > DimDw) P = (int)DimDw;
  if (P > DimUp) P = (int)DimUp;
  if (nthreads < 1) nthreads = 1;
  S->DimUp = DimUp;
  S->DimDw = DimDw;
  S->Dim = Dim;
  S->P = P;
  S->nthreads = nthreads;
  int64_t nzu = ora_build_hop_csr(p, 0, nup_el, NULL, NULL, NULL);
  int64_t nzd = ora_build_hop_csr(p, 1, ndw_el, NULL, NULL, NULL);
  S->rpu = (int64_t *)malloc(sizeof(int64_t) * (DimUp + 1));
  S->rpd = (int64_t *)malloc(sizeof(int64_t) * (DimDw + 1));
  S->cu = (int32_t *)malloc(sizeof(int32_t) * (nzu + 1));
  S->cd = (int32_t *)malloc(sizeof(int32_t) * (nzd + 1));
  S->vu = (double *)malloc(sizeof(double) * (nzu + 1));
  S->vd = (double *)malloc(sizeof(double) * (nzd + 1));
  ora_build_hop_csr(p, 0, nup_el, S->rpu, S->cu, S->vu);
  ora_build_hop_csr(p, 1, ndw_el, S->rpd, S->cd, S->vd);
  /* spH0d over local rows: built in parallel like the reference (each rank its rows) */
  double *diag = S->diag = (double *)malloc(sizeof(double) * Dim);
  int32_t *mapu = (int32_t *)malloc(sizeof(int32_t) * DimUp);
  int32_t *mapd = (int32_t *)malloc(sizeof(int32_t) * DimDw);
  ora_build_map(Ns, nup_el, mapu);
  ora_build_map(Ns, ndw_el, mapd);
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t i = 0; i < Dim; i++) {
    int nu[32], nd[32];
    bdecomp(mapu[i % DimUp], Ns, nu);
    bdecomp(mapd[i / DimUp], Ns, nd);
    diag[i] = diag_energy(p, nu, nd);
  }
  free(mapu);
  free(mapd);
  S->vt = (double **)calloc(P, sizeof(double *));
  S->hvt = (double **)calloc(P, sizeof(double *));
  for (int r = 0; r < P; r++) {
    int64_t Qup, u0;
    block_split(DimUp, P, r, &Qup, &u0);
    S->vt[r] = (double *)malloc(sizeof(double) * (size_t)(Qup * DimDw));
    S->hvt[r] = (double *)malloc(sizeof(double) * (size_t)(Qup * DimDw));
  }
  /* spH0nd (stored/H_non_local.f90), applied to the all-gathered vector (:906-927) */
  if (nonloc_condition(p)) {
    int64_t nz = ora_build_nonlocal_csr(p, nup_el, ndw_el, NULL, NULL, NULL);
    S->rpn = (int64_t *)malloc(sizeof(int64_t) * (Dim + 1));
    S->cn = (int64_t *)malloc(sizeof(int64_t) * (nz + 1));
    S->vn = (double *)malloc(sizeof(double) * (nz + 1));
    ora_build_nonlocal_csr(p, nup_el, ndw_el, S->rpn, S->cn, S->vn);
  }
}

/* one product Hv = H v (:765-929): local part, transpose, dw part, transpose back, non-local */
static void stored_apply(const stored_sector *S, const double *v, double *Hv) {
  const int64_t DimUp = S->DimUp, DimDw = S->DimDw;
  const int P = S->P;
  const int64_t *rpu = S->rpu, *rpd = S->rpd, *rpn = S->rpn, *cn = S->cn;
  const int32_t *cu = S->cu, *cd = S->cd;
  const double *vu = S->vu, *vd = S->vd, *vn = S->vn, *diag = S->diag;
  double **vt = S->vt, **hvt = S->hvt;
#pragma omp parallel num_threads(S->nthreads)
  {
#pragma omp for schedule(static)
    for (int r = 0; r < P; r++) {
      int64_t Qdw, d0;
      block_split(DimDw, P, r, &Qdw, &d0);
      const int64_t ishift = d0 * DimUp, Nloc = DimUp * Qdw;
      const double *vin = v + ishift;
      double *hv = Hv + ishift;
      const double *dg = diag + ishift;
      for (int64_t i = 0; i < Nloc; i++) hv[i] = dg[i] * vin[i];
      for (int64_t idw = 0; idw < Qdw; idw++)
        for (int64_t iup = 0; iup < DimUp; iup++) {
          double acc = 0.0;
          for (int64_t jj = rpu[iup]; jj < rpu[iup + 1]; jj++)
            acc += vu[jj] * vin[(cu[jj] - 1) + idw * DimUp];
          hv[iup + idw * DimUp] += acc;
        }
      int64_t Qup, u0;
      block_split(DimUp, P, r, &Qup, &u0);
      for (int64_t iu = 0; iu < Qup; iu++)
        for (int64_t id = 0; id < DimDw; id++) vt[r][id + iu * DimDw] = v[(u0 + iu) + id * DimUp];
      for (int64_t iu = 0; iu < Qup; iu++)
        for (int64_t id = 0; id < DimDw; id++) {
          double acc = 0.0;
          for (int64_t jj = rpd[id]; jj < rpd[id + 1]; jj++)
            acc += vd[jj] * vt[r][(cd[jj] - 1) + iu * DimDw];
          hvt[r][id + iu * DimDw] = acc;
        }
    }
#pragma omp for schedule(static)
    for (int r = 0; r < P; r++) {
      int64_t Qdw, d0;
      block_split(DimDw, P, r, &Qdw, &d0);
      for (int s = 0; s < P; s++) {
        int64_t Qup, u0;
        block_split(DimUp, P, s, &Qup, &u0);
        for (int64_t id = d0; id < d0 + Qdw; id++)
          for (int64_t iu = 0; iu < Qup; iu++)
            Hv[(u0 + iu) + id * DimUp] += hvt[s][id + iu * DimDw];
      }
      if (rpn)
        for (int64_t i = d0 * DimUp; i < (d0 + Qdw) * DimUp; i++)
          for (int64_t jj = rpn[i]; jj < rpn[i + 1]; jj++) Hv[i] += vn[jj] * v[cn[jj] - 1];
    }
  }
}

int ora_stored_hxv_mpi(const ora_params *p, int nup_el, int ndw_el, int P, int nthreads, int ncalls,
                       const double *v, double *Hv, double *seconds_per_call) {
  stored_sector S;
  stored_build(&S, p, nup_el, ndw_el, P, nthreads);
  double t0 = now_s();
  for (int call = 0; call < ncalls; call++) stored_apply(&S, v, Hv);
  double t1 = now_s();
  if (seconds_per_call) *seconds_per_call = (t1 - t0) / (ncalls > 0 ? ncalls : 1);
  stored_free(&S);
  return 0;
}

/* ------------------------------------------------------------------ */
/* Symmetric tridiagonal eigenvalues by bisection on the Sturm sequence (the role of
 * SciFortran's eigh/tql2 on the Lanczos matrix inside sp_lanc_eigh): lowest eigenvalue only. */
static int sturm_count(int n, const double *a, const double *b, double x) {
  int cnt = 0;
  double q = a[0] - x;
  if (q < 0.0) cnt++;
  for (int i = 1; i < n; i++) {
    const double d = (fabs(q) < 1e-300) ? 1e-300 : q;
    q = a[i] - x - b[i - 1] * b[i - 1] / d;
    if (q < 0.0) cnt++;
  }
  return cnt;
}
static double tridiag_lowest(int n, const double *a, const double *b) {
  double lo = a[0], hi = a[0];
  for (int i = 0; i < n; i++) {
    const double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i + 1 < n ? fabs(b[i]) : 0.0);
    if (a[i] - r < lo) lo = a[i] - r;
    if (a[i] + r > hi) hi = a[i] + r;
  }
  for (int it = 0; it < 200 && hi - lo > 4e-16 * (fabs(lo) + fabs(hi)) + 1e-300; it++) {
    const double mid = 0.5 * (lo + hi);
    if (sturm_count(n, a, b, mid) >= 1) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

/*
 * Pass 1 of sp_lanc_eigh (SciFortran SF_SP_LINALG, call site ED_DIAG_NORMAL.f90:206-213) on the
 * stored (ED_SPARSE_H=T) operator above, all in C so that the BASELINE-size sectors finish in
 * tens of seconds: three-term recurrence of lanczos_iteration
 *   iter==1: vin/=|vin| ; else (vin,vout) <- (vout/beta, -beta*vin)
 *   vout += H vin ; alfa = <vin,vout> ; vout -= alfa*vin ; beta = |vout|
 * the lowest Ritz value is taken every iteration once nlanc >= ncheck; stop when it moved by
 * <= threshold, or |beta| < threshold, or nitermax.  v0: start vector (normalised inside).
 * alanc/blanc (size nitermax) receive the coefficients; *seconds the wall time of the loop.
 */
int ora_stored_lanczos_gs(const ora_params *p, int nup_el, int ndw_el, int P, int nthreads,
                          int nitermax, double threshold, int ncheck, const double *v0, double *egs,
                          int *niter, double *alanc, double *blanc, double *seconds) {
  stored_sector S;
  stored_build(&S, p, nup_el, ndw_el, P, nthreads);
  const int64_t n = S.Dim;
  if (nitermax > n) nitermax = (int)n;
  double *vin = (double *)malloc(sizeof(double) * n);
  double *vout = (double *)malloc(sizeof(double) * n);
  double *hv = (double *)malloc(sizeof(double) * n);
  double *bsub = (double *)calloc((size_t)nitermax + 1, sizeof(double));
  double nrm = 0.0;
#pragma omp parallel for reduction(+ : nrm) num_threads(S.nthreads)
  for (int64_t i = 0; i < n; i++) nrm += v0[i] * v0[i];
  nrm = sqrt(nrm);
#pragma omp parallel for num_threads(S.nthreads)
  for (int64_t i = 0; i < n; i++) {
    vin[i] = v0[i] / nrm;
    vout[i] = 0.0;
  }
  double beta = 0.0, elast = 0.0;
  int nlanc = 0, have_last = 0;
  const double t0 = now_s();
  for (int it = 1; it <= nitermax; it++) {
    if (it > 1) {
      const double ib = 1.0 / beta, mb = -beta;
#pragma omp parallel for num_threads(S.nthreads)
      for (int64_t i = 0; i < n; i++) {
        const double t = vin[i];
        vin[i] = vout[i] * ib;
        vout[i] = mb * t;
      }
    }
    stored_apply(&S, vin, hv);
    double alfa = 0.0;
#pragma omp parallel for reduction(+ : alfa) num_threads(S.nthreads)
    for (int64_t i = 0; i < n; i++) {
      vout[i] += hv[i];
      alfa += vin[i] * vout[i];
    }
    double b2 = 0.0;
#pragma omp parallel for reduction(+ : b2) num_threads(S.nthreads)
    for (int64_t i = 0; i < n; i++) {
      vout[i] -= alfa * vin[i];
      b2 += vout[i] * vout[i];
    }
    beta = sqrt(b2);
    alanc[it - 1] = alfa;
    nlanc = it;
    if (fabs(beta) < threshold && it > 1) break;
    bsub[it - 1] = beta;
    if (it < nitermax) blanc[it] = beta;
    if (nlanc >= ncheck) {
      const double e0 = tridiag_lowest(nlanc, alanc, bsub);
      if (have_last && fabs(e0 - elast) <= threshold) break;
      elast = e0;
      have_last = 1;
    }
  }
  if (seconds) *seconds = now_s() - t0;
  *egs = tridiag_lowest(nlanc, alanc, bsub);
  *niter = nlanc;
  free(vin);
  free(vout);
  free(hv);
  free(bsub);
  stored_free(&S);
  return 0;
}

/* ------------------------------------------------------------------ */
/* ED_SECTOR.f90:465-531 apply_op_C_d / :654 apply_op_CDG_d, ed_total_ud=T, DimPh=1:
 * OV(j) = sgn * V(i), sign from the operated spin's own integer only. */
int ora_apply_op(int Ns, int op, int iorb, int spin, int nup_el, int ndw_el, const double *v,
                 double *ov) {
  const int jnup = nup_el + (spin == 0 ? op : 0), jndw = ndw_el + (spin == 1 ? op : 0);
  if (jnup < 0 || jnup > Ns || jndw < 0 || jndw > Ns) return -1;
  const int64_t IUp = ora_binomial(Ns, nup_el), IDw = ora_binomial(Ns, ndw_el);
  const int64_t JUp = ora_binomial(Ns, jnup), JDw = ora_binomial(Ns, jndw);
  int32_t *imu = (int32_t *)malloc(sizeof(int32_t) * IUp), *imd = (int32_t *)malloc(sizeof(int32_t) * IDw);
  int32_t *jmu = (int32_t *)malloc(sizeof(int32_t) * JUp), *jmd = (int32_t *)malloc(sizeof(int32_t) * JDw);
  ora_build_map(Ns, nup_el, imu);
  ora_build_map(Ns, ndw_el, imd);
  ora_build_map(Ns, jnup, jmu);
  ora_build_map(Ns, jndw, jmd);
  memset(ov, 0, sizeof(double) * (size_t)(JUp * JDw));
  for (int64_t i = 1; i <= IUp * IDw; i++) {
    int64_t iu = (i - 1) % IUp + 1, id = (i - 1) / IUp + 1; /* state2indices :1691 */
    int32_t m = (spin == 0) ? imu[iu - 1] : imd[id - 1];
    int occ = (m >> iorb) & 1;
    int32_t r;
    double sgn;
    if (op < 0) {
      if (!occ) continue;
      ora_c(iorb + 1, m, &r, &sgn);
    } else {
      if (occ) continue;
      ora_cdg(iorb + 1, m, &r, &sgn);
    }
    int64_t ju = iu, jd = id;
    if (spin == 0)
      ju = ora_binary_search(jmu, JUp, r);
    else
      jd = ora_binary_search(jmd, JDw, r);
    ov[(ju - 1) + (jd - 1) * JUp] = sgn * v[i - 1];
  }
  free(imu);
  free(imd);
  free(jmu);
  free(jmd);
  return 0;
}
