/*
 * ed_oracle.h -- CPU restatement of EDIpack's NORMAL-mode Lanczos HxV hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call this code, and only as the checker / CPU baseline.
 *
 * The reference (Fortran 90 + MPI, /root/reference) cannot be compiled in this image
 * (no Fortran compiler, no MPI, no SciFortran), so this file restates its loops in C,
 * function by function, with 64-bit linear indices.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference/).
 *
 * Parity status: pinned END-TO-END against the reference's golden files
 * test/src/NORMAL_NORMAL/{evals,dens,docc,Sigma_momenta}.check (see
 * tests/test_oracle_golden.py).  The SciFortran Lanczos drivers are not in the
 * reference tree (unpinned dependency): at the sp_lanc_* boundary parity is UNPINNED
 * except through those end-to-end observables.
 */
#ifndef ED_ORACLE_H
#define ED_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAXORB 5
#define ORA_MAXBATH 32

/* bath_type codes (ED_INPUT_VARS.f90:598) */
enum { ORA_BATH_NORMAL = 0, ORA_BATH_HYBRID = 1, ORA_BATH_REPLICA = 2, ORA_BATH_GENERAL = 3 };

/*
 * Everything the NORMAL-mode H x v reads from module globals in the reference
 * (SURVEY 8b "Semantics"), flattened into one POD.  Spin index s: 0 = up, 1 = dw;
 * for Nspin = 1 the caller fills both spin slots with the same numbers, which is what
 * the reference's index "Nspin" (=1) does (direct/HxV_dw.f90:14, HxV_local.f90:18).
 * Orbital/bath indices are 0-based here, 1-based in the reference.
 */
typedef struct {
  int32_t Ns;        /* levels per spin (ED_SETUP.f90:118-126) */
  int32_t Norb;
  int32_t Nbath;
  int32_t bath_type; /* ORA_BATH_* */
  int32_t hfmode;    /* Hartree-shifted interaction (HxV_local.f90:58) */
  int32_t Nfoo;      /* size(bath_diag,2): Norb, or 1 for hybrid (DIRECT_HxV.f90:60) */
  int32_t pad0, pad1;
  double xmu;
  double eloc[2][ORA_MAXORB][ORA_MAXORB]; /* impHloc(s,s,a,b)+mfHloc(s,s,a,b) */
  double spin_field_z[ORA_MAXORB];        /* spin_field(a,3) */
  double exc_field[4];
  double Uloc[ORA_MAXORB];
  double Ust[ORA_MAXORB][ORA_MAXORB]; /* Ust_internal */
  double Jh[ORA_MAXORB][ORA_MAXORB];  /* Jh_internal  */
  double Jx[ORA_MAXORB][ORA_MAXORB];  /* Jx_internal  */
  double Jp[ORA_MAXORB][ORA_MAXORB];  /* Jp_internal  */
  double diag_hybr[2][ORA_MAXORB][ORA_MAXBATH];         /* diag_hybr(s,a,k) */
  double bath_diag[2][ORA_MAXORB][ORA_MAXBATH];         /* bath_diag(s,a|1,k) */
  double hbath[2][ORA_MAXORB][ORA_MAXORB][ORA_MAXBATH]; /* hbath_tmp(s,s,a,b,k), replica/general */
  int32_t stride[ORA_MAXORB][ORA_MAXBATH];              /* getBathStride(a,k), 1-based site */
} ora_params;

/* ---- ED_AUX_FUNX.f90 bit operators ---- */
int ora_c(int pos, int32_t in, int32_t *out, double *sgn);   /* :334 */
int ora_cdg(int pos, int32_t in, int32_t *out, double *sgn); /* :360 */
int64_t ora_binary_search(const int32_t *a, int64_t n, int32_t value); /* :463, 1-based, 0 = not found */
int64_t ora_binomial(int n, int k);                                    /* ED_SECTOR.f90:1925 */

/* ---- ED_SECTOR.f90:217-242 build_sector (normal, ed_total_ud=T) ---- */
int64_t ora_build_map(int Ns, int nel, int32_t *map); /* returns dim; map may be NULL to count */

/* ---- direct H x v, serial (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:23 + direct/ *.f90) ---- */
int ora_direct_hxv(const ora_params *p, int nup, int ndw, const double *v, double *Hv);

/* ---- a10: user two-body terms and phonons of the direct path ----
 * coulomb_matrix_element (ED_VARS_GLOBAL.f90:24-31): U cd_i cd_j c_k c_l, every operator as
 * (orbital 1-based, spin 1 = up / 2 = dw) -- element (1) is the orbital and (2) the spin in
 * direct/HxV_sundry.f90:18-21. */
typedef struct {
  int32_t cd_i[2], cd_j[2], c_k[2], c_l[2];
  double U;
} ora_sundry_term;
/* Nph, w0_ph, A_ph, g_ph (ED_INPUT_VARS.f90:184-198); DimPh = Nph+1 (ED_SETUP.f90:137).
 * The direct path has no A_ph term (direct/HxV_ph.f90:1-6), the stored one has
 * (stored/H_ph.f90:6-17): A_ph is applied here as the stored path does, 0 reproduces both. */
typedef struct {
  int32_t Nph, pad;
  double w0, A;
  double g[ORA_MAXORB][ORA_MAXORB];
} ora_phonons;
/* directMatVec_normal_main with DimPh>=1 and coulomb_sundry
 * (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:23-130; direct/HxV_ph.f90, HxV_eph.f90:1-81,
 * HxV_sundry.f90:1-109).  v/Hv: DimUp*DimDw*(Nph+1), i = iup+(idw-1)DimUp+(iph-1)DimUp*DimDw.
 * terms may be NULL (nsundry=0), ph may be NULL (Nph=0).  Returns -2 for a spin-unbalanced
 * sundry term (the reference STOPs, HxV_sundry.f90:35). */
int ora_direct_hxv_ext(const ora_params *p, int nup, int ndw, int nsundry, const ora_sundry_term *terms,
                       const ora_phonons *ph, const double *v, double *Hv);

/* ---- direct H x v, MPI algorithm with P emulated ranks
 *      (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:236 + direct_mpi/ *.f90,
 *       dw split ED_HAMILTONIAN_NORMAL.f90:128-142, transposes ..._COMMON.f90:66-178).
 *      v/Hv are the concatenation of the rank chunks (= the global vector, because the
 *      dw-column split is contiguous).  nthreads>1 runs ranks on OpenMP threads. ---- */
int ora_direct_hxv_mpi(const ora_params *p, int nup, int ndw, int P, int nthreads,
                       const double *v, double *Hv);

/* timing sample of the above: only ranks [0,nrun) of P are executed, *seconds = product time */
int ora_direct_hxv_mpi_sample(const ora_params *p, int nup, int ndw, int P, int nrun, int nthreads,
                              const double *v, double *Hv, double *seconds);

/* ---- same decomposition with precomputed H_up/H_dw hop tables (what ED_SPARSE_H=T
 *      does, ..._STORED_HxV.f90:765-867): "optimised CPU variant" of BASELINE.md 3.2 ---- */
int ora_stored_hxv_mpi(const ora_params *p, int nup, int ndw, int P, int nthreads, int ncalls,
                       const double *v, double *Hv, double *seconds_per_call);

/* ---- pass 1 of sp_lanc_eigh (SciFortran; call site ED_DIAG_NORMAL.f90:206-213) on that stored
 *      operator, recurrence and stopping rule as edipack_oracle.lanc_eigh; the lowest Ritz value
 *      by Sturm bisection.  v0 = start vector, alanc/blanc sized nitermax. ---- */
int ora_stored_lanczos_gs(const ora_params *p, int nup, int ndw, int P, int nthreads, int nitermax,
                          double threshold, int ncheck, const double *v0, double *egs, int *niter,
                          double *alanc, double *blanc, double *seconds);

/* ---- stored path pieces (ED_HAMILTONIAN_NORMAL_STORED_HxV.f90:26 + stored/ *.f90):
 *      H_up / H_dw as list-of-rows in insertion order, duplicates accumulated
 *      (ED_SPARSE_MATRIX.f90:328-357).  Output CSR arrays sized by a first counting call
 *      (pass NULL arrays).  spin = 0 (up) / 1 (dw); cols are 1-based like the reference. ---- */
int64_t ora_build_hop_csr(const ora_params *p, int spin, int nel, int64_t *rowptr, int32_t *cols,
                          double *vals);
/* diagonal spH0d over global rows (stored/H_local.f90) */
int ora_build_diag(const ora_params *p, int nup, int ndw, double *diag);
/* spH0nd rows in insertion order (stored/H_non_local.f90); cols are 1-based global */
int64_t ora_build_nonlocal_csr(const ora_params *p, int nup, int ndw, int64_t *rowptr,
                               int64_t *cols, double *vals);
/* stored SpMV, serial (..._STORED_HxV.f90:517) built from the pieces above */
int ora_stored_hxv(const ora_params *p, int nup, int ndw, const double *v, double *Hv);

/* ---- ED_SECTOR.f90:465 / :654 apply_op_C / apply_op_CDG (normal, ed_total_ud=T) ----
 * op = -1 destroys, +1 creates orbital iorb (0-based) of spin (0 up / 1 dw);
 * in sector (nup,ndw) -> out sector; out must hold the target sector dimension. */
int ora_apply_op(int Ns, int op, int iorb, int spin, int nup, int ndw, const double *v, double *ov);

#ifdef __cplusplus
}
#endif
#endif
