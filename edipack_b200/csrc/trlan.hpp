// Thick-restart Lanczos driver (host logic only, no CUDA in this header): the multi-eigenpair
// solver that stands where the reference calls SciFortran's sp_eigh = (P-)ARPACK dsaupd/dseupd
// with which='SA' (call sites ED_DIAG_NORMAL.f90:179-192, ED_DIAG_NONSU2.f90:179-192,
// ED_DIAG_SUPERC.f90:161-174):
//
//     call sp_eigh([MpiComm,]MatVec, eval(Neigen), evec(Nloc,Neigen), Nblock, Nitermax, tol=)
//
// ARPACK itself is not part of the reference tree.  Its method for symmetric problems -- an
// implicitly restarted Lanczos process with full re-orthogonalisation of an `ncv`-vector basis
// -- is mathematically equivalent to thick restarting (Wu & Simon, SIAM J. Matrix Anal. 22, 602)
// which is what is written here; the interface keeps ARPACK's knobs (nev, ncv=Nblock,
// maxiter=Nitermax restarts, tol with the `tol*max(eps^(2/3),|theta|)` Ritz-estimate test).
//
// The driver is a template over the vector backend so that the same code runs on the device
// backend of eigs.cu (all basis vectors in HBM) and on the dense mock of tests/cpp/test_trlan.cpp.
//
// Backend concept (vectors are addressed by slot 0..ncv):
//   int matvec(int src, int dst)                  V[dst] = H V[src]
//   int project_out(int m, int w, double *h, double *nrm2_before, double *nrm2_after)
//                                                 h = V[0..m)^H V[w] (real parts returned),
//                                                 V[w] -= V[0..m) h, squared norms of V[w]
//   int scale(int w, double s)                    V[w] *= s
//   int rotate(int m, int k, const double *Y)     V[0..k) = V[0..m) Y   (Y column-major m x k)
//   int swap(int a, int b)                        exchange two slots
//   int randomize(int w, uint64_t seed)           fill V[w] with the seeded start vector
//   int norm2(int w, double *out)
// Every call returns 0 on success.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace edgpu {

// Cyclic Jacobi eigen-solver for the small dense projected matrix (n <= ~128): A symmetric,
// row-major n x n (destroyed); evals ascending; Z column-major (Z[j*n+i] = component i of
// eigenvector j).  Plays the role of LAPACK dsteqr inside ARPACK's dseigt.
inline int sym_eig_jacobi(int n, std::vector<double> &A, std::vector<double> &evals,
                          std::vector<double> &Z) {
  std::vector<double> V((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < n; i++) {
      dg += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; j++) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-34 * (dg + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; k++) {  // columns p,q
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {  // rows p,q
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(),
            [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
  evals.resize(n);
  Z.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; j++) {
    evals[j] = A[(size_t)idx[j] * n + idx[j]];
    for (int i = 0; i < n; i++) Z[(size_t)j * n + i] = V[(size_t)i * n + idx[j]];
  }
  return 0;
}

struct TrlanResult {
  int nconv = 0;      // leading wanted Ritz pairs that passed the convergence test
  int nmatvec = 0;    // H x v products
  int nrestart = 0;   // restarts done
  int nbasis = 0;     // size of the final projected problem
};

// Lowest `nev` eigenpairs of the symmetric / Hermitian operator behind `ops`.  On return the
// eigenvectors sit in slots 0..nev-1 (orthonormal), evals[0..nev) ascending, resid[0..nev) the
// Ritz estimates |beta * y_last|.  dim = global dimension of the operator.
template <class Ops>
int trlan_solve(Ops &ops, int64_t dim, int nev, int ncv, int maxiter, double tol, uint64_t seed,
                double *evals, double *resid, TrlanResult *res) {
  const double eps = 2.220446049250313e-16;
  const double eps23 = std::pow(eps, 2.0 / 3.0);
  if ((int64_t)ncv > dim) ncv = (int)dim;
  if (nev > ncv) nev = ncv;
  if (nev < 1 || ncv < 1) return 1;
  if (tol < eps) tol = eps;  // ARPACK: tol <= 0 means machine precision; nothing below it is meaningful here
  if (maxiter < 1) maxiter = 1;
  std::vector<double> T((size_t)ncv * ncv, 0.0), A, theta, Y, h(ncv + 1), h2(ncv + 1);
  int k = 0;  // kept Ritz vectors (thick part of the basis)
  int rc;
  if ((rc = ops.randomize(0, seed))) return rc;
  double n2;
  if ((rc = ops.norm2(0, &n2))) return rc;
  if (n2 <= 0.0) return 2;
  if ((rc = ops.scale(0, 1.0 / std::sqrt(n2)))) return rc;
  TrlanResult R;
  double beta_last = 0.0;
  double tnorm = 0.0;  // running estimate of |H| for the breakdown test
  int m = ncv;         // size of the projected problem of this cycle
  bool exhausted = false;
  for (int cycle = 0;; cycle++) {
    m = ncv;
    for (int j = k; j < ncv; j++) {
      if ((rc = ops.matvec(j, j + 1))) return rc;
      R.nmatvec++;
      double nb, na;
      if ((rc = ops.project_out(j + 1, j + 1, h.data(), &nb, &na))) return rc;
      double alpha = h[j];
      // "twice is enough" (Kahan/Parlett; the DGKS refinement of ARPACK's dsaitr): repeat the
      // projection when it removed most of the vector.  eta^2 = 0.1: a Lanczos step legitimately
      // removes alpha v_j + beta v_{j-1}, typically half of |H v_j|^2, which DGKS's eta = 1/sqrt(2)
      // would answer with a second pass over the whole basis on every other step; the rounding
      // error of one pass is amplified by at most 1/eta ~ 3.2, i.e. orthogonality stays at a few eps.
      if (na < 0.1 * nb) {
        double nb2;
        if ((rc = ops.project_out(j + 1, j + 1, h2.data(), &nb2, &na))) return rc;
        alpha += h2[j];
      }
      T[(size_t)j * ncv + j] = alpha;
      const double beta = std::sqrt(std::max(na, 0.0));
      tnorm = std::max(tnorm, std::fabs(alpha) + beta);
      beta_last = beta;
      if (beta <= eps * std::max(tnorm, 1e-300) * 16.0) {
        // invariant subspace: the Ritz values of the (j+1)-dimensional problem are exact
        beta_last = 0.0;
        if (j + 1 >= nev || (int64_t)(j + 1) >= dim) {
          m = j + 1;
          exhausted = true;
          break;
        }
        // fewer exact pairs than wanted: continue with a fresh direction orthogonal to the basis
        if ((rc = ops.randomize(j + 1, seed + 7919u * (uint64_t)(R.nmatvec + 1)))) return rc;
        double b0, a0;
        if ((rc = ops.project_out(j + 1, j + 1, h2.data(), &b0, &a0))) return rc;
        if ((rc = ops.project_out(j + 1, j + 1, h2.data(), &b0, &a0))) return rc;
        if (a0 <= 0.0) return 3;
        if ((rc = ops.scale(j + 1, 1.0 / std::sqrt(a0)))) return rc;
        continue;  // T(j,j+1) stays 0
      }
      if ((rc = ops.scale(j + 1, 1.0 / beta))) return rc;
      if (j + 1 < ncv) T[(size_t)j * ncv + j + 1] = T[(size_t)(j + 1) * ncv + j] = beta;
    }
    if ((int64_t)m >= dim) exhausted = true;  // the basis spans the whole space: T is exact
    if (exhausted) beta_last = 0.0;
    // Ritz pairs of the m x m projected matrix
    A.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++) A[(size_t)i * m + j] = T[(size_t)i * ncv + j];
    sym_eig_jacobi(m, A, theta, Y);
    R.nbasis = m;
    const int nwant = std::min(nev, m);
    R.nconv = 0;
    for (int i = 0; i < nwant; i++) {
      const double r = std::fabs(beta_last * Y[(size_t)i * m + (m - 1)]);
      if (resid) resid[i] = r;
      if (r <= tol * std::max(eps23, std::fabs(theta[i])) && R.nconv == i) R.nconv = i + 1;
    }
    if (exhausted) R.nconv = nwant;
    R.nrestart = cycle;
    const bool done = R.nconv >= nwant || cycle + 1 >= maxiter || m < ncv || ncv <= nev;
    if (done) {
      // eigenvectors -> slots 0..nwant-1
      if ((rc = ops.rotate(m, nwant, Y.data()))) return rc;
      for (int i = 0; i < nwant; i++) evals[i] = theta[i];
      for (int i = nwant; i < nev; i++) evals[i] = 0.0;
      break;
    }
    // thick restart: keep the k lowest Ritz vectors (the wanted ones plus a share of the rest,
    // moving with the converged count like ARPACK's kev adjustment), then the residual vector
    int keep = nev + std::min(R.nconv, (ncv - nev) / 2) + (ncv - nev) / 3;
    keep = std::max(keep, nev);
    keep = std::min(keep, ncv - 1);
    if ((rc = ops.rotate(ncv, keep, Y.data()))) return rc;
    if ((rc = ops.swap(keep, ncv))) return rc;
    std::fill(T.begin(), T.end(), 0.0);
    for (int i = 0; i < keep; i++) {
      T[(size_t)i * ncv + i] = theta[i];
      const double s = beta_last * Y[(size_t)i * ncv + (ncv - 1)];
      T[(size_t)i * ncv + keep] = T[(size_t)keep * ncv + i] = s;
    }
    k = keep;
  }
  if (res) *res = R;
  return 0;
}

}  // namespace edgpu
