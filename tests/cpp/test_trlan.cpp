// CPU check of the thick-restart Lanczos host logic (edipack_b200/csrc/trlan.hpp) on a dense
// mock backend: test infrastructure only, never linked into libedgpu.so.
//   usage: test_trlan n nev ncv seed [ndegenerate]
// prints "ok <max eigenvalue error> <max residual> <orthogonality defect> <matvecs> <restarts>" or "FAIL ...".
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../edipack_b200/csrc/trlan.hpp"

struct DenseOps {
  int n, ncv;
  std::vector<double> H;                // row-major n x n
  std::vector<std::vector<double>> V;   // ncv+1 slots
  int matvec(int s, int d) {
    for (int i = 0; i < n; i++) {
      double acc = 0.0;
      for (int j = 0; j < n; j++) acc += H[(size_t)i * n + j] * V[s][j];
      V[d][i] = acc;
    }
    return 0;
  }
  int project_out(int m, int w, double *h, double *nb, double *na) {
    double b = 0.0;
    for (int i = 0; i < n; i++) b += V[w][i] * V[w][i];
    for (int k = 0; k < m; k++) {
      double acc = 0.0;
      for (int i = 0; i < n; i++) acc += V[k][i] * V[w][i];
      h[k] = acc;
    }
    for (int k = 0; k < m; k++)
      for (int i = 0; i < n; i++) V[w][i] -= h[k] * V[k][i];
    double a = 0.0;
    for (int i = 0; i < n; i++) a += V[w][i] * V[w][i];
    *nb = b;
    *na = a;
    return 0;
  }
  int scale(int w, double s) {
    for (auto &x : V[w]) x *= s;
    return 0;
  }
  int rotate(int m, int k, const double *Y) {
    std::vector<double> x(m);
    for (int i = 0; i < n; i++) {
      for (int a = 0; a < m; a++) x[a] = V[a][i];
      for (int j = 0; j < k; j++) {
        double acc = 0.0;
        for (int a = 0; a < m; a++) acc += x[a] * Y[(size_t)j * m + a];
        V[j][i] = acc;
      }
    }
    return 0;
  }
  int swap(int a, int b) {
    V[a].swap(V[b]);
    return 0;
  }
  int randomize(int w, uint64_t seed) {
    std::mt19937_64 g(seed);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    for (auto &x : V[w]) x = u(g);
    return 0;
  }
  int norm2(int w, double *out) {
    double a = 0.0;
    for (double x : V[w]) a += x * x;
    *out = a;
    return 0;
  }
};

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  const int n = atoi(argv[1]), nev = atoi(argv[2]), ncv = atoi(argv[3]);
  const uint64_t seed = strtoull(argv[4], nullptr, 10);
  const int ndeg = argc > 5 ? atoi(argv[5]) : 0;
  DenseOps ops;
  ops.n = n;
  ops.ncv = ncv;
  ops.H.assign((size_t)n * n, 0.0);
  std::mt19937_64 g(seed * 77 + 1);
  std::normal_distribution<double> nd(0.0, 1.0);
  if (ndeg > 0) {
    // H = Q D Q^T with the lowest `ndeg` eigenvalues equal (Q from Householder reflectors)
    std::vector<double> D(n);
    for (int i = 0; i < n; i++) D[i] = i < ndeg ? -3.0 : -2.0 + 0.05 * i;
    for (int i = 0; i < n; i++) ops.H[(size_t)i * n + i] = D[i];
    for (int r = 0; r < 3; r++) {
      std::vector<double> u(n);
      double nn = 0.0;
      for (auto &x : u) { x = nd(g); nn += x * x; }
      for (auto &x : u) x /= std::sqrt(nn);
      // H <- P H P, P = I - 2 u u^T
      std::vector<double> Hu(n, 0.0);
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Hu[i] += ops.H[(size_t)i * n + j] * u[j];
      double uHu = 0.0;
      for (int i = 0; i < n; i++) uHu += u[i] * Hu[i];
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
          ops.H[(size_t)i * n + j] += -2.0 * u[i] * Hu[j] - 2.0 * Hu[i] * u[j] + 4.0 * uHu * u[i] * u[j];
    }
  } else {
    // sparse-ish symmetric matrix with a spread diagonal (ED-like spectrum)
    for (int i = 0; i < n; i++) {
      ops.H[(size_t)i * n + i] = 4.0 * nd(g);
      for (int t = 0; t < 6; t++) {
        int j = (int)(g() % n);
        if (j == i) continue;
        double x = nd(g);
        ops.H[(size_t)i * n + j] += x;
        ops.H[(size_t)j * n + i] += x;
      }
    }
  }
  ops.V.assign(ncv + 1, std::vector<double>(n, 0.0));
  std::vector<double> A = ops.H, ref, Zref;
  edgpu::sym_eig_jacobi(n, A, ref, Zref);
  std::vector<double> ev(nev), rs(nev);
  edgpu::TrlanResult R;
  int rc = edgpu::trlan_solve(ops, n, nev, ncv, 500, 1e-14, seed, ev.data(), rs.data(), &R);
  if (rc) { printf("FAIL rc=%d\n", rc); return 1; }
  const int nw = std::min(nev, std::min(ncv, n));
  double eerr = 0.0, rmax = 0.0, orth = 0.0;
  for (int i = 0; i < nw; i++) {
    eerr = std::max(eerr, std::fabs(ev[i] - ref[i]));
    // true residual |H x - e x|
    double r2 = 0.0;
    for (int a = 0; a < n; a++) {
      double acc = 0.0;
      for (int b = 0; b < n; b++) acc += ops.H[(size_t)a * n + b] * ops.V[i][b];
      acc -= ev[i] * ops.V[i][a];
      r2 += acc * acc;
    }
    rmax = std::max(rmax, std::sqrt(r2));
    for (int j = 0; j <= i; j++) {
      double d = 0.0;
      for (int a = 0; a < n; a++) d += ops.V[i][a] * ops.V[j][a];
      orth = std::max(orth, std::fabs(d - (i == j ? 1.0 : 0.0)));
    }
  }
  const bool ok = R.nconv >= nw && eerr < 1e-10 && rmax < 1e-7 && orth < 1e-10;
  printf("%s %.3e %.3e %.3e %d %d nconv=%d\n", ok ? "ok" : "FAIL", eerr, rmax, orth, R.nmatvec, R.nrestart, R.nconv);
  return ok ? 0 : 1;
}
