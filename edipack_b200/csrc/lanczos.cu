// Device-resident Lanczos drivers: replacements for SciFortran SF_SP_LINALG
// sp_lanc_eigh / sp_lanc_tridiag as called from ED_DIAG_NORMAL.f90:206-213 and
// ED_HAMILTONIAN_NORMAL.f90:360-365.  SciFortran is not part of the reference tree
// (unpinned external dependency); the recurrence below is its published three-term form:
//
//   iter==1 : vin /= |vin|                       else : (vin,vout) <- (vout/beta, -beta*vin)
//   vout += H vin ; alfa = <vin,vout> ; vout -= alfa*vin ; beta = |vout|
//
// All Lanczos vectors stay in HBM; per iteration the host only sees alfa and beta.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>

#include "edgpu_internal.cuh"

namespace edgpu {

// Symmetric tridiagonal eigen-solver (implicit QL with Wilkinson shifts), the role of
// SciFortran's tql2 / LAPACK dstev used at ED_GF_NORMAL.f90:416.
// diag[n], sub[n-1] (sub[i] couples i,i+1).  evals ascending.  evecs: column-major n x n
// (evecs[j*n+i] = component i of eigenvector j) when want_vecs, else only ONE ROW of Z
// (track_row: 0 = Z(1,j), the weights of ED_GF_NORMAL.f90:416; n-1 = Z(n,j), the Ritz residual
// estimates) is returned in evecs[0..n-1].
int tridiag_eig(int n, const double *diag, const double *sub, double *evals, double *evecs,
                bool want_vecs, int track_row) {
  if (n <= 0) return 0;
  std::vector<double> d(diag, diag + n), e(n, 0.0);
  for (int i = 0; i + 1 < n; i++) e[i] = sub[i];
  const int zr = want_vecs ? n : 1;  // rows of Z that are tracked
  std::vector<double> z((size_t)zr * n, 0.0);  // z[r*n + j] : row r, eigenvector j
  if (want_vecs) {
    for (int r = 0; r < zr; r++) z[(size_t)r * n + r] = 1.0;
  } else {
    z[track_row] = 1.0;  // the single tracked row starts as row `track_row` of the identity
  }
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) return set_error("tridiag_eig: no convergence");
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i], b = c * e[i];
          e[i + 1] = (r = std::hypot(f, g));
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (p = s * r);
          g = c * r - b;
          for (int k = 0; k < zr; k++) {
            double *zk = &z[(size_t)k * n];
            f = zk[i + 1];
            zk[i + 1] = s * zk[i] + c * f;
            zk[i] = c * zk[i] - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
  for (int j = 0; j < n; j++) {
    evals[j] = d[idx[j]];
    if (evecs) {
      if (want_vecs)
        for (int i = 0; i < n; i++) evecs[(size_t)j * n + i] = z[(size_t)i * n + idx[j]];
      else
        evecs[j] = z[idx[j]];
    }
  }
  return 0;
}

// Un-normalised three-term recurrence: x = X_j with |X_j| = nx (v_j = X_j / nx), y = X_{j-1}
// with |X_{j-1}| = ny.  One step costs three passes over the vectors:
//   T   = (1/nx) H X_j - (beta_{j-1}/ny) X_{j-1}   written over X_{j-1}   (k_fast + k_slow; the
//         scalings ride in the kernels' epilogues, alfa_j = <X_j,T>/nx is fused into k_slow)
//   X_{j+1} = T - (alfa_j/nx) X_j ,  beta_j = |X_{j+1}|                   (k_axpy_norm)
// = 72 B/state instead of the 120 B/state of scale+swap / accumulate / dot / axpy, and the
// host sees alfa and beta only.
int g_lanczos_last_stored = 0, g_lanczos_last_hxv = 0;  // edgpu_lanczos_last_info
double g_lanczos_last_resid = 0.0;  // Ritz estimate |beta z_last| at exit (resid_tol mode)

void lanczos_release(Engine &E) {
  if (E.stream) cudaStreamSynchronize(E.stream);
  for (auto &c : E.lz_chunks) cudaFree(c.first);
  E.lz_chunks.clear();
  for (int k = 0; k < 3; k++) {
    cudaFree(E.lz_work[k]);
    E.lz_work[k] = nullptr;
    E.lz_work_len[k] = 0;
  }
}

int lanczos_work(Engine &E, int k, int64_t n, double **p) {
  if (!E.lz_work[k] || E.lz_work_len[k] < n) {
    cudaFree(E.lz_work[k]);
    E.lz_work[k] = nullptr;
    E.lz_work_len[k] = 0;
    EDGPU_CUDA(cudaMalloc(&E.lz_work[k], sizeof(double) * std::max<int64_t>(n, 1)));
    E.lz_work_len[k] = n;
  }
  *p = E.lz_work[k];
  return 0;
}

cudaError_t dev_malloc(void **p, size_t bytes) {
  cudaError_t e = (cudaMalloc)(p, bytes);  // the runtime's cudaMalloc, not the macro
  // (not while a ground-state solve is reading the pool: its stored vectors would be freed under it)
  if (e == cudaErrorMemoryAllocation && (!g.lz_chunks.empty() || g.lz_work[0]) && !g.lz_in_use) {
    cudaGetLastError();
    lanczos_release(g);
    e = (cudaMalloc)(p, bytes);
  }
  return e;
}

struct LanczosVecs {
  double *x = nullptr, *y = nullptr;
  double nx = 1.0, ny = 1.0, beta_prev = 0.0;
  double *work[2] = {nullptr, nullptr};  // scratch vectors the recurrence may overwrite
  bool is_work(const double *p) const { return p == work[0] || p == work[1]; }
};

// One step; returns alfa, beta on the host and leaves L advanced (x = X_{j+1}, nx = beta).
// v_j / nx (the normalised Lanczos vector of this step) is L_before.x: callers that need it
// (pass 2 of the ground-state driver) read x and nx BEFORE calling.
// slot != nullptr (ground-state driver with the HBM vector store): X_{j+1} is produced directly in
// `slot` -- pass B reads the old term from X_{j-1} (L.y, left intact in its own slot) and writes T
// there, pass A and the update work in place -- so no copy of the new vector is needed
// (72 B/state/iteration instead of 80).  Without a slot T overwrites X_{j-1}, which must then be one
// of the scratch vectors (it is copied into one the first time it is not).
static int lanczos_step(Engine &E, int iter, LanczosVecs &L, double *alfa, double *beta,
                        double *slot = nullptr) {
  if (iter == 1) {
    double n2;
    EDGPU_TRY(vec_dot(E, L.x, L.x, &n2));
    if (n2 == 0.0) return set_error("lanczos_iteration: norm = 0");
    L.nx = std::sqrt(n2);
    L.ny = 1.0;
    L.beta_prev = 0.0;
    EDGPU_TRY(vec_zero(E, L.y, E.veclen()));
  }
  double *d_dot = E.d_scal + 1;
  double *t = slot;  // where T and then X_{j+1} live
  if (!t) {
    if (!L.is_work(L.y)) {  // X_{j-1} sits in a store slot: continue on a scratch copy
      double *w = (L.x == L.work[0]) ? L.work[1] : L.work[0];
      EDGPU_CUDA(cudaMemcpyAsync(w, L.y, sizeof(double) * (size_t)E.veclen(), cudaMemcpyDeviceToDevice,
                                 E.stream));
      L.y = w;
    }
    t = L.y;
  }
  EDGPU_TRY(hxv_device_ex(E, L.x, t, /*accum=*/true, /*timed=*/false, 1.0 / L.nx,
                          -L.beta_prev / L.ny, d_dot, /*d_old=*/L.y));
  double xt;
  EDGPU_TRY(scalar_to_host(E, d_dot, &xt));
  *alfa = xt / L.nx;
  double b2;
  EDGPU_TRY(vec_axpy_norm(E, t, L.x, *alfa / L.nx, &b2));  // t <- X_{j+1}
  *beta = std::sqrt(b2);
  L.y = L.x;
  L.x = t;
  L.ny = L.nx;
  L.nx = *beta;
  L.beta_prev = *beta;
  return 0;
}

// sp_lanc_tridiag: alanc(iter)=alfa, blanc(iter+1)=beta, early exit when |beta|<threshold.
// d_seed is consumed (used as vin).  d_work is a second vector of the same padded length.
int lanczos_tridiag_dev(Engine &E, double *d_seed, double *d_work, int nlanc, double threshold,
                        double *alanc, double *blanc, int *nused) {
  LanczosVecs L;
  L.x = d_seed;
  L.y = d_work;
  L.work[0] = d_seed;
  L.work[1] = d_work;
  for (int i = 0; i < nlanc; i++) alanc[i] = blanc[i] = 0.0;
  double alfa = 0.0, beta = 0.0;
  *nused = 0;
  for (int it = 1; it <= nlanc; it++) {
    EDGPU_TRY(lanczos_step(E, it, L, &alfa, &beta));
    alanc[it - 1] = alfa;
    *nused = it;
    if (std::fabs(beta) < threshold) break;
    if (it < nlanc) blanc[it] = beta;
  }
  return 0;
}

// sp_lanc_eigh: pass 1 builds T until the lowest Ritz value is stationary (checked every
// iteration once nlanc >= ncheck) or beta -> 0 or nitermax; pass 2 forms
// vect = sum_iter Z(iter,1) * v_iter, normalised.
// SciFortran's driver keeps two vectors and therefore re-runs the whole recurrence for pass 2
// (twice the H x v count).  With 180 GB of HBM per GPU the Lanczos vectors of pass 1 are kept
// instead (written by the same kernel that produces them, +8 B/state per iteration) as long as
// device memory lasts: pass 2 is then one streaming linear combination, and only the iterations
// whose vectors no longer fitted are replayed, restarting the recurrence from the last two stored
// vectors.  The result equals a full replay up to the summation order of the final combination
// (the replay regenerates exactly the stored vectors).  EDGPU_LANCZOS_STORE=0 disables the store
// (two-vector mode).
// d_start: start vector (kept intact, copied) or nullptr for the seeded random start.
// d_vect: output (padded length).
// resid_tol > 0 replaces the stationarity test by ARPACK's test on the Ritz estimate of the lowest
// pair, |beta_j Z(j,1)| <= resid_tol * max(eps^(2/3), |theta_1|) (the Neigen=1 route of edgpu_eigh).
int lanczos_gs_dev(Engine &E, int nitermax, double threshold, int ncheck, const double *d_start,
                   uint64_t seed, double *egs, double *d_vect, int *niter, double resid_tol) {
  const int64_t n = E.veclen();
  const int64_t dim_global = E.csr.open ? E.csr.nglobal : E.sec.up.dim * E.sec.dw.dim;
  // EDGPU_LANCZOS_TIMING=1: host wall-clock of the phases of one solve on stderr (diagnostics)
  static const bool timing = getenv("EDGPU_LANCZOS_TIMING") && getenv("EDGPU_LANCZOS_TIMING")[0] == '1';
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  const auto t_begin = now();
  auto t_alloc = t_begin, t_loop = t_begin, t_pass2 = t_begin;
  double t_hxv = 0.0, t_sync = 0.0, t_eig = 0.0;
  if (nitermax > dim_global) nitermax = (int)dim_global;
  if (nitermax < 1) nitermax = 1;
  if (ncheck < 1) ncheck = 1;
  double *vin = nullptr, *vout = nullptr;
  std::vector<double *> store;  // store[j] = X_{j+1} (un-normalised Lanczos vector j+1)
  EDGPU_TRY(lanczos_work(E, 0, n, &vin));   // cached across solves (see Engine::lz_work)
  EDGPU_TRY(lanczos_work(E, 1, n, &vout));
  auto cleanup = [&]() {
    store.clear();  // slots of the pooled chunks (E.lz_chunks), kept for the next solve
    E.lz_in_use = false;
  };
  const char *env = getenv("EDGPU_LANCZOS_STORE");
  bool storing = !(env && env[0] == '0');
  // Vector slots are carved from the pooled chunks; new chunks (8 vectors, at least 256 MB) are
  // added on demand while device memory lasts, keeping a reserve for everything else that is
  // allocated while sectors are open (stored states, seeds, transposes, the callers' buffers).
  // With nranks > 1 every decision about the store is COLLECTIVE: pass 2 replays the unstored
  // tail with all-reduces and transposes, so all ranks must end with the same number of stored
  // vectors.  The slot size is the largest local vector of any rank (the dw split differs by one
  // column), the usable part of the existing pool is the minimum over ranks, and a new chunk is
  // kept only if every rank got one.
  size_t vbytes = sizeof(double) * (size_t)n;
  int64_t pool_slots = 0;  // slots of the existing pool that every rank has
  auto slots_in_pool = [&]() {
    int64_t c = 0;
    for (auto &ch : E.lz_chunks) c += (int64_t)(ch.second / vbytes);
    return c;
  };
  if (E.nranks > 1) {
    double mx = (double)vbytes;
    if (comm_allreduce_host(E, &mx, 1, /*op max*/ 2)) { cleanup(); return g_status; }
    vbytes = (size_t)mx;
    double mn = (double)slots_in_pool();
    if (comm_allreduce_host(E, &mn, 1, /*op min*/ 3)) { cleanup(); return g_status; }
    pool_slots = (int64_t)mn;
  } else {
    pool_slots = slots_in_pool();
  }
  E.lz_in_use = true;  // dev_malloc must not release the pool under this solve
  size_t chunk_i = 0, chunk_used = 0;  // carving position
  // EDGPU_LANCZOS_MAXSTORE=k (testing): keep at most k vectors, as if device memory had run out
  const int max_store = getenv("EDGPU_LANCZOS_MAXSTORE") ? atoi(getenv("EDGPU_LANCZOS_MAXSTORE")) : -1;
  auto try_store_slot = [&]() -> double * {
    if (max_store >= 0 && (int)store.size() >= max_store) storing = false;
    if (!storing || (int)store.size() > nitermax) {
      storing = false;
      return nullptr;
    }
    if ((int64_t)store.size() >= pool_slots) {
      // every rank arrives here at the same iteration: collective growth of the pool
      size_t fr = 0, tot = 0;
      const size_t reserve = std::max<size_t>((size_t)4 << 30, 6 * vbytes);
      const size_t left = (size_t)nitermax + 1 - store.size();
      const size_t want = std::min(left, std::max<size_t>(8, ((size_t)256 << 20) / vbytes)) * vbytes;
      void *p = nullptr;
      bool ok = cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr >= reserve + want &&
                (cudaMalloc)(&p, want) == cudaSuccess;
      if (!ok) cudaGetLastError();
      if (E.nranks > 1) {
        double f = ok ? 1.0 : 0.0;
        if (comm_allreduce_host(E, &f, 1, /*op min*/ 3)) f = 0.0;
        if (f == 0.0 && ok) {
          cudaFree(p);
          ok = false;
        }
      }
      if (!ok) {
        storing = false;
        return nullptr;
      }
      E.lz_chunks.emplace_back((double *)p, want);
      pool_slots += (int64_t)(want / vbytes);
    }
    while (chunk_i < E.lz_chunks.size() && chunk_used + vbytes > E.lz_chunks[chunk_i].second) {
      chunk_i++;
      chunk_used = 0;
    }
    if (chunk_i == E.lz_chunks.size()) {  // cannot happen: pool_slots counts whole slots
      storing = false;
      return nullptr;
    }
    double *slot = (double *)((char *)E.lz_chunks[chunk_i].first + chunk_used);
    chunk_used += vbytes;
    store.push_back(slot);
    return slot;
  };
  if (timing) {
    cudaStreamSynchronize(E.stream);
    t_alloc = now();
  }
  if (d_start) {
    EDGPU_CUDA(cudaMemcpyAsync(vin, d_start, sizeof(double) * n, cudaMemcpyDeviceToDevice, E.stream));
  } else {
    int rc0 = vec_fill_random(E, vin, seed);
    if (rc0) { cleanup(); return rc0; }
  }
  if (double *p0 = try_store_slot())  // X_1
    cudaMemcpyAsync(p0, vin, sizeof(double) * n, cudaMemcpyDeviceToDevice, E.stream);
  std::vector<double> a, b(1, 0.0), ev, esave, nrm, zlast;  // nrm[j] = |X_{j+1}|
  g_lanczos_last_resid = 0.0;
  LanczosVecs L;
  L.x = vin;
  L.y = vout;
  L.work[0] = vin;
  L.work[1] = vout;
  double alfa = 0.0, beta = 0.0;
  int nlanc = 0, rc = 0;
  for (int it = 1; it <= nitermax; it++) {
    // X_{it+1} is only worth keeping while the chain of stored vectors is unbroken
    const auto t_s0 = now();
    double *slot = (int)store.size() == it ? try_store_slot() : nullptr;
    if (timing) t_sync += secs(t_s0, now());  // (time spent growing the vector pool)
    const auto t_s1 = now();
    rc = lanczos_step(E, it, L, &alfa, &beta, slot);
    if (timing) t_hxv += secs(t_s1, now());
    if (rc) { cleanup(); return rc; }
    if (it == 1) nrm.push_back(L.ny);  // |X_1| (the step normalised with it)
    nrm.push_back(beta);               // |X_{it+1}|
    a.push_back(alfa);
    nlanc = it;
    // beta -> 0: the Krylov space is exhausted (also at it == 1, when the start vector is an
    // eigenvector: dividing by |X_2| = 0 in the next step would give NaNs)
    if (std::fabs(beta) < threshold) break;
    b.push_back(beta);
    if (nlanc >= ncheck) {
      ev.resize(nlanc);
      if (resid_tol > 0.0) {
        zlast.resize(nlanc);
        rc = tridiag_eig(nlanc, a.data(), b.data() + 1, ev.data(), zlast.data(), false, nlanc - 1);
        if (rc) { cleanup(); return rc; }
        g_lanczos_last_resid = std::fabs(beta * zlast[0]);
        if (g_lanczos_last_resid <= resid_tol * std::max(3.666852862501036e-11, std::fabs(ev[0]))) break;
      } else {
        rc = tridiag_eig(nlanc, a.data(), b.data() + 1, ev.data(), nullptr, false);
        if (rc) { cleanup(); return rc; }
        esave.push_back(ev[0]);
        if (esave.size() >= 2 &&
            std::fabs(esave[esave.size() - 1] - esave[esave.size() - 2]) <= threshold)
          break;
      }
    }
  }
  t_loop = now();
  ev.resize(nlanc);
  std::vector<double> Z((size_t)nlanc * nlanc);
  rc = tridiag_eig(nlanc, a.data(), b.data() + 1, ev.data(), Z.data(), true);
  if (timing) t_eig = secs(t_loop, now());
  if (rc) { cleanup(); return rc; }
  *egs = ev[0];
  *niter = nlanc;
  // pass 2: vect = sum_j Z(j,1) X_j / |X_j|
  const int ns = std::min((int)store.size(), nlanc);  // stored: X_1 .. X_ns
  g_lanczos_last_stored = ns;
  g_lanczos_last_hxv = nlanc + ((ns >= 2 || ns >= nlanc) ? nlanc - ns : nlanc - 1);
  if (ns >= 2 || ns >= nlanc) {
    std::vector<double *> vs(store.begin(), store.begin() + ns);
    std::vector<double> cf(ns);
    for (int j = 0; j < ns; j++) cf[j] = Z[j] / nrm[j];
    rc = vec_lincomb(E, d_vect, vs, cf);
    if (rc) { cleanup(); return rc; }
    if (ns < nlanc) {
      // replay the tail: recurrence state after step ns-1 is x = X_ns, y = X_{ns-1}
      EDGPU_CUDA(cudaMemcpyAsync(vin, store[ns - 1], sizeof(double) * n, cudaMemcpyDeviceToDevice, E.stream));
      EDGPU_CUDA(cudaMemcpyAsync(vout, store[ns - 2], sizeof(double) * n, cudaMemcpyDeviceToDevice, E.stream));
      L = LanczosVecs();
      L.x = vin;
      L.y = vout;
      L.work[0] = vin;
      L.work[1] = vout;
      L.nx = nrm[ns - 1];
      L.ny = nrm[ns - 2];
      L.beta_prev = nrm[ns - 1];  // beta_{ns-1} = |X_ns|
      for (int it = ns; it < nlanc; it++) {
        rc = lanczos_step(E, it, L, &alfa, &beta);  // -> x = X_{it+1}
        if (rc) { cleanup(); return rc; }
        rc = vec_axpy(E, d_vect, L.x, Z[it] / L.nx);
        if (rc) { cleanup(); return rc; }
      }
    }
  } else {
    // two-vector mode: replay the recurrence from the same start vector
    if (d_start) {
      EDGPU_CUDA(cudaMemcpyAsync(vin, d_start, sizeof(double) * n, cudaMemcpyDeviceToDevice, E.stream));
    } else {
      rc = vec_fill_random(E, vin, seed);
      if (rc) { cleanup(); return rc; }
    }
    vec_zero(E, d_vect, n);
    L = LanczosVecs();
    L.x = vin;
    L.y = vout;
    L.work[0] = vin;
    L.work[1] = vout;
    for (int it = 1; it <= nlanc; it++) {
      // v_it = x / nx with (x, nx) as they are BEFORE the step (iteration 1 normalises inside)
      if (it == 1) {
        double n2s;
        rc = vec_dot(E, L.x, L.x, &n2s);
        if (rc) { cleanup(); return rc; }
        L.nx = std::sqrt(n2s);
      }
      rc = vec_axpy(E, d_vect, L.x, Z[it - 1] / L.nx);  // Z(iter,1): component iter of eigenvector 1
      if (rc) { cleanup(); return rc; }
      if (it == nlanc) break;  // the last vector needs no further H x v
      rc = lanczos_step(E, it, L, &alfa, &beta);
      if (rc) { cleanup(); return rc; }
    }
  }
  double n2;
  rc = vec_dot(E, d_vect, d_vect, &n2);
  if (!rc) rc = vec_scale(E, d_vect, 1.0 / std::sqrt(n2));
  cudaStreamSynchronize(E.stream);
  t_pass2 = now();
  cleanup();
  if (timing)
    fprintf(stderr,
            "[edgpu lanczos rank %d] alloc %.4f s, recurrence %.4f s (steps %.4f, pool growth %.4f), final eig %.4f, "
            "pass 2 %.4f s, free %.4f s; %d iterations, %d vectors kept\n",
            E.rank, secs(t_begin, t_alloc), secs(t_alloc, t_loop), t_hxv, t_sync, t_eig, secs(t_loop, t_pass2),
            secs(t_pass2, now()), nlanc, ns);
  return rc;
}

}  // namespace edgpu
