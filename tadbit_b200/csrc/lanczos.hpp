// Largest eigenpairs of a dense symmetric matrix by Lanczos with full re-orthogonalisation:
// the host-side control of N4 (compartments).  The reference asks
// scipy.sparse.linalg.eigsh(matrix, k=max_ev) for them (_pytadbit/hic_data.py:926-937: ARPACK,
// the k eigenvalues of largest magnitude, returned in ascending algebraic order and read back
// to front) and, when k >= n, gets every eigenpair from a dense solver instead.
//
// This header holds no device code: the vector operations come from `Ops` (compartments.cu
// runs them as CUDA kernels over the correlation matrix resident in HBM;
// tests/native/lanczos_host.cpp instantiates the same control code over plain arrays so that
// break-down, restart, selection and ordering are tested without a GPU).
//
//   Ops:  void   random_w(uint64_t seed);            w = reproducible pseudo-random vector
//         void   project(int m);                     w -= sum_k (V[k] . w) V[k], k < m
//         double norm_w();                           ||w||                              (synchronises)
//         void   keep(int slot, double scale);       V[slot] = scale * w
//         void   step(int j, bool store_next);       one Lanczos step WITHOUT a round trip to the host:
//                                                    w = A V[j]; project(j + 1) twice; alpha[j] = the
//                                                    two coefficients on V[j]; beta[j] = ||w||;
//                                                    V[j + 1] = w / beta[j] if store_next
//         void   read_ab(int j0, int j1, double *alpha, double *beta);
//                                                    alpha[j], beta[j] of steps j0 <= j < j1   (synchronises)
//         void   combine(int m, int k, const double *coef, double *out);
//                                                    out[i] = normalised sum_j coef[i * m + j] V[j]   (host, k x n)
//
// The steps between two convergence tests are queued back to back and their scalars read in one go:
// on a chromosome of a thousand bins a step is a few microseconds of kernels, and three round
// trips per step (the first version) were 90 % of the solver's time.  A break-down inside such a
// burst shows as a vanishing beta when the scalars arrive; the steps queued after it worked on
// garbage and are dropped.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace tb {

// Eigen-decomposition of a symmetric tridiagonal matrix (implicit QL with Wilkinson shifts).
// d[0..n): diagonal, e[0..n-1): sub-diagonal (e[i] couples i and i + 1).  On return d holds the
// eigenvalues (unordered) and z (n x n, row-major) the eigenvectors: z[r * n + i] = component r of
// the eigenvector of d[i].  Returns false if an eigenvalue needed more than 120 sweeps.
inline bool tridiag_eig(int n, std::vector<double> &d, std::vector<double> &e_in, std::vector<double> &z) {
  std::vector<double> e(n, 0.0);
  for (int i = 0; i + 1 < n; ++i) e[i] = e_in[i];
  z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) z[(size_t)i * n + i] = 1.0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= eps * dd) break;
      }
      if (m != l) {
        if (iter++ == 120) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          r = hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (int k = 0; k < n; ++k) {
            double *row = &z[(size_t)k * n];
            f = row[i + 1];
            row[i + 1] = s * row[i] + c * f;
            row[i] = c * row[i] - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

struct LanczosResult {
  std::vector<double> evals;      // the selected eigenvalues, descending (algebraic)
  int steps = 0;                  // Lanczos vectors built
  int restarts = 0;               // invariant subspaces met (new start vector each time)
  int round_trips = 0;            // bursts of steps = reads of the scalars
  bool converged = false;
  double residual = 0.0;          // largest residual estimate of the selected pairs, relative to max |theta|
};

// The k eigenpairs of largest magnitude (all_by_value: every eigenpair, k == n) of the n x n
// matrix behind `ops`; eigenvectors (unit norm, rows of h_vecs[k x n]) in descending algebraic
// order of their eigenvalues, which is the order the reference reads them in.
template <typename Ops>
LanczosResult lanczos_largest(Ops &ops, int64_t n, int k, int m_cap, double tol, bool all_by_value,
                              double *h_vecs) {
  LanczosResult res;
  if (m_cap > n) m_cap = (int)n;
  if (k > m_cap) k = m_cap;
  // T: alpha[j] on the diagonal; resid[j] = ||w|| after step j couples j and j + 1 unless the
  // iteration was restarted there (cut[j])
  std::vector<double> alpha, resid, d, e, z, a, b;
  std::vector<char> cut;
  const uint64_t seed = 0x7ad0b17b200ULL;

  ops.random_w(seed);
  ops.keep(0, 1.0 / ops.norm_w());
  double anorm = 0.0;                        // running estimate of ||A||
  int next_check = (int)std::min<int64_t>(n, std::max(2 * k + 1, 8));
  int m = 0;                                 // accepted steps = size of T
  bool solved = false;
  while (m < m_cap) {
    const int stop = std::min(std::max(next_check, m + 1), m_cap);
    for (int j = m; j < stop; ++j) ops.step(j, j + 1 < m_cap);
    a.resize(stop - m);
    b.resize(stop - m);
    ops.read_ab(m, stop, a.data(), b.data());
    ++res.round_trips;
    bool breakdown = false;
    for (int j = m; j < stop && !breakdown; ++j) {
      alpha.push_back(a[j - m]);
      resid.push_back(b[j - m]);
      cut.push_back(0);
      anorm = std::max(anorm, fabs(alpha[j]) + resid[j] + ((j && !cut[j - 1]) ? resid[j - 1] : 0.0));
      breakdown = !(resid[j] > 1e-13 * std::max(anorm, 1e-300));      // also catches NaN
    }
    m = (int)alpha.size();
    const bool last = m == m_cap;
    d = alpha;
    e.assign(m, 0.0);
    for (int j = 0; j + 1 < m; ++j) e[j] = cut[j] ? 0.0 : resid[j];
    solved = tridiag_eig(m, d, e, z);
    if (!solved) break;
    // residual estimate of a Ritz pair: ||w|| * |last component of its eigenvector of T|
    std::vector<int> idx(m);
    for (int i = 0; i < m; ++i) idx[i] = i;
    if (all_by_value) std::sort(idx.begin(), idx.end(), [&](int p, int q) { return d[p] > d[q]; });
    else std::sort(idx.begin(), idx.end(), [&](int p, int q) { return fabs(d[p]) > fabs(d[q]); });
    double tmax = 0.0, worst = 0.0;
    for (int i = 0; i < m; ++i) tmax = std::max(tmax, fabs(d[i]));
    const int kk = std::min(k, m);
    for (int i = 0; i < kk; ++i)
      worst = std::max(worst, (breakdown ? 0.0 : resid[m - 1]) * fabs(z[(size_t)(m - 1) * m + idx[i]]));
    res.residual = tmax > 0.0 ? worst / tmax : 0.0;
    const bool ok = m >= k && worst <= tol * std::max(tmax, 1e-300);
    if (ok || m == n || last) {
      // out of vectors: a pair whose residual is below 1e-9 ||A|| is still an eigenpair to ~1e-9
      // (eigenvalue to ~1e-18 / gap); anything worse is reported as not converged
      res.converged = ok || m == n || (m >= k && worst <= 1e-9 * std::max(tmax, 1e-300));
      break;
    }
    next_check = m + std::max(2, m / 6);
    if (breakdown) {
      // the Krylov space is invariant: carry on from a fresh vector orthogonal to it
      ++res.restarts;
      ops.random_w(seed + (uint64_t)res.restarts);
      ops.project(m);
      ops.project(m);
      const double nb = ops.norm_w();
      if (!(nb > 1e-8)) {                      // nothing left outside the space: T is exact
        res.converged = true;
        break;
      }
      ops.keep(m, 1.0 / nb);
      cut[m - 1] = 1;
    }
  }
  res.steps = m;
  if (!solved || m == 0) {
    res.converged = false;
    return res;
  }
  // selection: largest magnitude (or all), then descending algebraic order
  std::vector<int> idx(m);
  for (int i = 0; i < m; ++i) idx[i] = i;
  if (all_by_value) std::sort(idx.begin(), idx.end(), [&](int p, int q) { return d[p] > d[q]; });
  else std::stable_sort(idx.begin(), idx.end(), [&](int p, int q) { return fabs(d[p]) > fabs(d[q]); });
  const int kk = std::min(k, m);
  idx.resize(kk);
  std::sort(idx.begin(), idx.end(), [&](int p, int q) { return d[p] > d[q]; });
  std::vector<double> coef((size_t)kk * m);
  for (int i = 0; i < kk; ++i) {
    res.evals.push_back(d[idx[i]]);
    for (int j = 0; j < m; ++j) coef[(size_t)i * m + j] = z[(size_t)j * m + idx[i]];
  }
  ops.combine(m, kk, coef.data(), h_vecs);
  return res;
}

}  // namespace tb
