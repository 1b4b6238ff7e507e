// TEST INFRASTRUCTURE (not part of the product library): instantiates the Lanczos control code of
// tadbit_b200/csrc/lanczos.hpp over plain host arrays, so that break-down, restart, selection and
// ordering can be tested without a GPU (tests/test_lanczos_host.py compiles this with g++ and
// compares with numpy.linalg.eigh).  The product runs the same control code over CUDA kernels
// (compartments.cu:DeviceOps).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../tadbit_b200/csrc/lanczos.hpp"

namespace {

struct HostOps {
  const double *A;
  int64_t n;
  std::vector<double> V, w;
  static uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
  }
  void random_w(uint64_t seed) {
    for (int64_t i = 0; i < n; ++i) {
      const uint64_t r = splitmix64(seed * 0x100000001b3ULL + (uint64_t)i);
      w[i] = (double)(r >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
  }
  void matvec(int slot) {
    const double *v = &V[(size_t)slot * n];
    for (int64_t r = 0; r < n; ++r) {
      double acc = 0.0;
      for (int64_t k = 0; k < n; ++k) acc += A[r * n + k] * v[k];
      w[r] = acc;
    }
  }
  std::vector<double> al, be;
  void project(int m) { project_h(m, nullptr); }
  double project_h(int m, double *last) {
    std::vector<double> h(m);
    for (int k = 0; k < m; ++k) {
      double acc = 0.0;
      for (int64_t i = 0; i < n; ++i) acc += V[(size_t)k * n + i] * w[i];
      h[k] = acc;
    }
    for (int k = 0; k < m; ++k)
      for (int64_t i = 0; i < n; ++i) w[i] -= h[k] * V[(size_t)k * n + i];
    if (last) *last = h[m - 1];
    return 0.0;
  }
  void step(int j, bool store_next) {
    matvec(j);
    double h1 = 0.0, h2 = 0.0;
    project_h(j + 1, &h1);
    project_h(j + 1, &h2);
    if ((int)al.size() <= j) { al.resize(j + 1); be.resize(j + 1); }
    al[j] = h1 + h2;
    be[j] = norm_w();
    const double inv = be[j] > 0.0 ? 1.0 / be[j] : 0.0;
    if (store_next) keep(j + 1, inv);
  }
  void read_ab(int j0, int j1, double *a, double *b) {
    for (int j = j0; j < j1; ++j) { a[j - j0] = al[j]; b[j - j0] = be[j]; }
  }
  double norm_w() {
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) acc += w[i] * w[i];
    return sqrt(acc);
  }
  void keep(int slot, double scale) {
    for (int64_t i = 0; i < n; ++i) V[(size_t)slot * n + i] = w[i] * scale;
  }
  void combine(int m, int k, const double *coef, double *out) {
    for (int v = 0; v < k; ++v) {
      double nv = 0.0;
      for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int j = 0; j < m; ++j) acc += coef[(size_t)v * m + j] * V[(size_t)j * n + i];
        out[(size_t)v * n + i] = acc;
        nv += acc * acc;
      }
      nv = sqrt(nv);
      if (nv > 0.0)
        for (int64_t i = 0; i < n; ++i) out[(size_t)v * n + i] /= nv;
    }
  }
};

}  // namespace

extern "C" int lanczos_host(int64_t n, const double *A, int k, int m_cap, double tol, double *evals, double *evecs,
                            int *info) {
  const bool all = k >= n;
  if (all) k = (int)n;
  if (m_cap <= 0 || m_cap > n) m_cap = (int)n;
  HostOps ops;
  ops.A = A;
  ops.n = n;
  ops.V.assign((size_t)m_cap * n, 0.0);
  ops.w.assign(n, 0.0);
  tb::LanczosResult r = tb::lanczos_largest(ops, n, k, m_cap, tol, all, evecs);
  for (size_t i = 0; i < r.evals.size(); ++i) evals[i] = r.evals[i];
  info[0] = (int)r.evals.size();
  info[1] = r.steps;
  info[2] = r.restarts;
  info[3] = r.converged ? 1 : 0;
  info[4] = r.round_trips;
  return 0;
}
