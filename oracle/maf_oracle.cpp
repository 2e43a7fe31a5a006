// =============================================================================
// oracle/maf_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++17, std::complex<double>, complex-step tangent with
// eps_k = 1e-15) of the reference's residual + tangent assembly `calc_r_K`
// and of just enough of its input generation to build meshes without Julia.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may load this library, and only as the checker / the CPU baseline.
//
// Parity status: the reference (Julia) cannot run in this image, and it ships
// no fixtures for residuals / tangents / sparsity / Newton histories.  The
// oracle is pinned against every known answer in the reference's own test/
// directory that touches this path (spline values + derivatives, knot spans,
// unique-element ids, Gauss points, tensor-product table layout, GeoDynStress
// geometry + viscous stresses; see tests/test_oracle_reference_known_answers.py)
// and against an independent numpy formulation (tests/ref_numpy.py).  For the
// assembled r / K themselves: PARITY UNPINNED BY THE REFERENCE (no vectors exist).
//
// Every function cites the reference file:line (relative to /root/reference)
// it restates.  Operation order follows the Julia source.
// =============================================================================
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

// Two kinds of build of this ONE source (oracle/Makefile):
//   libmaf_oracle.so  real_t = double, cd = std::complex<double>: the restated reference algorithm.
//   libmaf_truth.so   -DORC_TRUTH=1 (long double, 64-bit mantissa) / libmaf_truthq.so -DORC_TRUTH=2 (__float128,
//                     113 bits): the very same algorithm, operation for operation, evaluated in extended precision
//                     on the same double inputs and rounded to double at the very end. It measures how far a double
//                     evaluation (this oracle, the GPU kernels) is from the exact value of the reference's formulas.
//                     Every number additionally carries E, the first-order running bound of the rounding error the
//                     REFERENCE ALGORITHM ITSELF commits in double precision, in units of eps (Wilkinson forward
//                     analysis, one rounding per operation):
//                         z = x +- y : E(z) = E(x) + E(y) + |z|          z = x * y : E(z) = |x| E(y) + |y| E(x) + |z|
//                         z = x / y  : E(z) = E(x)/|y| + |x| E(y)/y^2 + |z|      z = sqrt(x) : E(z) = E(x)/(2 z) + |z|
//                     (inputs are exact doubles, E = 0). E is what "sum of |terms|" means once the cancellations
//                     inside the terms are counted too; the parity rule's floor is a multiple of eps * E per entry.
#if defined(ORC_TRUTH)
#if ORC_TRUTH == 2
typedef __float128 real_t;
// sqrt without libquadmath: long-double estimate, two Newton steps (64 -> 128 -> 256 correct bits, rounded to 113)
static inline real_t r_sqrt(real_t x) {
  if (!(x > 0)) return 0;
  real_t y = (real_t)sqrtl((long double)x);
  y = (y + x / y) / 2;
  y = (y + x / y) / 2;
  return y;
}
#else
typedef long double real_t;
static inline real_t r_sqrt(real_t x) { return sqrtl(x); }
#endif
static inline real_t r_abs(real_t x) { return x < 0 ? -x : x; }
struct cd {   // complex number over real_t, plain formulas (what -fcx-limited-range gives std::complex), + error bounds
  real_t re, im, er, ei;
  cd() : re(0), im(0), er(0), ei(0) {}
  cd(real_t r) : re(r), im(0), er(0), ei(0) {}
  cd(double r) : re(r), im(0), er(0), ei(0) {}
  cd(int r) : re(r), im(0), er(0), ei(0) {}
  cd(real_t r, real_t i) : re(r), im(i), er(0), ei(0) {}
  cd(real_t r, real_t i, real_t a, real_t b) : re(r), im(i), er(a), ei(b) {}
  real_t real() const { return re; }
  real_t imag() const { return im; }
  cd& operator+=(const cd& o);
  cd& operator-=(const cd& o);
  cd& operator/=(const cd& o);
};
static inline cd operator+(const cd& a, const cd& b) {
  const real_t r = a.re + b.re, i = a.im + b.im;
  return cd(r, i, a.er + b.er + r_abs(r), a.ei + b.ei + r_abs(i));
}
static inline cd operator-(const cd& a, const cd& b) {
  const real_t r = a.re - b.re, i = a.im - b.im;
  return cd(r, i, a.er + b.er + r_abs(r), a.ei + b.ei + r_abs(i));
}
static inline cd operator-(const cd& a) { return cd(-a.re, -a.im, a.er, a.ei); }
static inline cd operator*(const cd& a, const cd& b) {
  const real_t ar = r_abs(a.re), ai = r_abs(a.im), br = r_abs(b.re), bi = r_abs(b.im);
  const real_t r = a.re * b.re - a.im * b.im, i = a.re * b.im + a.im * b.re;
  // each product: |x| E(y) + |y| E(x) + |x y|; then the sum / difference
  const real_t e_rr = ar * b.er + br * a.er + ar * br, e_ii = ai * b.ei + bi * a.ei + ai * bi;
  const real_t e_ri = ar * b.ei + bi * a.er + ar * bi, e_ir = ai * b.er + br * a.ei + ai * br;
  return cd(r, i, e_rr + e_ii + r_abs(r), e_ri + e_ir + r_abs(i));
}
static inline cd operator/(const cd& a, const cd& b) {
  // plain formula (a conj(b)) / |b|^2, analysed as it is evaluated
  const real_t d = b.re * b.re + b.im * b.im;
  const real_t br = r_abs(b.re), bi = r_abs(b.im);
  const real_t e_d = 2 * br * b.er + br * br + 2 * bi * b.ei + bi * bi + d;
  const cd num = a * cd(b.re, -b.im, b.er, b.ei);
  const real_t r = num.re / d, i = num.im / d;
  return cd(r, i, num.er / d + r_abs(num.re) * e_d / (d * d) + r_abs(r), num.ei / d + r_abs(num.im) * e_d / (d * d) + r_abs(i));
}
inline cd& cd::operator+=(const cd& o) { *this = *this + o; return *this; }
inline cd& cd::operator-=(const cd& o) { *this = *this - o; return *this; }
inline cd& cd::operator/=(const cd& o) { *this = *this / o; return *this; }
#define ORC_MIXED(T)                                                                   \
  static inline cd operator+(const cd& a, T b) { return a + cd((real_t)b); }          \
  static inline cd operator+(T a, const cd& b) { return cd((real_t)a) + b; }          \
  static inline cd operator-(const cd& a, T b) { return a - cd((real_t)b); }          \
  static inline cd operator-(T a, const cd& b) { return cd((real_t)a) - b; }          \
  static inline cd operator*(const cd& a, T b) { return a * cd((real_t)b); }          \
  static inline cd operator*(T a, const cd& b) { return cd((real_t)a) * b; }          \
  static inline cd operator/(const cd& a, T b) { return a / cd((real_t)b); }          \
  static inline cd operator/(T a, const cd& b) { return cd((real_t)a) / b; }
ORC_MIXED(double)
ORC_MIXED(int)
static inline cd c_conj(const cd& a) { return cd(a.re, -a.im, a.er, a.ei); }
// principal square root. The path only takes roots of numbers with a positive real part and an O(eps_k) imaginary
// part (metric determinants, squared lengths): sqrt(z) = (t, im / 2t), t = sqrt((|z| + re) / 2).
static inline cd c_sqrt(const cd& z) {
  const real_t m = r_sqrt(z.re * z.re + z.im * z.im);
  if (z.re >= 0) {
    const real_t t = r_sqrt((m + z.re) / 2);
    if (t == 0) return cd(0.0);
    const real_t i = z.im / (2 * t);
    // first order in the (tiny) imaginary part: t ~ sqrt(re), im' ~ im / (2 sqrt(re))
    const real_t e_t = z.er / (2 * t) + 3 * t;
    return cd(t, i, e_t, z.ei / (2 * t) + r_abs(z.im) * e_t / (2 * t * t) + r_abs(i));
  }
  const real_t t = r_sqrt((m - z.re) / 2);
  return cd(r_abs(z.im) / (2 * t), z.im < 0 ? -t : t, z.er / (2 * t) + 3 * t, z.er / (2 * t) + 3 * t);
}
static inline real_t c_er(const cd& z) { return z.er; }
static inline real_t c_ei(const cd& z) { return z.ei; }
#else
typedef double real_t;
typedef std::complex<double> cd;
static inline real_t r_abs(real_t x) { return std::fabs(x); }
static inline cd c_conj(const cd& a) { return std::conj(a); }
static inline cd c_sqrt(const cd& z) { return std::sqrt(z); }
static inline real_t c_er(const cd&) { return 0.0; }
static inline real_t c_ei(const cd&) { return 0.0; }
#endif

// --- enums: src/input/Enums.jl:32-156, src/input/Dof.jl:28-33 -----------------
enum Scenario { F_CAVI = 1, F_COUE = 2, F_POIS = 3, F_PULL = 4, F_BEND = 5 };
enum Motion { STATIC = 1, EUL = 2, LAG = 3, ALEV = 4, ALEVB = 5 };
enum Boundary { BOTTOM = 1, RIGHT = 2, TOP = 3, LEFT = 4 };
enum Neumann { SHEAR = 1, STRETCH = 2, MOMENT = 3 };
enum Curve { CLAMPED = 1, CLOSED = 2 };
enum Unknown { U_vx = 1, U_vy, U_vz, U_vmx, U_vmy, U_vmz, U_lam, U_pm };  // Dof.jl

// compile-time constants: src/input/Params.jl:161-179
static const int POLY = 2, GP1D = 3, NDERS = 2, ZDIM = 2, XDIM = 3, NEN = 9, VOIGT = 3, NEDBDF = 3;

#define ORC_CHECK(cond, msg)                                   \
  do {                                                         \
    if (!(cond)) throw std::runtime_error(std::string(msg));   \
  } while (0)

// =============================================================================
// Spline.jl
// =============================================================================
struct KnotVector {  // Spline.jl:20-33  (zs is 0-based storage of the 1-based Julia vector)
  std::vector<double> zs;
  int nel = 0, poly = 0, curve = CLAMPED;
};

// Julia isapprox for Float64 scalars: rtol = sqrt(eps), atol = 0.
static bool isapprox(double x, double y) {
  if (x == y) return true;
  if (!std::isfinite(x) || !std::isfinite(y)) return false;
  const double rtol = std::sqrt(2.220446049250313e-16);
  return std::fabs(x - y) <= rtol * std::max(std::fabs(x), std::fabs(y));
}

// Spline.jl:46-61
static KnotVector knot_vector_from_list(const std::vector<double>& zs, int poly, int curve) {
  const int n = (int)zs.size();
  if (curve == CLAMPED) {
    for (int i = 0; i + 1 < n; ++i) ORC_CHECK(zs[i] <= zs[i + 1], "knots must be non-decreasing");
    ORC_CHECK(n >= 2 * poly + 2, "knot vector too short");
    for (int i = 0; i <= poly; ++i) ORC_CHECK(zs[i] == zs[0], "need poly+1 repeated knots at start");
    for (int i = n - poly - 1; i < n; ++i) ORC_CHECK(zs[i] == zs[n - 1], "need poly+1 repeated knots at end");
  } else if (curve == CLOSED) {
    // zs[2:end]-zs[1:end-1] ≈ (zs[2]-zs[1])*ones  (vector isapprox: norm based)
    double d0 = zs[1] - zs[0], num = 0, den = 0;
    for (int i = 0; i + 1 < n; ++i) {
      double d = zs[i + 1] - zs[i];
      num += (d - d0) * (d - d0);
      den = std::max(den, 0.0);
    }
    double na = 0, nb = 0;
    for (int i = 0; i + 1 < n; ++i) {
      double d = zs[i + 1] - zs[i];
      na += d * d;
      nb += d0 * d0;
    }
    const double rtol = std::sqrt(2.220446049250313e-16);
    ORC_CHECK(std::sqrt(num) <= rtol * std::max(std::sqrt(na), std::sqrt(nb)), "closed knots must be uniform");
  } else {
    ORC_CHECK(false, "knot vector for curve not implemented");
  }
  KnotVector kv;
  kv.zs = zs;
  kv.poly = poly;
  kv.curve = curve;
  kv.nel = n - 2 * poly - 1;
  return kv;
}

// Spline.jl:76-96
static KnotVector knot_vector_uniform(int nel, int poly, int curve) {
  const int num = nel + 2 * poly + 1;
  std::vector<double> zs(num, 0.0);
  if (curve == CLAMPED) {
    for (int idx = poly + 2; idx <= num - poly - 1; ++idx) zs[idx - 1] = (double)(idx - poly - 1) / (double)nel;
    for (int idx = num - poly; idx <= num; ++idx) zs[idx - 1] = 1.0;
  } else if (curve == CLOSED) {
    for (int idx = 1; idx <= num; ++idx) zs[idx - 1] = (double)(idx - poly - 1) / (double)nel;
  }
  return knot_vector_from_list(zs, poly, curve);
}

// Spline.jl:122-183
static std::vector<double> get_fine_zs(int nel, int poly) {
  ORC_CHECK(nel >= 18, "fine mesh requires at least 18 1-D elements");
  const int num = nel + 2 * poly + 1;
  std::vector<double> z(num + 1, 0.0);  // 1-based
  const int num_wide_l1 = 6;
  const int num_fine_l1 = nel - 2 * num_wide_l1 - 1;
  const double z_wide_l1 = 1.0 / 3.0;
  const double z_fine_l1 = 1.0 - 2 * z_wide_l1;
  for (int idx = poly + 2; idx <= poly + num_wide_l1 + 1; ++idx) {
    double dz = z_wide_l1 / num_wide_l1;
    z[idx] = (idx - poly - 1) * dz;
  }
  for (int idx = num - poly - num_wide_l1; idx <= num - poly - 1; ++idx) {
    double dz = z_wide_l1 / num_wide_l1;
    z[idx] = 1.0 - z_wide_l1 + (idx - num + poly + num_wide_l1) * dz;
  }
  const int num_wide_l2 = num_fine_l1 / 4;  // floor
  const int num_fine_l2 = num_fine_l1 - 2 * num_wide_l2;
  const double z_wide_l2 = 1.0 / 9.0;
  const double z_fine_l2 = z_fine_l1 - 2 * z_wide_l2;
  if (2 * num_wide_l2 <= num_wide_l1) {
    double dz = z_fine_l1 / (num_fine_l1 + 1);
    for (int idx = poly + num_wide_l1 + 2; idx <= num - poly - num_wide_l1 - 1; ++idx)
      z[idx] = z_wide_l1 + (idx - poly - num_wide_l1 - 1) * dz;
  } else {
    double dz = z_wide_l2 / num_wide_l2;
    for (int idx = num_wide_l1 + poly + 2; idx <= num_wide_l1 + poly + 1 + num_wide_l2; ++idx)
      z[idx] = z_wide_l1 + (idx - num_wide_l1 - poly - 1) * dz;
    for (int idx = num - poly - num_wide_l1 - num_wide_l2; idx <= num - poly - num_wide_l1 - 1; ++idx)
      z[idx] = 1.0 - z_wide_l1 - z_wide_l2 + (idx - num + poly + num_wide_l1 + num_wide_l2) * dz;
    dz = z_fine_l2 / (num_fine_l2 + 1);
    for (int idx = num_wide_l1 + poly + 2 + num_wide_l2; idx <= num - poly - num_wide_l1 - num_wide_l2 - 1; ++idx)
      z[idx] = z_wide_l1 + z_wide_l2 + (idx - num_wide_l1 - poly - 1 - num_wide_l2) * dz;
  }
  for (int idx = num - poly; idx <= num; ++idx) z[idx] = 1.0;
  return std::vector<double>(z.begin() + 1, z.end());
}

// Spline.jl:198-248 (NURBS book A2.1, extended to CLOSED). Returns the 1-based index.
static int get_knot_span_index(const KnotVector& kv, double zeta) {
  std::vector<double> z(kv.zs.size() + 1);  // 1-based copy
  for (size_t i = 0; i < kv.zs.size(); ++i) z[i + 1] = kv.zs[i];
  const int nk = (int)kv.zs.size();
  if (kv.curve == CLOSED) {
    for (int i = 1; i <= kv.poly; ++i) z[i] = z[kv.poly + 1];
    for (int i = nk - kv.poly + 1; i <= nk; ++i) z[i] = z[nk - kv.poly];
  }
  ORC_CHECK(zeta >= z[1], "zeta smaller than first active knot");
  ORC_CHECK(zeta <= z[nk], "zeta larger than last active knot");
  int m = 1;
  while (z[m] == z[1]) ++m;
  int n = nk;
  while (z[n] == z[nk]) --n;
  if (zeta == z[n + 1]) return n;
  int low = m - 1, high = n + 1;
  int mid = (low + high) / 2;
  while (zeta < z[mid] || zeta >= z[mid + 1]) {
    if (zeta < z[mid]) high = mid; else low = mid;
    mid = (low + high) / 2;
  }
  return mid;
}

// Spline.jl:264-296 (A2.2)
static void get_bspline_vals(const KnotVector& kv, double zeta, double* out /*poly+1*/) {
  ORC_CHECK(zeta >= kv.zs.front(), "zeta is less than smallest knot");
  ORC_CHECK(zeta <= kv.zs.back(), "zeta is greater than largest knot");
  const int p = kv.poly;
  std::vector<double> left(p + 2, 0.0), right(p + 2, 0.0), bf(p + 2, 0.0);  // 1-based
  const int span = get_knot_span_index(kv, zeta);
  auto Z = [&](int i) { return kv.zs[i - 1]; };
  bf[1] = 1.0;
  for (int j = 1; j <= p; ++j) {
    left[j + 1] = zeta - Z(span + 1 - j);
    right[j + 1] = Z(span + j) - zeta;
    double saved = 0.0;
    for (int r = 1; r <= j; ++r) {
      double temp = bf[r] / (right[r + 1] + left[j + 2 - r]);
      bf[r] = saved + right[r + 1] * temp;
      saved = left[j + 2 - r] * temp;
    }
    bf[j + 1] = saved;
  }
  for (int j = 0; j <= p; ++j) out[j] = bf[j + 1];
}

// Spline.jl:319-422 (A2.3). out is (poly+1) x (num_ders+1), row-major out[r*(nd+1)+k].
static void get_bspline_ders(const KnotVector& kv, double zeta, int num_ders, double* out) {
  ORC_CHECK(zeta >= kv.zs.front(), "zeta is less than smallest knot");
  ORC_CHECK(zeta <= kv.zs.back(), "zeta is greater than largest knot");
  const int p = kv.poly;
  ORC_CHECK(num_ders >= 0, "cannot have fewer than zero derivatives");
  ORC_CHECK(num_ders <= p, "basis functions have only `poly` derivatives");
  std::vector<double> left(p + 2, 0.0), right(p + 2, 0.0);
  const int span = get_knot_span_index(kv, zeta);
  auto Z = [&](int i) { return kv.zs[i - 1]; };
  const int P1 = p + 2;  // 1-based square storage
  std::vector<double> ndu(P1 * P1, 0.0), a(3 * P1, 0.0), ders(P1 * (num_ders + 2), 0.0);
  auto NDU = [&](int i, int j) -> double& { return ndu[i * P1 + j]; };
  auto A = [&](int i, int j) -> double& { return a[i * P1 + j]; };
  auto D = [&](int i, int k) -> double& { return ders[i * (num_ders + 2) + k]; };
  NDU(1, 1) = 1.0;
  for (int j = 1; j <= p; ++j) {
    left[j + 1] = zeta - Z(span + 1 - j);
    right[j + 1] = Z(span + j) - zeta;
    double saved = 0.0;
    for (int r = 1; r <= j; ++r) {
      NDU(j + 1, r) = right[r + 1] + left[j + 2 - r];
      double temp = NDU(r, j) / NDU(j + 1, r);
      NDU(r, j + 1) = saved + right[r + 1] * temp;
      saved = left[j + 2 - r] * temp;
    }
    NDU(j + 1, j + 1) = saved;
  }
  for (int j = 1; j <= p + 1; ++j) D(j, 1) = NDU(j, p + 1);
  for (int r = 1; r <= p + 1; ++r) {
    int s1 = 1, s2 = 2;
    A(1, 1) = 1.0;
    for (int k = 1; k <= num_ders; ++k) {
      double d = 0.0;
      int rk = r - k, pk = p - k;
      if (r > k) {
        A(s2, 1) = A(s1, 1) / NDU(pk + 2, rk);
        d = A(s2, 1) * NDU(rk, pk + 1);
      }
      int j1 = rk >= 0 ? 1 : -rk + 1;
      int j2 = (r - 2 <= pk) ? k - 1 : p - r + 1;
      for (int j = j1; j <= j2; ++j) {
        A(s2, j + 1) = (A(s1, j + 1) - A(s1, j)) / NDU(pk + 2, rk + j);
        d += A(s2, j + 1) * NDU(rk + j, pk + 1);
      }
      if (r - 1 <= pk) {
        A(s2, k + 1) = -A(s1, k) / NDU(pk + 2, r);
        d += A(s2, k + 1) * NDU(r, pk + 1);
      }
      D(r, k + 1) = d;
      std::swap(s1, s2);
    }
  }
  int r = p;
  for (int k = 1; k <= num_ders; ++k) {
    for (int j = 1; j <= p + 1; ++j) D(j, k + 1) *= r;
    r *= (p - k);
  }
  for (int j = 1; j <= p + 1; ++j)
    for (int k = 1; k <= num_ders + 1; ++k) out[(j - 1) * (num_ders + 1) + (k - 1)] = D(j, k);
}

// Spline.jl:465-481 (ks_id = knot-span id, 1-based)
static void get_bspline_indices_ks(const KnotVector& kv, int ks_id, int* ids /*poly+1, 1-based values*/) {
  const int p = kv.poly;
  for (int i = 1; i <= p + 1; ++i) {
    int id = i + (ks_id - p - 1);
    if (kv.curve == CLOSED) id = (id - 1) % kv.nel + 1;
    ids[i - 1] = id;
  }
}
// Spline.jl:441-449
static void get_bspline_indices(const KnotVector& kv, double zeta, int* ids) {
  get_bspline_indices_ks(kv, get_knot_span_index(kv, zeta), ids);
}

// Spline.jl:580-626
static std::vector<double> collocate_zeta(const KnotVector& kv) {
  const int p = kv.poly;
  const int nk = (int)kv.zs.size();
  if (kv.curve == CLOSED) {
    std::vector<double> out;
    for (int i = p + 1; i <= nk - p - 1; ++i) out.push_back((kv.zs[i - 1] + kv.zs[i]) / 2);
    return out;
  }
  std::vector<double> u;  // unique(kv.zs)
  for (double z : kv.zs)
    if (std::find(u.begin(), u.end(), z) == u.end()) u.push_back(z);
  ORC_CHECK(p > 1, "interpolation for poly >= 2 only when clamped");
  ORC_CHECK(p < 4, "interpolation for poly <= 3 only when clamped");
  ORC_CHECK((int)u.size() == nk - 2 * p, "no repeated interior knots");
  const int nb = nk - p - 1;
  std::vector<double> zl(nb, 0.0);
  zl[0] = u.front();
  zl[nb - 1] = u.back();
  if (p == 2) {
    for (int i = 2; i <= nb - 1; ++i) zl[i - 1] = (u[i - 2] + u[i - 1]) / 2;
  } else {
    zl[1] = (u[0] + u[1]) / 2;
    zl[nb - 2] = (u[u.size() - 2] + u[u.size() - 1]) / 2;
    for (int i = 3; i <= nb - 2; ++i) zl[i - 1] = u[i - 2];
  }
  return zl;
}

// dense solve, Gaussian elimination with partial pivoting (stand-in for Julia `\` = LAPACK LU)
static std::vector<double> dense_solve(std::vector<double> Amat, std::vector<double> b, int n) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(Amat[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(Amat[(size_t)i * n + k]) > best) best = std::fabs(Amat[(size_t)i * n + k]), piv = i;
    ORC_CHECK(best > 0.0, "singular collocation matrix");
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(Amat[(size_t)k * n + j], Amat[(size_t)piv * n + j]);
      std::swap(b[k], b[piv]);
    }
    for (int i = k + 1; i < n; ++i) {
      double f = Amat[(size_t)i * n + k] / Amat[(size_t)k * n + k];
      if (f == 0.0) continue;
      for (int j = k; j < n; ++j) Amat[(size_t)i * n + j] -= f * Amat[(size_t)k * n + j];
      b[i] -= f * b[k];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= Amat[(size_t)i * n + j] * b[j];
    b[i] = s / Amat[(size_t)i * n + i];
  }
  return b;
}

// Spline.jl:510-528; xvals = x.(collocate_zeta(kv)) supplied by the caller
static std::vector<double> get_1d_bspline_cps(const KnotVector& kv, const std::vector<double>& xvals) {
  std::vector<double> zl = collocate_zeta(kv);
  const int nb = (int)zl.size();
  ORC_CHECK((int)xvals.size() == nb, "xvals length mismatch");
  std::vector<double> M((size_t)nb * nb, 0.0);
  std::vector<int> ids(kv.poly + 1);
  std::vector<double> vals(kv.poly + 1);
  for (int j = 0; j < nb; ++j) {
    get_bspline_indices(kv, zl[j], ids.data());
    get_bspline_vals(kv, zl[j], vals.data());
    for (int k = 0; k <= kv.poly; ++k) M[(size_t)j * nb + (ids[k] - 1)] = vals[k];
  }
  return dense_solve(M, xvals, nb);
}

// Spline.jl:540-567; xvals[j + k*num1] = x(z1list[j], z2list[k])
static std::vector<double> get_2d_bspline_cps(const KnotVector& kv1, const KnotVector& kv2,
                                              const std::vector<double>& xvals) {
  std::vector<double> z1 = collocate_zeta(kv1), z2 = collocate_zeta(kv2);
  const int n1 = (int)z1.size(), n2 = (int)z2.size(), nb = n1 * n2;
  ORC_CHECK((int)xvals.size() == nb, "xvals length mismatch");
  ORC_CHECK(nb <= 4096, "dense 2-D collocation restricted to small meshes in the oracle");
  std::vector<double> M((size_t)nb * nb, 0.0);
  std::vector<int> i1(kv1.poly + 1), i2(kv2.poly + 1);
  std::vector<double> v1(kv1.poly + 1), v2(kv2.poly + 1);
  for (int k = 0; k < n2; ++k)
    for (int j = 0; j < n1; ++j) {
      get_bspline_indices(kv1, z1[j], i1.data());
      get_bspline_indices(kv2, z2[k], i2.data());
      get_bspline_vals(kv1, z1[j], v1.data());
      get_bspline_vals(kv2, z2[k], v2.data());
      for (int q = 0; q <= kv2.poly; ++q)
        for (int p = 0; p <= kv1.poly; ++p)
          M[(size_t)(j + k * n1) * nb + ((i1[p] - 1) + (i2[q] - 1) * n1)] = v1[p] * v2[q];
    }
  return dense_solve(M, xvals, nb);
}

// Spline.jl:640-671
struct Unique1D {
  int uel_num = 0, num_el = 0;
  std::vector<int> uel_ids;                          // 1-based values
  std::vector<std::pair<double, double>> uel_list;   // (zlo, zhi)
};
static Unique1D get_unique_1d_elements(const KnotVector& kv) {
  const int p = kv.poly, nk = (int)kv.zs.size();
  auto Z = [&](int i) { return kv.zs[i - 1]; };
  std::vector<double> prev(2 * p + 1, 0.0), ctx(2 * p + 1);
  Unique1D u;
  u.num_el = nk - 2 * p - 1;
  u.uel_ids.assign(u.num_el, 0);
  for (int k = p + 1; k <= nk - p - 1; ++k) {
    int el = k - p;
    for (int q = 0; q < 2 * p + 1; ++q) ctx[q] = Z(k - p + 1 + q) - Z(k - p + q);
    bool same = true;
    for (int q = 0; q < 2 * p + 1; ++q) same = same && isapprox(ctx[q], prev[q]);
    if (!same) {
      u.uel_num += 1;
      u.uel_list.push_back({Z(k), Z(k + 1)});
    }
    u.uel_ids[el - 1] = u.uel_num;
    prev = ctx;
  }
  return u;
}

// =============================================================================
// GaussPoint.jl
// =============================================================================
// GaussPoint.jl:106-133, 143-171
static void gauss_xi(int ngp, double* xs, double* ws) {
  ORC_CHECK(ngp <= 4, "have at most 4 1-D Gauss points");
  if (ngp == 1) { xs[0] = -1.0; ws[0] = 0.0; }
  else if (ngp == 3) {
    xs[0] = -std::sqrt(3.0 / 5.0); xs[1] = 0.0; xs[2] = std::sqrt(3.0 / 5.0);
    ws[0] = 5.0 / 9.0; ws[1] = 8.0 / 9.0; ws[2] = 5.0 / 9.0;
  } else if (ngp == 4) {
    xs[0] = -std::sqrt(3.0 / 7.0 + std::sqrt(6.0 / 5.0) * 2 / 7);
    xs[1] = -std::sqrt(3.0 / 7.0 - std::sqrt(6.0 / 5.0) * 2 / 7);
    xs[2] = std::sqrt(3.0 / 7.0 - std::sqrt(6.0 / 5.0) * 2 / 7);
    xs[3] = std::sqrt(3.0 / 7.0 + std::sqrt(6.0 / 5.0) * 2 / 7);
    ws[0] = (18 - std::sqrt(30.0)) / 36; ws[1] = (18 + std::sqrt(30.0)) / 36;
    ws[2] = (18 + std::sqrt(30.0)) / 36; ws[3] = (18 - std::sqrt(30.0)) / 36;
  } else ORC_CHECK(false, "Gauss points not implemented");
}
// GaussPoint.jl:72-76
static void gauss_zeta(int ngp, double lo, double hi, double* zs, double* ws) {
  double xs[4], wx[4];
  gauss_xi(ngp, xs, wx);
  for (int k = 0; k < ngp; ++k) {
    zs[k] = xs[k] * (hi - lo) / 2 + (hi + lo) / 2;
    ws[k] = wx[k] * (hi - lo) / 2;
  }
}

// =============================================================================
// GpBasisFn.jl
// =============================================================================
struct Fn1 { double w, N[3], dN[3], ddN[3]; };               // GpBasisFn.jl:20-34
struct Fn2 { double w, N[9], dN[9][2], ddN[9][3]; };         // GpBasisFn.jl:96-112

// GpBasisFn.jl:45-55
static Fn1 gp_basis_fns_1d(double w, double zeta, const KnotVector& kv) {
  ORC_CHECK(kv.poly == POLY, "oracle tables are compiled for POLY = 2");
  double d[3 * 3];
  get_bspline_ders(kv, zeta, NDERS, d);
  Fn1 f;
  f.w = w;
  for (int i = 0; i < 3; ++i) { f.N[i] = d[i * 3 + 0]; f.dN[i] = d[i * 3 + 1]; f.ddN[i] = d[i * 3 + 2]; }
  return f;
}
// GpBasisFn.jl:102-110 (index id1 + 3*(id2-1); second derivative columns 11, 22, 12)
static Fn2 gp_basis_fns_2d(const Fn1& f1, const Fn1& f2) {
  Fn2 g;
  g.w = f1.w * f2.w;
  for (int i2 = 0; i2 < 3; ++i2)
    for (int i1 = 0; i1 < 3; ++i1) {
      int a = i1 + 3 * i2;
      g.N[a] = f1.N[i1] * f2.N[i2];
      g.dN[a][0] = f1.dN[i1] * f2.N[i2];
      g.dN[a][1] = f1.N[i1] * f2.dN[i2];
      g.ddN[a][0] = f1.ddN[i1] * f2.N[i2];
      g.ddN[a][1] = f1.N[i1] * f2.ddN[i2];
      g.ddN[a][2] = f1.dN[i1] * f2.dN[i2];
    }
  return g;
}

struct LineFns {  // GpBasisFn.jl:143-202
  int nel = 0, nuel = 0, ngp = 0;
  std::vector<int> uel_ids;   // 1-based values
  std::vector<Fn1> ufns;      // [(uel-1)*ngp + gp-1]
  Fn1 zmin, zmax;
};
static LineFns line_gp_basis_fns(const KnotVector& kv, int ngp) {
  Unique1D u = get_unique_1d_elements(kv);
  LineFns L;
  L.nel = u.num_el; L.nuel = u.uel_num; L.ngp = ngp; L.uel_ids = u.uel_ids;
  L.ufns.resize((size_t)L.nuel * ngp);
  for (int ue = 0; ue < L.nuel; ++ue) {
    double zs[4], ws[4];
    gauss_zeta(ngp, u.uel_list[ue].first, u.uel_list[ue].second, zs, ws);
    for (int g = 0; g < ngp; ++g) L.ufns[(size_t)ue * ngp + g] = gp_basis_fns_1d(ws[g], zs[g], kv);
  }
  const int nk = (int)kv.zs.size();
  if (kv.curve == CLAMPED) {
    L.zmin = gp_basis_fns_1d(1.0, kv.zs[0], kv);
    L.zmax = gp_basis_fns_1d(1.0, kv.zs[nk - 1], kv);
  } else {
    L.zmin = gp_basis_fns_1d(1.0, kv.zs[kv.poly], kv);
    L.zmax = gp_basis_fns_1d(1.0, kv.zs[nk - kv.poly - 1], kv);
  }
  return L;
}

struct BdryFns {  // GpBasisFn.jl:231-277
  int nel = 0, nuel = 0;
  std::vector<int> uel_ids;
  std::vector<Fn2> ufns;  // [(uel-1)*GP1D + gp-1]
};
static BdryFns bdry_gp_basis_fns(const LineFns& line, const Fn1& perp, int bdry) {
  BdryFns B;
  B.nel = line.nel; B.nuel = line.nuel; B.uel_ids = line.uel_ids;
  B.ufns.resize((size_t)line.nuel * line.ngp);
  for (int id = 0; id < line.nuel; ++id)
    for (int g = 0; g < line.ngp; ++g) {
      const Fn1& f = line.ufns[(size_t)id * line.ngp + g];
      B.ufns[(size_t)id * line.ngp + g] =
          (bdry == BOTTOM || bdry == TOP) ? gp_basis_fns_2d(f, perp) : gp_basis_fns_2d(perp, f);
    }
  return B;
}

struct AreaFns {  // GpBasisFn.jl:300-356
  int nel = 0, nuel = 0;
  std::vector<int> uel_ids;   // 1-based values
  std::vector<Fn2> ufns;      // [(uel-1)*9 + gp-1]
};
static AreaFns area_gp_basis_fns(const LineFns& l1, const LineFns& l2) {
  AreaFns A;
  A.nel = l1.nel * l2.nel;
  A.nuel = l1.nuel * l2.nuel;
  A.uel_ids.assign(A.nel, 0);
  for (int id1 = 1; id1 <= l1.nel; ++id1)
    for (int id2 = 1; id2 <= l2.nel; ++id2)
      A.uel_ids[id1 + (id2 - 1) * l1.nel - 1] = l1.uel_ids[id1 - 1] + (l2.uel_ids[id2 - 1] - 1) * l1.nuel;
  A.ufns.resize((size_t)A.nuel * l1.ngp * l2.ngp);
  for (int id1 = 1; id1 <= l1.nuel; ++id1)
    for (int id2 = 1; id2 <= l2.nuel; ++id2)
      for (int g1 = 1; g1 <= l1.ngp; ++g1)
        for (int g2 = 1; g2 <= l2.ngp; ++g2) {
          int ue = id1 + (id2 - 1) * l1.nuel, gp = g1 + (g2 - 1) * l1.ngp;
          A.ufns[(size_t)(ue - 1) * (l1.ngp * l2.ngp) + (gp - 1)] = gp_basis_fns_2d(
              l1.ufns[(size_t)(id1 - 1) * l1.ngp + g1 - 1], l2.ufns[(size_t)(id2 - 1) * l2.ngp + g2 - 1]);
        }
  return A;
}

// =============================================================================
// Params.jl / Mesh.jl / Bc.jl
// =============================================================================
struct Params {  // Params.jl:36-55 (+ the keyword args the hot path reads)
  int motion = ALEVB, scenario = F_PULL, num1el = 17, num2el = 17;
  double length = 64.0, kb = 1.0, kg = -0.5, zv = 1.0, pn = 0.0;
  double adb = 4096.0, am = 1.0, ek = 1e-15, enr = 1e-12;
  double pull_speed = 0.0, bend_mf = 0.0, bend_tm = 1.0;
};

struct NeuBc { int bdry, type; double val; };
struct DirBc { int unknown, node; double val; };

struct Mesh {  // Mesh.jl:48-80
  Params p;
  int num1el, num2el, numel, num1np, num2np, numnp;
  std::vector<int64_t> IX;  // 9 x numel column-major, 1-based node ids
  std::vector<int64_t> bdry_elems[5], bdry_nodes[5], bdry_inner_nodes[5];  // index by Boundary enum
  int64_t crnr_nodes[5];    // BOTTOM_LEFT=1, BOTTOM_RIGHT=2, TOP_LEFT=3, TOP_RIGHT=4
  KnotVector kv1, kv2;
  LineFns line1, line2;
  AreaFns area;
  BdryFns bdry[5];
  int dofs[9];              // dofs[Unknown] = column (1-based) or 0
  int ndf;
  std::vector<int64_t> ID;  // ndf x numnp column-major
  std::vector<int64_t> ID_inv_node, ID_inv_dof;
  int64_t nmdf;
  std::vector<int64_t> LM;  // (9*ndf) x numel column-major
  std::vector<DirBc> inh_dir;
  std::vector<NeuBc> inh_neu;
};

// Bc.jl:414-431
static void get_dofs(int motion, int* dofs, int* ndf) {
  for (int i = 0; i < 9; ++i) dofs[i] = 0;
  if (motion == LAG || motion == STATIC) {
    dofs[U_vx] = 1; dofs[U_vy] = 2; dofs[U_vz] = 3; dofs[U_lam] = 4; *ndf = 4;
  } else if (motion == EUL) {
    dofs[U_vx] = 1; dofs[U_vy] = 2; dofs[U_vz] = 3; dofs[U_vmx] = 4; dofs[U_vmy] = 5; dofs[U_vmz] = 6;
    dofs[U_lam] = 7; *ndf = 7;
  } else if (motion == ALEV || motion == ALEVB) {
    dofs[U_vx] = 1; dofs[U_vy] = 2; dofs[U_vz] = 3; dofs[U_vmx] = 4; dofs[U_vmy] = 5; dofs[U_vmz] = 6;
    dofs[U_lam] = 7; dofs[U_pm] = 8; *ndf = 8;
  } else ORC_CHECK(false, "motion degrees of freedom not provided");
}

// PullForce.jl:8-13
static int64_t get_pull_el_id(int64_t numel) { return (numel + 1) / 2; }  // ceil(numel/2)

// Mesh.jl:574-593
static std::vector<int64_t> construct_IX(const KnotVector& kv1, const KnotVector& kv2, int num1np) {
  std::vector<int64_t> IX((size_t)NEN * kv1.nel * kv2.nel, 0);
  int s1[3], s2[3];
  for (int e2 = 1; e2 <= kv2.nel; ++e2)
    for (int e1 = 1; e1 <= kv1.nel; ++e1) {
      get_bspline_indices_ks(kv2, e2 + POLY, s2);
      get_bspline_indices_ks(kv1, e1 + POLY, s1);
      for (int n2 = 1; n2 <= POLY + 1; ++n2)
        for (int n1 = 1; n1 <= POLY + 1; ++n1)
          IX[(size_t)(n1 + (n2 - 1) * (POLY + 1) - 1) + (size_t)NEN * (e1 + (e2 - 1) * kv1.nel - 1)] =
              s1[n1 - 1] + (int64_t)num1np * (s2[n2 - 1] - 1);
    }
  return IX;
}

#define IDX(m, dof, node) (m).ID[(size_t)((dof)-1) + (size_t)(m).ndf * ((node)-1)]

// Bc.jl:58-106
static void bc_f_cavi(Mesh& m) {
  for (int i = 0; i < 9; ++i) m.dofs[i] = 0;
  m.dofs[U_vx] = 1; m.dofs[U_vy] = 2; m.dofs[U_lam] = 3; m.ndf = 3;
  m.ID.assign((size_t)m.ndf * m.numnp, 0);
  for (int64_t i : m.bdry_nodes[BOTTOM]) { IDX(m, 1, i) = -1; IDX(m, 2, i) = -1; }
  for (int64_t i : m.bdry_nodes[TOP]) {
    IDX(m, 1, i) = -1; IDX(m, 2, i) = -1;
    if (i != m.crnr_nodes[3] && i != m.crnr_nodes[4]) m.inh_dir.push_back({U_vx, (int)i, 1.0});
  }
  for (int64_t i : m.bdry_nodes[LEFT]) { IDX(m, 1, i) = -1; IDX(m, 2, i) = -1; }
  for (int64_t i : m.bdry_nodes[RIGHT]) { IDX(m, 1, i) = -1; IDX(m, 2, i) = -1; }
  int64_t c = m.numnp / 2 + 1;
  IDX(m, 3, c) = -1;
  m.inh_dir.push_back({U_lam, (int)c, 0.0});
}
// Bc.jl:122-157
static void bc_f_coue(Mesh& m) {
  for (int i = 0; i < 9; ++i) m.dofs[i] = 0;
  m.dofs[U_vx] = 1; m.dofs[U_vy] = 2; m.dofs[U_lam] = 3; m.ndf = 3;
  m.ID.assign((size_t)m.ndf * m.numnp, 0);
  for (int64_t i : m.bdry_nodes[BOTTOM]) { IDX(m, 1, i) = -1; IDX(m, 2, i) = -1; }
  for (int64_t i : m.bdry_nodes[TOP]) { IDX(m, 1, i) = -1; IDX(m, 2, i) = -1; m.inh_dir.push_back({U_vx, (int)i, 3.0}); }
  for (int64_t i : m.bdry_nodes[LEFT]) IDX(m, 2, i) = -1;
  for (int64_t i : m.bdry_nodes[RIGHT]) IDX(m, 2, i) = -1;
  m.inh_neu = {{LEFT, STRETCH, 4.0}, {RIGHT, STRETCH, 4.0}};
}
// Bc.jl:286-316
static void bc_f_pois(Mesh& m) {
  for (int i = 0; i < 9; ++i) m.dofs[i] = 0;
  m.dofs[U_vx] = 1; m.dofs[U_vy] = 2; m.dofs[U_lam] = 3; m.ndf = 3;
  m.ID.assign((size_t)m.ndf * m.numnp, 0);
  for (int b : {TOP, BOTTOM})
    for (int64_t i : m.bdry_nodes[b]) { IDX(m, 1, i) = -1; IDX(m, 2, i) = -1; }
  for (int b : {LEFT, RIGHT})
    for (int64_t i : m.bdry_nodes[b]) IDX(m, 2, i) = -1;
  m.inh_neu = {{LEFT, STRETCH, 4.0}, {RIGHT, STRETCH, 8.0}};
}
// Bc.jl:188-269
static void bc_f_pull(Mesh& m) {
  const Params& p = m.p;
  get_dofs(p.motion, m.dofs, &m.ndf);
  m.ID.assign((size_t)m.ndf * m.numnp, 0);
  for (int b = 1; b <= 4; ++b)
    for (int64_t i : m.bdry_nodes[b]) {
      IDX(m, m.dofs[U_vz], i) = -1;
      if (p.motion == ALEV || p.motion == ALEVB) {
        IDX(m, m.dofs[U_vmx], i) = -1; IDX(m, m.dofs[U_vmy], i) = -1; IDX(m, m.dofs[U_vmz], i) = -1;
      }
    }
  for (int b = 1; b <= 4; ++b)
    for (int64_t i : m.bdry_inner_nodes[b]) {
      IDX(m, m.dofs[U_vz], i) = -1;
      if (p.motion == ALEVB) IDX(m, m.dofs[U_vmz], i) = -1;
    }
  int64_t ce = get_pull_el_id(m.numel);
  for (int a = 0; a < NEN; ++a) {
    int64_t nd = m.IX[(size_t)a + (size_t)NEN * (ce - 1)];
    IDX(m, m.dofs[U_vx], nd) = -1; IDX(m, m.dofs[U_vy], nd) = -1; IDX(m, m.dofs[U_vz], nd) = -1;
    m.inh_dir.push_back({U_vz, (int)nd, p.pull_speed});
    if (p.motion != LAG) {
      IDX(m, m.dofs[U_vmx], nd) = -1; IDX(m, m.dofs[U_vmy], nd) = -1; IDX(m, m.dofs[U_vmz], nd) = -1;
      m.inh_dir.push_back({U_vmz, (int)nd, p.pull_speed});
    }
  }
  for (int b = 1; b <= 4; ++b) {
    const auto& bn = m.bdry_nodes[b];
    int64_t c = bn[bn.size() / 2];  // bdry_nodes[bdry][floor(end/2)+1], 1-based
    IDX(m, m.dofs[U_vx], c) = -1; IDX(m, m.dofs[U_vy], c) = -1;
    if (p.motion != LAG) { IDX(m, m.dofs[U_vmx], c) = -1; IDX(m, m.dofs[U_vmy], c) = -1; }
  }
  double lval = p.kb / 4;
  m.inh_neu = {{LEFT, STRETCH, lval}, {RIGHT, STRETCH, lval}, {TOP, STRETCH, lval}, {BOTTOM, STRETCH, lval}};
  ORC_CHECK(p.pn == 0.0, "F_PULL with a normal pressure is not implemented");
}
// Bc.jl:347-393
static void bc_f_bend(Mesh& m) {
  const Params& p = m.p;
  get_dofs(p.motion, m.dofs, &m.ndf);
  m.ID.assign((size_t)m.ndf * m.numnp, 0);
  bool ale = (p.motion == ALEV || p.motion == ALEVB);
  for (int b : {TOP, BOTTOM})
    for (int64_t i : m.bdry_nodes[b]) {
      IDX(m, m.dofs[U_vy], i) = -1;
      if (ale) IDX(m, m.dofs[U_vmy], i) = -1;
    }
  for (int64_t i : m.bdry_nodes[LEFT]) {
    IDX(m, m.dofs[U_vx], i) = -1; IDX(m, m.dofs[U_vy], i) = -1; IDX(m, m.dofs[U_vz], i) = -1;
    if (ale) { IDX(m, m.dofs[U_vmx], i) = -1; IDX(m, m.dofs[U_vmy], i) = -1; IDX(m, m.dofs[U_vmz], i) = -1; }
  }
  for (int64_t i : m.bdry_nodes[RIGHT]) {
    IDX(m, m.dofs[U_vz], i) = -1;
    if (ale) IDX(m, m.dofs[U_vmz], i) = -1;
  }
  m.inh_neu = {{LEFT, MOMENT, p.bend_mf}, {RIGHT, MOMENT, p.bend_mf}};
}

// Mesh.jl:94-302 (FLAT topology; CYLINDER has no implemented scenario, Mesh.jl:554-555)
static Mesh* generate_mesh(const Params& p) {
  Mesh* mp = new Mesh();
  Mesh& m = *mp;
  m.p = p;
  m.num1el = p.num1el; m.num2el = p.num2el;
  m.numel = p.num1el * p.num2el;
  m.num1np = p.num1el + POLY; m.num2np = p.num2el + POLY;
  m.numnp = m.num1np * m.num2np;
  const int64_t n1 = m.num1np, nn = m.numnp, ne = m.numel, e1 = p.num1el;
  for (int64_t i = 1; i <= e1; ++i) m.bdry_elems[BOTTOM].push_back(i);
  for (int64_t i = e1; i <= ne; i += e1) m.bdry_elems[RIGHT].push_back(i);
  for (int64_t i = ne - e1 + 1; i <= ne; ++i) m.bdry_elems[TOP].push_back(i);
  for (int64_t i = 1; i <= ne - e1 + 1; i += e1) m.bdry_elems[LEFT].push_back(i);
  for (int64_t i = 1; i <= n1; ++i) { m.bdry_nodes[BOTTOM].push_back(i); m.bdry_inner_nodes[BOTTOM].push_back(i + n1); }
  for (int64_t i = n1; i <= nn; i += n1) { m.bdry_nodes[RIGHT].push_back(i); m.bdry_inner_nodes[RIGHT].push_back(i - 1); }
  for (int64_t i = nn - n1 + 1; i <= nn; ++i) { m.bdry_nodes[TOP].push_back(i); m.bdry_inner_nodes[TOP].push_back(i - n1); }
  for (int64_t i = 1; i <= nn - n1 + 1; i += n1) { m.bdry_nodes[LEFT].push_back(i); m.bdry_inner_nodes[LEFT].push_back(i + 1); }
  m.crnr_nodes[1] = 1; m.crnr_nodes[2] = n1; m.crnr_nodes[3] = nn - n1 + 1; m.crnr_nodes[4] = nn;
  if (p.scenario == F_PULL && p.num1el >= 18 && p.num2el >= 18) {  // Mesh.jl:176-181
    m.kv1 = knot_vector_from_list(get_fine_zs(p.num1el, POLY), POLY, CLAMPED);
    m.kv2 = knot_vector_from_list(get_fine_zs(p.num2el, POLY), POLY, CLAMPED);
  } else {
    m.kv1 = knot_vector_uniform(p.num1el, POLY, CLAMPED);
    m.kv2 = knot_vector_uniform(p.num2el, POLY, CLAMPED);
  }
  m.line1 = line_gp_basis_fns(m.kv1, GP1D);
  m.line2 = line_gp_basis_fns(m.kv2, GP1D);
  m.area = area_gp_basis_fns(m.line1, m.line2);
  m.bdry[BOTTOM] = bdry_gp_basis_fns(m.line1, m.line2.zmin, BOTTOM);  // Mesh.jl:220-225
  m.bdry[RIGHT] = bdry_gp_basis_fns(m.line2, m.line1.zmax, RIGHT);
  m.bdry[TOP] = bdry_gp_basis_fns(m.line1, m.line2.zmax, TOP);
  m.bdry[LEFT] = bdry_gp_basis_fns(m.line2, m.line1.zmin, LEFT);
  m.IX = construct_IX(m.kv1, m.kv2, m.num1np);
  // generate_scenario, Mesh.jl:262-302
  switch (p.scenario) {
    case F_CAVI: bc_f_cavi(m); break;
    case F_COUE: bc_f_coue(m); break;
    case F_POIS: bc_f_pois(m); break;
    case F_PULL: bc_f_pull(m); break;
    case F_BEND: bc_f_bend(m); break;
    default: ORC_CHECK(false, "Need boundary conditions for scenario");
  }
  int64_t dof_index = 1;
  for (int64_t node = 1; node <= nn; ++node)
    for (int d = 1; d <= m.ndf; ++d) {
      if (IDX(m, d, node) != -1) { IDX(m, d, node) = dof_index; dof_index += 1; }
      else IDX(m, d, node) = 0;
    }
  m.nmdf = 0;
  for (int64_t v : m.ID) m.nmdf = std::max(m.nmdf, v);
  for (int64_t node = 1; node <= nn; ++node)
    for (int d = 1; d <= m.ndf; ++d)
      if (IDX(m, d, node) != 0) { m.ID_inv_node.push_back(node); m.ID_inv_dof.push_back(d); }
  m.LM.assign((size_t)m.ndf * NEN * ne, 0);  // LM = reshape(ID[:,IX], (ndf*NEN, numel))
  for (int64_t e = 0; e < ne; ++e)
    for (int a = 0; a < NEN; ++a)
      for (int d = 0; d < m.ndf; ++d)
        m.LM[(size_t)d + (size_t)m.ndf * a + (size_t)m.ndf * NEN * e] =
            m.ID[(size_t)d + (size_t)m.ndf * (m.IX[(size_t)a + (size_t)NEN * e] - 1)];
  return mp;
}

// Mesh.jl:477-542
static void get_v_order(const int* dofs, int* o) { o[0] = dofs[U_vx]; o[1] = dofs[U_vy]; o[2] = dofs[U_vz]; }
static void get_m_order(const int* dofs, int* o) { o[0] = dofs[U_vmx]; o[1] = dofs[U_vmy]; o[2] = dofs[U_vmz]; }
static void get_m_motion_order(int motion, const int* dofs, int* o) {
  if (motion == STATIC) { o[0] = o[1] = o[2] = 0; }
  else if (motion == LAG) get_v_order(dofs, o);
  else get_m_order(dofs, o);
}

// =============================================================================
// GeoDynStress.jl
// =============================================================================
struct GDS {  // GeoDynStress.jl:42-80
  cd x[3], a_[2][3], ddx[3][3], aco[2][2], acon[2][2], aup[2][3], J, Gam[3][2], n[3], b[2][2], H, K;
  cd v[3], dv[2][3], ddv[3][3], lam, pm, vm[3], dvm[2][3], ddvm[3][3];
  cd sig[3], sigm[3], M[3];
  cd Ba[3][27], Bb[3][27];
};

// GeoDynStress.jl:190-205 : column of cps for `unknown`, zero if absent
static inline cd dof_cp(const cd* cps /*9 x ndf col-major*/, const int* dofs, int unknown, int a) {
  int c = dofs[unknown];
  return c == 0 ? cd(0.0, 0.0) : cps[(size_t)a + (size_t)NEN * (c - 1)];
}

// GeoDynStress.jl:91-178. xms: 9x3 col-major, cps: 9 x ndf col-major.
static void geo_dyn_stress(const cd* xms, const cd* cps, const int* dofs, const double* N,
                           const double (*dN)[2], const double (*ddN)[3], double kb, double kg, double zv,
                           GDS& g) {
  cd vc[9][3], vmc[9][3], lc[9], pc[9];
  for (int a = 0; a < 9; ++a) {
    vc[a][0] = dof_cp(cps, dofs, U_vx, a); vc[a][1] = dof_cp(cps, dofs, U_vy, a); vc[a][2] = dof_cp(cps, dofs, U_vz, a);
    vmc[a][0] = dof_cp(cps, dofs, U_vmx, a); vmc[a][1] = dof_cp(cps, dofs, U_vmy, a); vmc[a][2] = dof_cp(cps, dofs, U_vmz, a);
    lc[a] = dof_cp(cps, dofs, U_lam, a); pc[a] = dof_cp(cps, dofs, U_pm, a);
  }
  // geometry (:111-130)
  for (int i = 0; i < 3; ++i) {
    g.x[i] = 0;
    for (int a = 0; a < 9; ++a) g.x[i] += xms[a + 9 * i] * N[a];
    for (int al = 0; al < 2; ++al) {
      g.a_[al][i] = 0;
      for (int a = 0; a < 9; ++a) g.a_[al][i] += xms[a + 9 * i] * dN[a][al];
    }
    for (int k = 0; k < 3; ++k) {
      g.ddx[k][i] = 0;
      for (int a = 0; a < 9; ++a) g.ddx[k][i] += xms[a + 9 * i] * ddN[a][k];
    }
  }
  for (int al = 0; al < 2; ++al)
    for (int be = 0; be < 2; ++be) {
      g.aco[al][be] = 0;
      for (int i = 0; i < 3; ++i) g.aco[al][be] += g.a_[al][i] * g.a_[be][i];
    }
  cd det = g.aco[0][0] * g.aco[1][1] - g.aco[0][1] * g.aco[1][0];
  cd idet = 1.0 / det;  // StaticArrays 2x2 inv
  g.acon[0][0] = g.aco[1][1] * idet; g.acon[0][1] = -g.aco[0][1] * idet;
  g.acon[1][0] = -g.aco[1][0] * idet; g.acon[1][1] = g.aco[0][0] * idet;
  for (int al = 0; al < 2; ++al)
    for (int i = 0; i < 3; ++i) g.aup[al][i] = g.a_[0][i] * g.acon[0][al] + g.a_[1][i] * g.acon[1][al];
  g.J = c_sqrt(det);
  for (int k = 0; k < 3; ++k)
    for (int mu = 0; mu < 2; ++mu) {
      g.Gam[k][mu] = 0;
      for (int i = 0; i < 3; ++i) g.Gam[k][mu] += g.ddx[k][i] * g.aup[mu][i];
    }
  const cd* a1 = g.a_[0];
  const cd* a2 = g.a_[1];
  g.n[0] = (a1[1] * a2[2] - a1[2] * a2[1]) / g.J;
  g.n[1] = (a1[2] * a2[0] - a1[0] * a2[2]) / g.J;
  g.n[2] = (a1[0] * a2[1] - a1[1] * a2[0]) / g.J;
  cd bf[3];
  for (int k = 0; k < 3; ++k) {
    bf[k] = 0;
    for (int i = 0; i < 3; ++i) bf[k] += g.ddx[k][i] * g.n[i];
  }
  g.b[0][0] = bf[0]; g.b[0][1] = bf[2]; g.b[1][0] = bf[2]; g.b[1][1] = bf[1];
  cd tmp[2][2], bcon[2][2];
  for (int al = 0; al < 2; ++al)
    for (int be = 0; be < 2; ++be) tmp[al][be] = g.acon[al][0] * g.b[0][be] + g.acon[al][1] * g.b[1][be];
  for (int al = 0; al < 2; ++al)
    for (int be = 0; be < 2; ++be) bcon[al][be] = tmp[al][0] * g.acon[0][be] + tmp[al][1] * g.acon[1][be];
  g.H = (g.acon[0][0] * g.b[0][0] + g.acon[1][0] * g.b[1][0] + g.acon[0][1] * g.b[0][1] + g.acon[1][1] * g.b[1][1]) / 2.0;
  g.K = (g.b[0][0] * g.b[1][1] - g.b[0][1] * g.b[1][0]) / det;
  // dynamics (:135-144)
  for (int i = 0; i < 3; ++i) {
    g.v[i] = 0; g.vm[i] = 0;
    for (int a = 0; a < 9; ++a) { g.v[i] += vc[a][i] * N[a]; g.vm[i] += vmc[a][i] * N[a]; }
    for (int al = 0; al < 2; ++al) {
      g.dv[al][i] = 0; g.dvm[al][i] = 0;
      for (int a = 0; a < 9; ++a) { g.dv[al][i] += vc[a][i] * dN[a][al]; g.dvm[al][i] += vmc[a][i] * dN[a][al]; }
    }
    for (int k = 0; k < 3; ++k) {
      g.ddv[k][i] = 0; g.ddvm[k][i] = 0;
      for (int a = 0; a < 9; ++a) { g.ddv[k][i] += vc[a][i] * ddN[a][k]; g.ddvm[k][i] += vmc[a][i] * ddN[a][k]; }
    }
  }
  g.lam = 0; g.pm = 0;
  for (int a = 0; a < 9; ++a) { g.lam += lc[a] * N[a]; g.pm += pc[a] * N[a]; }
  // stresses (:148-159)
  auto visc = [&](const cd (*dvel)[3], cd out[2][2]) {
    // pi = transpose(a^alpha) * (dv * a^{alpha beta}); then (pi + pi^T) * zv
    cd w[2][3];  // (dv * acon): 3 x 2, stored [beta][i]
    for (int be = 0; be < 2; ++be)
      for (int i = 0; i < 3; ++i) w[be][i] = dvel[0][i] * g.acon[0][be] + dvel[1][i] * g.acon[1][be];
    cd pi[2][2];
    for (int al = 0; al < 2; ++al)
      for (int be = 0; be < 2; ++be) {
        pi[al][be] = 0;
        for (int i = 0; i < 3; ++i) pi[al][be] += g.aup[al][i] * w[be][i];
      }
    for (int al = 0; al < 2; ++al)
      for (int be = 0; be < 2; ++be) out[al][be] = (pi[al][be] + pi[be][al]) * zv;
  };
  cd pi[2][2], pim[2][2], sg[2][2], Mm[2][2], sgm[2][2];
  visc(g.dv, pi);
  visc(g.dvm, pim);
  for (int al = 0; al < 2; ++al)
    for (int be = 0; be < 2; ++be) {
      sg[al][be] = g.acon[al][be] * (kb * g.H * g.H - kg * g.K + g.lam) - bcon[al][be] * 2.0 * kb * g.H + pi[al][be];
      Mm[al][be] = g.acon[al][be] * g.H * (kb + 2 * kg) - bcon[al][be] * kg;
      sgm[al][be] = g.acon[al][be] * (kb * g.H * g.H - kg * g.K) - bcon[al][be] * 2.0 * kb * g.H + pim[al][be];
    }
  g.sig[0] = sg[0][0]; g.sig[1] = sg[1][1]; g.sig[2] = 2.0 * sg[0][1];
  g.M[0] = Mm[0][0]; g.M[1] = Mm[1][1]; g.M[2] = Mm[0][1] + Mm[1][0];
  g.sigm[0] = sgm[0][0]; g.sigm[1] = sgm[1][1]; g.sigm[2] = 2.0 * sgm[0][1];
  // shape function matrices (:163-174), column index comp + 3*(node-1)
  cd ddNc[9][3];  // nabla nabla N = ddN - dN * transpose(Gam)
  for (int a = 0; a < 9; ++a)
    for (int k = 0; k < 3; ++k) ddNc[a][k] = ddN[a][k] - (dN[a][0] * g.Gam[k][0] + dN[a][1] * g.Gam[k][1]);
  for (int a = 0; a < 9; ++a)
    for (int i = 0; i < 3; ++i) {
      int c = i + 3 * a;
      g.Ba[0][c] = a1[i] * dN[a][0];
      g.Ba[1][c] = a2[i] * dN[a][1];
      g.Ba[2][c] = (a1[i] * dN[a][1] + a2[i] * dN[a][0]) * 0.5;
      for (int k = 0; k < 3; ++k) g.Bb[k][c] = g.n[i] * ddNc[a][k];
    }
}

// =============================================================================
// FiniteElement.jl
// =============================================================================
static const Fn2& area_fns(const Mesh& m, int64_t el, int gp) {  // Mesh.jl:311-319
  return m.area.ufns[(size_t)(m.area.uel_ids[el - 1] - 1) * 9 + (gp - 1)];
}
static const Fn2& bdry_fns(const Mesh& m, int bdry, int64_t el, int gp) {  // Mesh.jl:390-403
  const auto& be = m.bdry_elems[bdry];
  auto it = std::find(be.begin(), be.end(), el);
  ORC_CHECK(it != be.end(), "element id not found on boundary");
  size_t bel = (size_t)(it - be.begin());
  return m.bdry[bdry].ufns[(size_t)(m.bdry[bdry].uel_ids[bel] - 1) * GP1D + (gp - 1)];
}

// FiniteElement.jl:253-330. Outputs rv(27) rm(27) rl(9) rp(9) (rp unused when p absent).
static void calc_elem_dof_residuals(const Mesh& m, int64_t el, const cd* xms, const cd* cps, cd* rv, cd* rm,
                                    cd* rl, cd* rp) {
  const Params& p = m.p;
  const int p_order = m.dofs[U_pm];
  for (int i = 0; i < 27; ++i) rv[i] = rm[i] = 0;
  for (int i = 0; i < 9; ++i) rl[i] = rp[i] = 0;
  cd GDB[3][9], HDB[3][3];
  for (int i = 0; i < 3; ++i) {
    for (int a = 0; a < 9; ++a) GDB[i][a] = 0;
    for (int j = 0; j < 3; ++j) HDB[i][j] = 0;
  }
  double xs[4], wx[4];
  gauss_xi(GP1D, xs, wx);
  GDS g;
  for (int gp = 1; gp <= GP1D * GP1D; ++gp) {
    const Fn2& f = area_fns(m, el, gp);
    const double* N = f.N;
    const double gpw = f.w;
    geo_dyn_stress(xms, cps, m.dofs, f.N, f.dN, f.ddN, p.kb, p.kg, p.zv, g);
    for (int c = 0; c < 27; ++c) {  // :293-294
      cd s = 0;
      for (int k = 0; k < 3; ++k) s += g.Ba[k][c] * g.sig[k];
      cd t = 0;
      for (int k = 0; k < 3; ++k) t += g.Bb[k][c] * g.M[k];
      rv[c] += (s + t) * g.J * gpw;
    }
    if (p.pn != 0.0)  // :295-297
      for (int a = 0; a < 9; ++a)
        for (int i = 0; i < 3; ++i) rv[i + 3 * a] -= g.n[i] * N[a] * p.pn * g.J * gpw;
    cd trv = 0;  // tr(transpose(a^alpha) * dv_alpha)
    for (int al = 0; al < 2; ++al)
      for (int i = 0; i < 3; ++i) trv += g.aup[al][i] * g.dv[al][i];
    for (int a = 0; a < 9; ++a) rl[a] += N[a] * gpw * (g.J * trv - p.adb * g.lam / p.zv);  // :298-299
    if (p.motion == EUL) {  // :300-303
      cd ndv = 0;
      for (int i = 0; i < 3; ++i) ndv += g.n[i] * g.v[i];
      for (int a = 0; a < 9; ++a)
        for (int i = 0; i < 3; ++i) rm[i + 3 * a] += (g.vm[i] - g.n[i] * ndv) * N[a] * p.am * g.J * gpw;
    } else if (p.motion == ALEV || p.motion == ALEVB) {  // :304-313
      for (int c = 0; c < 27; ++c) {
        cd s = 0;
        for (int k = 0; k < 3; ++k) s += g.Ba[k][c] * g.sigm[k];
        rm[c] += s * g.J * gpw;
      }
      if (p.motion == ALEVB)
        for (int c = 0; c < 27; ++c) {
          cd t = 0;
          for (int k = 0; k < 3; ++k) t += g.Bb[k][c] * g.M[k];
          rm[c] += t * g.J * gpw;
        }
      for (int a = 0; a < 9; ++a)
        for (int i = 0; i < 3; ++i) rm[i + 3 * a] -= g.n[i] * N[a] * g.pm * g.J * gpw;
      cd ndm = 0;
      for (int i = 0; i < 3; ++i) ndm += g.n[i] * (g.vm[i] - g.v[i]);
      for (int a = 0; a < 9; ++a) {
        rp[a] -= N[a] * gpw * g.J * ndm;
        rp[a] -= N[a] * gpw * p.adb * g.pm / p.zv;
      }
    }
    double NDB[3] = {xs[(gp - 1) % GP1D], xs[(gp - 1) / GP1D], 1.0};  // :315-317
    for (int i = 0; i < 3; ++i) {
      for (int a = 0; a < 9; ++a) GDB[i][a] += NDB[i] * N[a] * gpw;
      for (int j = 0; j < 3; ++j) HDB[i][j] += NDB[i] * NDB[j] * gpw;
    }
  }
  // tmpDB = transpose(GDB) * inv(HDB) * GDB  (:323) -- 3x3 inverse by adjugate (StaticArrays)
  cd Hi[3][3];
  {
    cd a = HDB[0][0], b = HDB[0][1], c = HDB[0][2], d = HDB[1][0], e = HDB[1][1], f = HDB[1][2], gg = HDB[2][0],
       h = HDB[2][1], i = HDB[2][2];
    cd det = a * (e * i - f * h) - b * (d * i - f * gg) + c * (d * h - e * gg);
    cd id = 1.0 / det;
    Hi[0][0] = (e * i - f * h) * id; Hi[0][1] = (c * h - b * i) * id; Hi[0][2] = (b * f - c * e) * id;
    Hi[1][0] = (f * gg - d * i) * id; Hi[1][1] = (a * i - c * gg) * id; Hi[1][2] = (c * d - a * f) * id;
    Hi[2][0] = (d * h - e * gg) * id; Hi[2][1] = (b * gg - a * h) * id; Hi[2][2] = (a * e - b * d) * id;
  }
  cd T1[9][3];  // transpose(GDB) * inv(HDB)
  for (int a = 0; a < 9; ++a)
    for (int j = 0; j < 3; ++j) {
      T1[a][j] = 0;
      for (int i = 0; i < 3; ++i) T1[a][j] += GDB[i][a] * Hi[i][j];
    }
  for (int a = 0; a < 9; ++a) {
    cd sl = 0, sp = 0;
    for (int b = 0; b < 9; ++b) {
      cd t = 0;
      for (int j = 0; j < 3; ++j) t += T1[a][j] * GDB[j][b];
      sl += t * dof_cp(cps, m.dofs, U_lam, b);
      if (p_order != 0) sp += t * cps[(size_t)b + (size_t)NEN * (p_order - 1)];
    }
    rl[a] += sl * p.adb / p.zv;  // :324
    if (p.motion == ALEV || p.motion == ALEVB) rp[a] += sp * p.adb / p.zv;  // :325-327
  }
}

// FiniteElement.jl:208-242
static void calc_elem_residual(const Mesh& m, int64_t el, const cd* xms, const cd* cps, cd* r_el) {
  int vo[3], mo[3];
  get_v_order(m.dofs, vo);
  get_m_order(m.dofs, mo);
  const int lo = m.dofs[U_lam], po = m.dofs[U_pm];
  cd rv[27], rm[27], rl[9], rp[9];
  calc_elem_dof_residuals(m, el, xms, cps, rv, rm, rl, rp);
  for (int i = 0; i < NEN * m.ndf; ++i) r_el[i] = 0;
  for (int a = 0; a < NEN; ++a) {
    for (int j = 0; j < 3; ++j)
      if (vo[j] != 0) r_el[vo[j] - 1 + m.ndf * a] = rv[j + 3 * a];
    for (int j = 0; j < 3; ++j)
      if (mo[j] != 0) r_el[mo[j] - 1 + m.ndf * a] = rm[j + 3 * a];
    r_el[lo - 1 + m.ndf * a] = rl[a];
    if (po != 0) r_el[po - 1 + m.ndf * a] = rp[a];
  }
}

// FiniteElement.jl:431-452
static void calc_tau_nu(int bdry, const cd a_[2][3], const cd n[3], cd tau[3], cd nu[3]) {
  for (int i = 0; i < 3; ++i) {
    if (bdry == BOTTOM) tau[i] = a_[0][i];
    else if (bdry == RIGHT) tau[i] = a_[1][i];
    else if (bdry == TOP) tau[i] = -a_[0][i];
    else tau[i] = -a_[1][i];
  }
  cd d = 0;  // dot(tau,tau) conjugates its first argument (:448)
  for (int i = 0; i < 3; ++i) d += c_conj(tau[i]) * tau[i];
  cd s = c_sqrt(d);
  for (int i = 0; i < 3; ++i) tau[i] /= s;
  nu[0] = tau[1] * n[2] - tau[2] * n[1];
  nu[1] = tau[2] * n[0] - tau[0] * n[2];
  nu[2] = tau[0] * n[1] - tau[1] * n[0];
}

// FiniteElement.jl:338-400
static void calc_bdry_element_residual(const Mesh& m, int bdry, int ntype, double nval, int64_t el, const cd* xms,
                                       const cd* cps, double time, cd* r_el) {
  const Params& p = m.p;
  int vo[3], mo[3];
  get_v_order(m.dofs, vo);
  get_m_order(m.dofs, mo);
  cd rv[27], rm[27];
  for (int i = 0; i < 27; ++i) rv[i] = rm[i] = 0;
  GDS g;
  for (int gp = 1; gp <= GP1D; ++gp) {
    const Fn2& f = bdry_fns(m, bdry, el, gp);
    geo_dyn_stress(xms, cps, m.dofs, f.N, f.dN, f.ddN, p.kb, p.kg, p.zv, g);
    cd tau[3], nu[3];
    calc_tau_nu(bdry, g.a_, g.n, tau, nu);
    cd s2 = 0;
    for (int al = 0; al < 2; ++al) {
      cd t = 0;
      for (int i = 0; i < 3; ++i) t += g.aup[al][i] * tau[i];
      s2 += t * t;
    }
    cd JG = 1.0 / c_sqrt(s2);
    if (ntype == STRETCH || ntype == SHEAR) {
      for (int a = 0; a < 9; ++a)
        for (int i = 0; i < 3; ++i) {
          cd fi = (ntype == STRETCH) ? nval * nu[i] : nval * tau[i];
          rv[i + 3 * a] -= fi * f.N[a] * JG * f.w;
        }
    } else if (ntype == MOMENT && p.scenario == F_BEND) {
      cd nua[2];
      for (int al = 0; al < 2; ++al) {
        nua[al] = 0;
        for (int i = 0; i < 3; ++i) nua[al] += g.aup[al][i] * nu[i];
      }
      double Mval = nval * std::min(time / p.bend_tm, 1.0);
      for (int a = 0; a < 9; ++a) {
        cd dNnu = f.dN[a][0] * nua[0] + f.dN[a][1] * nua[1];
        for (int i = 0; i < 3; ++i) rv[i + 3 * a] -= g.n[i] * dNnu * Mval * JG * f.w;
      }
    } else ORC_CHECK(false, "Neumann boundary condition not implemented");
  }
  for (int i = 0; i < NEN * m.ndf; ++i) r_el[i] = 0;
  for (int a = 0; a < NEN; ++a) {
    for (int j = 0; j < 3; ++j)
      if (vo[j] != 0) r_el[vo[j] - 1 + m.ndf * a] = rv[j + 3 * a];
    for (int j = 0; j < 3; ++j)
      if (mo[j] != 0) r_el[mo[j] - 1 + m.ndf * a] = rm[j + 3 * a];
  }
}

// Element-level complex-step tangent shared by the area loop (FiniteElement.jl:100-126)
// and the Neumann loop (:156-184). K_el is (9 ndf)^2 column-major.
// r_mag / K_mag (may be NULL; zeros in the regular build): the error-bound scale E of every entry of r_el / K_el.
template <class ResFn>
static void elem_r_K(const Mesh& m, int64_t el, const double* xms_gl, const double* cps_gl, double dt,
                     const int* mmo, ResFn&& res, std::vector<real_t>& r_el, std::vector<real_t>& K_el,
                     std::vector<real_t>* r_mag = nullptr, std::vector<real_t>* K_mag = nullptr) {
  const int ndf = m.ndf, nd = NEN * ndf;
  const double ek = m.p.ek;
  std::vector<cd> xe(27), ce((size_t)9 * ndf), out(nd);
  for (int a = 0; a < 9; ++a) {
    int64_t nodeid = m.IX[(size_t)a + (size_t)NEN * (el - 1)] - 1;
    for (int i = 0; i < 3; ++i) xe[a + 9 * i] = xms_gl[(size_t)nodeid + (size_t)m.numnp * i];
    for (int d = 0; d < ndf; ++d) ce[a + 9 * d] = cps_gl[(size_t)nodeid + (size_t)m.numnp * d];
  }
  res(xe.data(), ce.data(), out.data());
  r_el.assign(nd, 0.0);
  K_el.assign((size_t)nd * nd, 0.0);
  for (int i = 0; i < nd; ++i) r_el[i] = out[i].real();
  if (K_mag) {
    r_mag->assign(nd, 0.0);
    K_mag->assign((size_t)nd * nd, 0.0);
    for (int i = 0; i < nd; ++i) (*r_mag)[i] = c_er(out[i]);
  }
  for (int a = 0; a < 9; ++a) {
    int64_t node = m.IX[(size_t)a + (size_t)NEN * (el - 1)];
    for (int d = 1; d <= ndf; ++d) {
      if (m.ID[(size_t)(d - 1) + (size_t)ndf * (node - 1)] == 0) continue;
      const int col = (d - 1) + ndf * a;
      cd save = ce[a + 9 * (d - 1)];
      ce[a + 9 * (d - 1)] += cd(0.0, ek);
      res(xe.data(), ce.data(), out.data());
      ce[a + 9 * (d - 1)] = save;
      for (int i = 0; i < nd; ++i) K_el[(size_t)i + (size_t)nd * col] += out[i].imag() / ek;
      if (K_mag)   // the division by eps_k and the addition into K_el: one rounding each
        for (int i = 0; i < nd; ++i)
          (*K_mag)[(size_t)i + (size_t)nd * col] += (c_ei(out[i]) + 2 * r_abs(out[i].imag())) / ek;
      int comp = -1;
      for (int j = 0; j < 3; ++j)
        if (mmo[j] == d) { comp = j; break; }
      if (comp >= 0) {
        cd sx = xe[a + 9 * comp];
        xe[a + 9 * comp] += cd(0.0, ek);
        res(xe.data(), ce.data(), out.data());
        xe[a + 9 * comp] = sx;
        for (int i = 0; i < nd; ++i) K_el[(size_t)i + (size_t)nd * col] += (out[i].imag() / ek) * dt;
        if (K_mag)
          for (int i = 0; i < nd; ++i)
            (*K_mag)[(size_t)i + (size_t)nd * col] += ((c_ei(out[i]) + 3 * r_abs(out[i].imag())) / ek) * r_abs((real_t)dt);
      }
    }
  }
}

// Julia SparseMatrixCSC scalar `K[i,j] += v` semantics: an entry is created only when the value to store
// is non-zero; once stored it stays stored (explicit zeros survive). Key = (col << 32) | row, 0-based.
struct SpAcc {
  std::unordered_map<uint64_t, real_t> m;
  void add(int64_t row, int64_t col, real_t v) {
    uint64_t key = ((uint64_t)col << 32) | (uint64_t)row;
    auto it = m.find(key);
    if (it == m.end()) {
      real_t nv = 0.0 + v;
      if (nv != 0.0) m.emplace(key, nv);
    } else it->second += v;
  }
};

struct Result {
  std::vector<double> r;
  std::vector<int64_t> colptr, rowval;  // 1-based CSC like SparseMatrixCSC
  std::vector<double> nzval;
};

struct FastAcc {  // accumulate into a caller-supplied 0-based CSC pattern (CPU-baseline mode)
  const int64_t* colptr; const int64_t* rowval; std::vector<real_t> nz;
  void add(int64_t row, int64_t col, real_t v) {
    const int64_t* b = rowval + colptr[col];
    const int64_t* e = rowval + colptr[col + 1];
    const int64_t* it = std::lower_bound(b, e, row);
    if (it != e && *it == row) nz[(size_t)(it - rowval)] += v;
  }
};

// FiniteElement.jl:75-200
template <class Acc>
static void area_chunk(const Mesh& m, const double* xms, const double* cps, double dt, const int* mmo, int64_t e0,
                       int64_t e1, std::vector<real_t>& r_th, Acc& K_th, std::vector<real_t>* rm_th = nullptr,
                       Acc* Km_th = nullptr) {
  const int nd = NEN * m.ndf;
  std::vector<real_t> r_el, K_el, rm_el, Km_el;
  std::vector<int> ids;
  for (int64_t el = e0; el <= e1; ++el) {
    elem_r_K(m, el, xms, cps, dt, mmo,
             [&](const cd* xe, const cd* ce, cd* out) { calc_elem_residual(m, el, xe, ce, out); }, r_el, K_el,
             Km_th ? &rm_el : nullptr, Km_th ? &Km_el : nullptr);
    ids.clear();
    for (int i = 0; i < nd; ++i)
      if (m.LM[(size_t)i + (size_t)nd * (el - 1)] != 0) ids.push_back(i);
    for (int ri : ids) {
      int64_t gr = m.LM[(size_t)ri + (size_t)nd * (el - 1)];
      r_th[gr - 1] += r_el[ri];
      if (Km_th) (*rm_th)[gr - 1] += rm_el[ri] + r_abs(r_el[ri]);   // + the rounding of the global addition
      for (int ci : ids) {
        int64_t gc = m.LM[(size_t)ci + (size_t)nd * (el - 1)];
        K_th.add(gr - 1, gc - 1, K_el[(size_t)ri + (size_t)nd * ci]);
        if (Km_th)
          Km_th->add(gr - 1, gc - 1, Km_el[(size_t)ri + (size_t)nd * ci] + r_abs(K_el[(size_t)ri + (size_t)nd * ci]));
      }
    }
  }
}

template <class Acc>
static void neumann_loop(const Mesh& m, const double* xms, const double* cps, double time, double dt, const int* mmo,
                         std::vector<real_t>& r_gl, Acc& K_gl, std::vector<real_t>* rm_gl = nullptr,
                         Acc* Km_gl = nullptr) {
  const int nd = NEN * m.ndf;
  std::vector<real_t> r_el, K_el, rm_el, Km_el;
  std::vector<int> ids;
  for (const NeuBc& bc : m.inh_neu)
    for (int64_t el : m.bdry_elems[bc.bdry]) {
      elem_r_K(m, el, xms, cps, dt, mmo,
               [&](const cd* xe, const cd* ce, cd* out) {
                 calc_bdry_element_residual(m, bc.bdry, bc.type, bc.val, el, xe, ce, time, out);
               },
               r_el, K_el, Km_gl ? &rm_el : nullptr, Km_gl ? &Km_el : nullptr);
      ids.clear();
      for (int i = 0; i < nd; ++i)
        if (m.LM[(size_t)i + (size_t)nd * (el - 1)] != 0) ids.push_back(i);
      for (int ri : ids) {
        int64_t gr = m.LM[(size_t)ri + (size_t)nd * (el - 1)];
        r_gl[gr - 1] += r_el[ri];
        if (Km_gl) (*rm_gl)[gr - 1] += rm_el[ri] + r_abs(r_el[ri]);
        for (int ci : ids) {
          int64_t gc = m.LM[(size_t)ci + (size_t)nd * (el - 1)];
          K_gl.add(gr - 1, gc - 1, K_el[(size_t)ri + (size_t)nd * ci]);
          if (Km_gl)
            Km_gl->add(gr - 1, gc - 1, Km_el[(size_t)ri + (size_t)nd * ci] + r_abs(K_el[(size_t)ri + (size_t)nd * ci]));
        }
      }
    }
}

static std::vector<std::pair<int64_t, int64_t>> make_chunks(int64_t numel, int nthreads) {
  // chunk_size = ceil(max(1, numel/nthreads)); partition(1:numel, chunk_size)   (:88-89)
  double q = std::max(1.0, (double)numel / (double)nthreads);
  int64_t cs = (int64_t)std::ceil(q);
  std::vector<std::pair<int64_t, int64_t>> ch;
  for (int64_t s = 1; s <= numel; s += cs) ch.push_back({s, std::min(numel, s + cs - 1)});
  return ch;
}

static Result* calc_r_K(const Mesh& m, const double* xms, const double* cps, double time, double dt, int nthreads) {
  int mmo[3];
  get_m_motion_order(m.p.motion, m.dofs, mmo);
  auto chunks = make_chunks(m.numel, std::max(1, nthreads));
  const size_t nc = chunks.size();
  std::vector<std::vector<real_t>> r_th(nc, std::vector<real_t>((size_t)m.nmdf, 0.0));
  std::vector<SpAcc> K_th(nc);
  std::vector<std::thread> th;
  std::vector<std::string> errs(nc);
  for (size_t c = 0; c < nc; ++c)
    th.emplace_back([&, c]() {
      try { area_chunk(m, xms, cps, dt, mmo, chunks[c].first, chunks[c].second, r_th[c], K_th[c]); }
      catch (std::exception& e) { errs[c] = e.what(); }
    });
  for (auto& t : th) t.join();
  for (auto& e : errs) ORC_CHECK(e.empty(), e);
  // sum over tasks (:146-147): sparse `+` keeps only non-zero results
  std::vector<real_t> r_gl = r_th[0];
  SpAcc K_gl = std::move(K_th[0]);
  for (size_t c = 1; c < nc; ++c) {
    for (int64_t i = 0; i < m.nmdf; ++i) r_gl[i] += r_th[c][i];
    for (auto& kv : K_th[c].m) {
      auto it = K_gl.m.find(kv.first);
      if (it == K_gl.m.end()) K_gl.m.emplace(kv.first, kv.second); else it->second += kv.second;
    }
    for (auto it = K_gl.m.begin(); it != K_gl.m.end();)
      if (it->second == 0.0) it = K_gl.m.erase(it); else ++it;
  }
  neumann_loop(m, xms, cps, time, dt, mmo, r_gl, K_gl);
  Result* R = new Result();
  R->r.resize(r_gl.size());
  for (size_t i = 0; i < r_gl.size(); ++i) R->r[i] = (double)r_gl[i];
  std::vector<std::pair<uint64_t, real_t>> ent(K_gl.m.begin(), K_gl.m.end());
  std::sort(ent.begin(), ent.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  R->colptr.assign((size_t)m.nmdf + 1, 0);
  R->rowval.resize(ent.size());
  R->nzval.resize(ent.size());
  for (size_t k = 0; k < ent.size(); ++k) {
    int64_t col = (int64_t)(ent[k].first >> 32), row = (int64_t)(ent[k].first & 0xffffffffu);
    R->colptr[col + 1] += 1;
    R->rowval[k] = row + 1;
    R->nzval[k] = (double)ent[k].second;
  }
  R->colptr[0] = 1;
  for (int64_t c = 0; c < m.nmdf; ++c) R->colptr[c + 1] += R->colptr[c];
  return R;
}

// CPU-baseline mode: same element algorithm and threading scheme, accumulation into a given pattern.
static void calc_r_K_fast(const Mesh& m, const double* xms, const double* cps, double time, double dt, int nthreads,
                          const int64_t* colptr0, const int64_t* rowval0, int64_t nnz, int64_t e_first,
                          int64_t e_last, int with_neumann, double* r_out, double* nz_out,
                          double* rmag_out = nullptr, double* nzmag_out = nullptr) {
  int mmo[3];
  get_m_motion_order(m.p.motion, m.dofs, mmo);
  const int64_t ne = e_last - e_first + 1;
  auto chunks = make_chunks(ne, std::max(1, nthreads));
  const size_t nc = chunks.size();
  const bool want_mag = rmag_out || nzmag_out;   // truth builds only (the regular build accumulates zeros)
  std::vector<std::vector<real_t>> r_th(nc, std::vector<real_t>((size_t)m.nmdf, 0.0));
  std::vector<FastAcc> K_th(nc);
  for (auto& k : K_th) { k.colptr = colptr0; k.rowval = rowval0; k.nz.assign((size_t)nnz, 0.0); }
  std::vector<std::vector<real_t>> rm_th(want_mag ? nc : 0, std::vector<real_t>((size_t)m.nmdf, 0.0));
  std::vector<FastAcc> Km_th(want_mag ? nc : 0);
  for (auto& k : Km_th) { k.colptr = colptr0; k.rowval = rowval0; k.nz.assign((size_t)nnz, 0.0); }
  std::vector<std::thread> th;
  for (size_t c = 0; c < nc; ++c)
    th.emplace_back([&, c]() {
      area_chunk(m, xms, cps, dt, mmo, e_first - 1 + chunks[c].first, e_first - 1 + chunks[c].second, r_th[c], K_th[c],
                 want_mag ? &rm_th[c] : nullptr, want_mag ? &Km_th[c] : nullptr);
    });
  for (auto& t : th) t.join();
  for (size_t c = 1; c < nc; ++c) {
    for (int64_t i = 0; i < m.nmdf; ++i) r_th[0][i] += r_th[c][i];
    for (int64_t i = 0; i < nnz; ++i) K_th[0].nz[i] += K_th[c].nz[i];
    if (want_mag) {
      for (int64_t i = 0; i < m.nmdf; ++i) rm_th[0][i] += rm_th[c][i];
      for (int64_t i = 0; i < nnz; ++i) Km_th[0].nz[i] += Km_th[c].nz[i];
    }
  }
  if (with_neumann)
    neumann_loop(m, xms, cps, time, dt, mmo, r_th[0], K_th[0], want_mag ? &rm_th[0] : nullptr,
                 want_mag ? &Km_th[0] : nullptr);
  if (r_out) for (int64_t i = 0; i < m.nmdf; ++i) r_out[i] = (double)r_th[0][i];
  if (nz_out) for (int64_t i = 0; i < nnz; ++i) nz_out[i] = (double)K_th[0].nz[i];
  if (rmag_out) for (int64_t i = 0; i < m.nmdf; ++i) rmag_out[i] = (double)rm_th[0][i];
  if (nzmag_out) for (int64_t i = 0; i < nnz; ++i) nzmag_out[i] = (double)Km_th[0].nz[i];
}

// =============================================================================
// C interface (ctypes) -- test infrastructure
// =============================================================================
static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret)                                         \
  }                                                            \
  catch (std::exception & e) { g_err = e.what(); return ret; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

struct orc_params {
  int32_t motion, scenario, num1el, num2el;
  double length, kb, kg, zv, pn, adb, am, ek, pull_speed, bend_mf, bend_tm;
};

void* orc_mesh_create(const orc_params* q) {
  ORC_TRY
  Params p;
  p.motion = q->motion; p.scenario = q->scenario; p.num1el = q->num1el; p.num2el = q->num2el;
  p.length = q->length; p.kb = q->kb; p.kg = q->kg; p.zv = q->zv; p.pn = q->pn; p.adb = q->adb; p.am = q->am;
  p.ek = q->ek; p.pull_speed = q->pull_speed; p.bend_mf = q->bend_mf; p.bend_tm = q->bend_tm;
  return generate_mesh(p);
  ORC_CATCH(nullptr)
}
void orc_mesh_destroy(void* h) { delete (Mesh*)h; }

// sizes: [numel, numnp, ndf, nmdf, num1el, num2el, num1np, num2np, nuel1, nuel2, n_dir, n_neu, nknots1, nknots2]
void orc_mesh_sizes(void* h, int64_t* s) {
  Mesh& m = *(Mesh*)h;
  s[0] = m.numel; s[1] = m.numnp; s[2] = m.ndf; s[3] = m.nmdf; s[4] = m.num1el; s[5] = m.num2el;
  s[6] = m.num1np; s[7] = m.num2np; s[8] = m.line1.nuel; s[9] = m.line2.nuel; s[10] = (int64_t)m.inh_dir.size();
  s[11] = (int64_t)m.inh_neu.size(); s[12] = (int64_t)m.kv1.zs.size(); s[13] = (int64_t)m.kv2.zs.size();
}
void orc_mesh_dofs(void* h, int32_t* dofs8) {
  Mesh& m = *(Mesh*)h;
  for (int u = 1; u <= 8; ++u) dofs8[u - 1] = m.dofs[u];
}
void orc_mesh_IX(void* h, int64_t* o) { Mesh& m = *(Mesh*)h; std::copy(m.IX.begin(), m.IX.end(), o); }
void orc_mesh_ID(void* h, int64_t* o) { Mesh& m = *(Mesh*)h; std::copy(m.ID.begin(), m.ID.end(), o); }
void orc_mesh_LM(void* h, int64_t* o) { Mesh& m = *(Mesh*)h; std::copy(m.LM.begin(), m.LM.end(), o); }
void orc_mesh_ID_inv(void* h, int64_t* node, int64_t* dof) {
  Mesh& m = *(Mesh*)h;
  std::copy(m.ID_inv_node.begin(), m.ID_inv_node.end(), node);
  std::copy(m.ID_inv_dof.begin(), m.ID_inv_dof.end(), dof);
}
void orc_mesh_knots(void* h, int dir, double* o) {
  Mesh& m = *(Mesh*)h;
  const KnotVector& kv = dir == 1 ? m.kv1 : m.kv2;
  std::copy(kv.zs.begin(), kv.zs.end(), o);
}
// line tables: uel_ids (nel, 1-based), tab (nuel x 3 x 10: w,N[3],dN[3],ddN[3]), edge (2 x 10: zmin, zmax)
void orc_mesh_line(void* h, int dir, int64_t* uel_ids, double* tab, double* edge) {
  Mesh& m = *(Mesh*)h;
  const LineFns& L = dir == 1 ? m.line1 : m.line2;
  for (int i = 0; i < L.nel; ++i) uel_ids[i] = L.uel_ids[i];
  auto put = [](const Fn1& f, double* o) {
    o[0] = f.w;
    for (int i = 0; i < 3; ++i) { o[1 + i] = f.N[i]; o[4 + i] = f.dN[i]; o[7 + i] = f.ddN[i]; }
  };
  for (size_t k = 0; k < L.ufns.size(); ++k) put(L.ufns[k], tab + 10 * k);
  put(L.zmin, edge);
  put(L.zmax, edge + 10);
}
static void put_fn2(const Fn2& f, double* o) {  // 55 doubles: w, N[9], dN[9x2 col-major], ddN[9x3 col-major]
  o[0] = f.w;
  for (int a = 0; a < 9; ++a) o[1 + a] = f.N[a];
  for (int al = 0; al < 2; ++al)
    for (int a = 0; a < 9; ++a) o[10 + a + 9 * al] = f.dN[a][al];
  for (int k = 0; k < 3; ++k)
    for (int a = 0; a < 9; ++a) o[28 + a + 9 * k] = f.ddN[a][k];
}
void orc_mesh_area_fns(void* h, int64_t el, int gp, double* o55) { put_fn2(area_fns(*(Mesh*)h, el, gp), o55); }
int orc_mesh_bdry_fns(void* h, int bdry, int64_t el, int gp, double* o55) {
  ORC_TRY
  put_fn2(bdry_fns(*(Mesh*)h, bdry, el, gp), o55);
  return 0;
  ORC_CATCH(1)
}
void orc_mesh_area_uel_ids(void* h, int64_t* o) {
  Mesh& m = *(Mesh*)h;
  for (int i = 0; i < m.area.nel; ++i) o[i] = m.area.uel_ids[i];
}
int64_t orc_mesh_bdry_count(void* h, int bdry) { return (int64_t)((Mesh*)h)->bdry_elems[bdry].size(); }
void orc_mesh_bdry_elems(void* h, int bdry, int64_t* o) {
  Mesh& m = *(Mesh*)h;
  std::copy(m.bdry_elems[bdry].begin(), m.bdry_elems[bdry].end(), o);
}
void orc_mesh_bdry_nodes(void* h, int bdry, int64_t* nodes, int64_t* inner) {
  Mesh& m = *(Mesh*)h;
  std::copy(m.bdry_nodes[bdry].begin(), m.bdry_nodes[bdry].end(), nodes);
  std::copy(m.bdry_inner_nodes[bdry].begin(), m.bdry_inner_nodes[bdry].end(), inner);
}
void orc_mesh_bcs(void* h, int32_t* dir_unknown, int64_t* dir_node, double* dir_val, int32_t* neu_bdry,
                  int32_t* neu_type, double* neu_val) {
  Mesh& m = *(Mesh*)h;
  for (size_t i = 0; i < m.inh_dir.size(); ++i) {
    dir_unknown[i] = m.inh_dir[i].unknown; dir_node[i] = m.inh_dir[i].node; dir_val[i] = m.inh_dir[i].val;
  }
  for (size_t i = 0; i < m.inh_neu.size(); ++i) {
    neu_bdry[i] = m.inh_neu[i].bdry; neu_type[i] = m.inh_neu[i].type; neu_val[i] = m.inh_neu[i].val;
  }
}

// ---- spline / gauss helpers exposed for the known-answer tests ---------------------------------
void* orc_kv_from_list(const double* zs, int n, int poly, int curve) {
  ORC_TRY
  return new KnotVector(knot_vector_from_list(std::vector<double>(zs, zs + n), poly, curve));
  ORC_CATCH(nullptr)
}
void* orc_kv_uniform(int nel, int poly, int curve) {
  ORC_TRY
  return new KnotVector(knot_vector_uniform(nel, poly, curve));
  ORC_CATCH(nullptr)
}
void* orc_kv_of_mesh(void* h, int dir) { Mesh& m = *(Mesh*)h; return new KnotVector(dir == 1 ? m.kv1 : m.kv2); }
void orc_kv_destroy(void* kv) { delete (KnotVector*)kv; }
int orc_kv_len(void* kv) { return (int)((KnotVector*)kv)->zs.size(); }
int orc_kv_nel(void* kv) { return ((KnotVector*)kv)->nel; }
void orc_kv_knots(void* kv, double* o) { auto& z = ((KnotVector*)kv)->zs; std::copy(z.begin(), z.end(), o); }
int orc_fine_zs(int nel, int poly, double* o) {
  ORC_TRY
  auto z = get_fine_zs(nel, poly);
  std::copy(z.begin(), z.end(), o);
  return 0;
  ORC_CATCH(1)
}
int orc_knot_span(void* kv, double z) {
  ORC_TRY
  return get_knot_span_index(*(KnotVector*)kv, z);
  ORC_CATCH(-1)
}
int orc_bspline_vals(void* kv, double z, double* o) {
  ORC_TRY
  get_bspline_vals(*(KnotVector*)kv, z, o);
  return 0;
  ORC_CATCH(1)
}
int orc_bspline_ders(void* kv, double z, int nd, double* o) {
  ORC_TRY
  get_bspline_ders(*(KnotVector*)kv, z, nd, o);
  return 0;
  ORC_CATCH(1)
}
int orc_bspline_indices(void* kv, double z, int32_t* o) {
  ORC_TRY
  get_bspline_indices(*(KnotVector*)kv, z, o);
  return 0;
  ORC_CATCH(1)
}
int orc_collocate(void* kv, double* o) {
  ORC_TRY
  auto z = collocate_zeta(*(KnotVector*)kv);
  std::copy(z.begin(), z.end(), o);
  return (int)z.size();
  ORC_CATCH(-1)
}
int orc_cps_1d(void* kv, const double* xvals, int n, double* o) {
  ORC_TRY
  auto c = get_1d_bspline_cps(*(KnotVector*)kv, std::vector<double>(xvals, xvals + n));
  std::copy(c.begin(), c.end(), o);
  return 0;
  ORC_CATCH(1)
}
int orc_cps_2d(void* kv1, void* kv2, const double* xvals, int n, double* o) {
  ORC_TRY
  auto c = get_2d_bspline_cps(*(KnotVector*)kv1, *(KnotVector*)kv2, std::vector<double>(xvals, xvals + n));
  std::copy(c.begin(), c.end(), o);
  return 0;
  ORC_CATCH(1)
}
// returns uel_num; uel_ids (nel), uel_list (2 x uel_num)
int orc_unique_1d(void* kv, int64_t* uel_ids, double* uel_list) {
  ORC_TRY
  Unique1D u = get_unique_1d_elements(*(KnotVector*)kv);
  for (int i = 0; i < u.num_el; ++i) uel_ids[i] = u.uel_ids[i];
  for (int i = 0; i < u.uel_num; ++i) { uel_list[2 * i] = u.uel_list[i].first; uel_list[2 * i + 1] = u.uel_list[i].second; }
  return u.uel_num;
  ORC_CATCH(-1)
}
int orc_gauss_xi(int ngp, double* xs, double* ws) {
  ORC_TRY
  gauss_xi(ngp, xs, ws);
  return 0;
  ORC_CATCH(1)
}
int orc_gauss_zeta(int ngp, double lo, double hi, double* zs, double* ws) {
  ORC_TRY
  gauss_zeta(ngp, lo, hi, zs, ws);
  return 0;
  ORC_CATCH(1)
}
// line basis fns for an arbitrary knot vector: returns nuel; uel_ids(nel), tab(nuel*ngp*10)
int orc_line_fns(void* kv, int ngp, int64_t* uel_ids, double* tab) {
  ORC_TRY
  LineFns L = line_gp_basis_fns(*(KnotVector*)kv, ngp);
  for (int i = 0; i < L.nel; ++i) uel_ids[i] = L.uel_ids[i];
  for (size_t k = 0; k < L.ufns.size(); ++k) {
    double* o = tab + 10 * k;
    o[0] = L.ufns[k].w;
    for (int i = 0; i < 3; ++i) { o[1 + i] = L.ufns[k].N[i]; o[4 + i] = L.ufns[k].dN[i]; o[7 + i] = L.ufns[k].ddN[i]; }
  }
  return L.nuel;
  ORC_CATCH(-1)
}
int orc_fn1(void* kv, double w, double z, double* o10) {
  ORC_TRY
  Fn1 f = gp_basis_fns_1d(w, z, *(KnotVector*)kv);
  o10[0] = f.w;
  for (int i = 0; i < 3; ++i) { o10[1 + i] = f.N[i]; o10[4 + i] = f.dN[i]; o10[7 + i] = f.ddN[i]; }
  return 0;
  ORC_CATCH(1)
}
void orc_fn2(const double* a10, const double* b10, double* o55) {
  Fn1 f1, f2;
  f1.w = a10[0]; f2.w = b10[0];
  for (int i = 0; i < 3; ++i) {
    f1.N[i] = a10[1 + i]; f1.dN[i] = a10[4 + i]; f1.ddN[i] = a10[7 + i];
    f2.N[i] = b10[1 + i]; f2.dN[i] = b10[4 + i]; f2.ddN[i] = b10[7 + i];
  }
  put_fn2(gp_basis_fns_2d(f1, f2), o55);
}

// ---- Gauss-point kernel for the GeoDynStress known-answer tests ---------------------------------
// xms_el 9x3, cps_el 9 x ndf (real, col-major). out: x[3] a_[6: a1,a2] acon[4 col-major] aco[4] J n[3] b[4] H K
//   sig[3] sigm[3] M[3] lam pm v[3] vm[3]  -> 3+6+4+4+1+3+4+1+1+3+3+3+1+1+3+3 = 44 doubles
void orc_geo_dyn_stress(void* h, int64_t el, int gp, const double* xms_el, const double* cps_el, double* o) {
  Mesh& m = *(Mesh*)h;
  std::vector<cd> xe(27), ce((size_t)9 * m.ndf);
  for (int i = 0; i < 27; ++i) xe[i] = xms_el[i];
  for (int i = 0; i < 9 * m.ndf; ++i) ce[i] = cps_el[i];
  const Fn2& f = area_fns(m, el, gp);
  GDS g;
  geo_dyn_stress(xe.data(), ce.data(), m.dofs, f.N, f.dN, f.ddN, m.p.kb, m.p.kg, m.p.zv, g);
  int k = 0;
  for (int i = 0; i < 3; ++i) o[k++] = (double)g.x[i].real();
  for (int al = 0; al < 2; ++al) for (int i = 0; i < 3; ++i) o[k++] = (double)g.a_[al][i].real();
  for (int be = 0; be < 2; ++be) for (int al = 0; al < 2; ++al) o[k++] = (double)g.acon[al][be].real();
  for (int be = 0; be < 2; ++be) for (int al = 0; al < 2; ++al) o[k++] = (double)g.aco[al][be].real();
  o[k++] = (double)g.J.real();
  for (int i = 0; i < 3; ++i) o[k++] = (double)g.n[i].real();
  for (int be = 0; be < 2; ++be) for (int al = 0; al < 2; ++al) o[k++] = (double)g.b[al][be].real();
  o[k++] = (double)g.H.real(); o[k++] = (double)g.K.real();
  for (int i = 0; i < 3; ++i) o[k++] = (double)g.sig[i].real();
  for (int i = 0; i < 3; ++i) o[k++] = (double)g.sigm[i].real();
  for (int i = 0; i < 3; ++i) o[k++] = (double)g.M[i].real();
  o[k++] = (double)g.lam.real(); o[k++] = (double)g.pm.real();
  for (int i = 0; i < 3; ++i) o[k++] = (double)g.v[i].real();
  for (int i = 0; i < 3; ++i) o[k++] = (double)g.vm[i].real();
}

// element-level r_el / K_el (area element), for fine-grained parity tests. K_el col-major (9ndf)^2.
int orc_elem_r_K(void* h, int64_t el, const double* xms, const double* cps, double dt, double* r_el, double* K_el) {
  ORC_TRY
  Mesh& m = *(Mesh*)h;
  int mmo[3];
  get_m_motion_order(m.p.motion, m.dofs, mmo);
  std::vector<real_t> r, K;
  elem_r_K(m, el, xms, cps, dt, mmo,
           [&](const cd* xe, const cd* ce, cd* out) { calc_elem_residual(m, el, xe, ce, out); }, r, K);
  for (size_t i = 0; i < r.size(); ++i) r_el[i] = (double)r[i];
  for (size_t i = 0; i < K.size(); ++i) K_el[i] = (double)K[i];
  return 0;
  ORC_CATCH(1)
}
// same plus the error-bound scale E of every entry (zeros in the regular build); with bdry > 0 the Neumann boundary
// element (bdry code, Neumann type, value) of element `el` instead of the area element (FiniteElement.jl:156-184)
int orc_elem_r_K_mag(void* h, int64_t el, int bdry, int ntype, double nval, const double* xms, const double* cps,
                     double time, double dt, double* r_el, double* K_el, double* r_mag, double* K_mag) {
  ORC_TRY
  Mesh& m = *(Mesh*)h;
  int mmo[3];
  get_m_motion_order(m.p.motion, m.dofs, mmo);
  std::vector<real_t> r, K, rm, Km;
  if (bdry > 0)
    elem_r_K(m, el, xms, cps, dt, mmo,
             [&](const cd* xe, const cd* ce, cd* out) {
               calc_bdry_element_residual(m, bdry, ntype, nval, el, xe, ce, time, out);
             }, r, K, &rm, &Km);
  else
    elem_r_K(m, el, xms, cps, dt, mmo,
             [&](const cd* xe, const cd* ce, cd* out) { calc_elem_residual(m, el, xe, ce, out); }, r, K, &rm, &Km);
  for (size_t i = 0; i < r.size(); ++i) { r_el[i] = (double)r[i]; if (r_mag) r_mag[i] = (double)rm[i]; }
  for (size_t i = 0; i < K.size(); ++i) { K_el[i] = (double)K[i]; if (K_mag) K_mag[i] = (double)Km[i]; }
  return 0;
  ORC_CATCH(1)
}
// 0: regular oracle (double); 1: truth build in long double; 2: truth build in __float128
int orc_truth_kind() {
#if defined(ORC_TRUTH)
  return ORC_TRUTH;
#else
  return 0;
#endif
}
// replace the mesh's inhomogeneous Neumann conditions (tests of the SHEAR / TOP-BOTTOM MOMENT branches,
// FiniteElement.jl:374-380, which no scenario of the reference's Bc.jl exercises)
int orc_mesh_set_neumann(void* h, int n, const int32_t* bdry, const int32_t* type, const double* val) {
  ORC_TRY
  Mesh& m = *(Mesh*)h;
  m.inh_neu.clear();
  for (int k = 0; k < n; ++k) {
    ORC_CHECK(bdry[k] >= 1 && bdry[k] <= 4 && type[k] >= 1 && type[k] <= 3, "bad Neumann condition");
    m.inh_neu.push_back(NeuBc{bdry[k], type[k], val[k]});
  }
  return 0;
  ORC_CATCH(1)
}
// rv(27) of calc_elem_dof_residuals (used by calc_pull_force, PullForce.jl:75) -- real parts
int orc_elem_dof_residuals(void* h, int64_t el, const double* xms, const double* cps, double* rv, double* rm,
                           double* rl, double* rp) {
  ORC_TRY
  Mesh& m = *(Mesh*)h;
  std::vector<cd> xe(27), ce((size_t)9 * m.ndf);
  for (int a = 0; a < 9; ++a) {
    int64_t nodeid = m.IX[(size_t)a + (size_t)NEN * (el - 1)] - 1;
    for (int i = 0; i < 3; ++i) xe[a + 9 * i] = xms[(size_t)nodeid + (size_t)m.numnp * i];
    for (int d = 0; d < m.ndf; ++d) ce[a + 9 * d] = cps[(size_t)nodeid + (size_t)m.numnp * d];
  }
  cd v[27], mm[27], l[9], pp[9];
  calc_elem_dof_residuals(m, el, xe.data(), ce.data(), v, mm, l, pp);
  for (int i = 0; i < 27; ++i) { rv[i] = (double)v[i].real(); rm[i] = (double)mm[i].real(); }
  for (int i = 0; i < 9; ++i) { rl[i] = (double)l[i].real(); rp[i] = (double)pp[i].real(); }
  return 0;
  ORC_CATCH(1)
}

void* orc_calc_r_K(void* h, const double* xms, const double* cps, double time, double dt, int nthreads) {
  ORC_TRY
  return calc_r_K(*(Mesh*)h, xms, cps, time, dt, nthreads);
  ORC_CATCH(nullptr)
}
int64_t orc_result_nnz(void* r) { return (int64_t)((Result*)r)->nzval.size(); }
void orc_result_get(void* r, double* rvec, int64_t* colptr, int64_t* rowval, double* nzval) {
  Result& R = *(Result*)r;
  std::copy(R.r.begin(), R.r.end(), rvec);
  std::copy(R.colptr.begin(), R.colptr.end(), colptr);
  std::copy(R.rowval.begin(), R.rowval.end(), rowval);
  std::copy(R.nzval.begin(), R.nzval.end(), nzval);
}
void orc_result_destroy(void* r) { delete (Result*)r; }

int orc_calc_r_K_fast(void* h, const double* xms, const double* cps, double time, double dt, int nthreads,
                      const int64_t* colptr0, const int64_t* rowval0, int64_t nnz, int64_t e_first, int64_t e_last,
                      int with_neumann, double* r_out, double* nz_out) {
  ORC_TRY
  calc_r_K_fast(*(Mesh*)h, xms, cps, time, dt, nthreads, colptr0, rowval0, nnz, e_first, e_last, with_neumann, r_out,
                nz_out);
  return 0;
  ORC_CATCH(1)
}
// same, plus the per-entry error-bound scale E on the same pattern (truth builds; zeros otherwise)
int orc_calc_r_K_fast_mag(void* h, const double* xms, const double* cps, double time, double dt, int nthreads,
                          const int64_t* colptr0, const int64_t* rowval0, int64_t nnz, int64_t e_first,
                          int64_t e_last, int with_neumann, double* r_out, double* nz_out, double* rmag_out,
                          double* nzmag_out) {
  ORC_TRY
  calc_r_K_fast(*(Mesh*)h, xms, cps, time, dt, nthreads, colptr0, rowval0, nnz, e_first, e_last, with_neumann, r_out,
                nz_out, rmag_out, nzmag_out);
  return 0;
  ORC_CATCH(1)
}

}  // extern "C"
