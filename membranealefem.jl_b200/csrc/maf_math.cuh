// maf_math.cuh -- Gauss-point physics of the membrane equations with exact (forward-mode) derivatives.
//
// One templated function, gp_eval, evaluates the "generalised stresses" S at a Gauss point: the coefficients
// that multiply the test-function channels {N, N_1, N_2, N_11, N_22, N_12} in the element residual of
//   /root/reference/src/analysis/FiniteElement.jl:284-313  (built there from GeoDynStress.jl:111-174).
// Instantiated with plain doubles it gives the residual; instantiated with Dual numbers on a subset of its
// inputs it gives one column of the Gauss-point tangent A = dS/dE exactly (this replaces the reference's
// complex-step differentiation, FiniteElement.jl:113-122, whose result equals the exact derivative to round-off).
//
// Everything here is __host__ __device__ so that tests can run the very same code on the CPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MAF_HD __host__ __device__ __forceinline__
#else
#define MAF_HD inline
#endif

namespace maf {

// ---- forward-mode dual number (value + one directional derivative) ----------------------------------------
struct Dual {
  double v, d;
  MAF_HD Dual() {}
  MAF_HD Dual(double v_) : v(v_), d(0.0) {}
  MAF_HD Dual(double v_, double d_) : v(v_), d(d_) {}
};
MAF_HD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
MAF_HD Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
MAF_HD Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
MAF_HD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
MAF_HD Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
MAF_HD Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
MAF_HD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
MAF_HD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.v * b.d + a.d * b.v); }
MAF_HD Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
MAF_HD Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
MAF_HD Dual operator/(Dual a, Dual b) {
  double q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
MAF_HD Dual operator/(double a, Dual b) {
  double q = a / b.v;
  return Dual(q, -q * b.d / b.v);
}
MAF_HD Dual operator/(Dual a, double b) { return Dual(a.v / b, a.d / b); }
MAF_HD Dual dsqrt(Dual a) {
  double s = sqrt(a.v);
  return Dual(s, 0.5 * a.d / s);
}
MAF_HD double dsqrt(double a) { return sqrt(a); }
MAF_HD double val(double a) { return a; }
MAF_HD double val(Dual a) { return a.v; }
MAF_HD double der(double) { return 0.0; }
MAF_HD double der(Dual a) { return a.d; }

template <class A, class B> struct Prom { typedef Dual T; };
template <> struct Prom<double, double> { typedef double T; };

// motion codes = reference's enum Motion (src/input/Enums.jl:67-73)
enum { M_STATIC = 1, M_EUL = 2, M_LAG = 3, M_ALEV = 4, M_ALEVB = 5 };
// test-function / trial-function channels
enum { CH_N = 0, CH_N1 = 1, CH_N2 = 2, CH_N11 = 3, CH_N22 = 4, CH_N12 = 5, NCH = 6 };

struct Material {
  double kb, kg, zv, pn, adb, am;
  double kdb;   // adb / zv, formed once on the host (an FP64 division per Gauss-point evaluation otherwise)
};

// Generalised stresses at one Gauss point. Sv/Sm: [channel][component]; Sl, Sp scalars (channel N only).
//   rv[a,i] = sum_gp w * sum_c Sv[c][i] * Phi^c_a      (FiniteElement.jl:293-297)
//   rm[a,i] = sum_gp w * sum_c Sm[c][i] * Phi^c_a      (:300-309)
//   rl[a]   = sum_gp w * Sl * N_a  (+ Dohrmann-Bochev, :323-324)          (:298-299)
//   rp[a]   = sum_gp w * Sp * N_a  (+ Dohrmann-Bochev, :325-327)          (:311-312)
template <class T> struct GpStress {
  T Sv[NCH][3];
  T Sm[NCH][3];
  T Sl, Sp;
};

// 1/x, sqrt(x), 1/sqrt(x) of the metric determinant with ONE division and one square root (FP64 division and square
// root are long instruction sequences): 1/sqrt(x) = sqrt(x) * (1/x); the derivative parts need no further division.
MAF_HD void inv_sqrt_inv(double x, double& ix, double& s, double& is) {
  ix = 1.0 / x;
  s = sqrt(x);
  is = s * ix;
}
MAF_HD void inv_sqrt_inv(Dual x, Dual& ix, Dual& s, Dual& is) {
  const double r = 1.0 / x.v, sq = sqrt(x.v), isq = sq * r;
  ix = Dual(r, -(r * r) * x.d);
  const double hd = 0.5 * x.d;
  s = Dual(sq, hd * isq);
  is = Dual(isq, -(hd * isq) * r);
}

// Metric quantities that depend on the tangent vectors only (GeoDynStress.jl:117-123).
template <class TA> struct GpGeom {
  TA A11, A12, A22;   // a^{alpha beta}
  TA idet, J, iJ;
  TA up[2][3];        // a^alpha
  TA n[3];            // unit normal
};

template <class TA> MAF_HD void gp_geom(const TA a[2][3], GpGeom<TA>& g) {
  TA a11 = a[0][0] * a[0][0] + a[0][1] * a[0][1] + a[0][2] * a[0][2];
  TA a12 = a[0][0] * a[1][0] + a[0][1] * a[1][1] + a[0][2] * a[1][2];
  TA a22 = a[1][0] * a[1][0] + a[1][1] * a[1][1] + a[1][2] * a[1][2];
  TA det = a11 * a22 - a12 * a12;
  inv_sqrt_inv(det, g.idet, g.J, g.iJ);
  g.A11 = a22 * g.idet;
  g.A22 = a11 * g.idet;
  g.A12 = -(a12 * g.idet);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g.up[0][i] = a[0][i] * g.A11 + a[1][i] * g.A12;
    g.up[1][i] = a[0][i] * g.A12 + a[1][i] * g.A22;
  }
  g.n[0] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * g.iJ;
  g.n[1] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * g.iJ;
  g.n[2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * g.iJ;
}

// Core of the Gauss-point evaluation. The second derivatives of x enter ONLY through the curvature components
// b[k] = x_{,k} . n (k = 11, 22, 12; GeoDynStress.jl:124-126) and the Christoffel symbols
// Gam[k][mu] = x_{,k} . a^mu (:121), which is what lets the tangent w.r.t. x_{,k} be assembled from three
// b-directions plus a closed-form Gamma term (maf_element.cuh, IT_GEO_B).
//   a[al][i] = x_{,al} (:112), dv[al][i] = v_{,al} (:136), v[i] (:135), dm[al][i] = vm_{,al} (:143), vm[i] (:142),
//   lam (:139), pm (:140).
template <int MOTION, class TA, class TB, class TGm, class TV, class TM, class TS, class TO>
MAF_HD void gp_core(const GpGeom<TA>& g, const TA a[2][3], const TB b[3], const TGm Gam[3][2], const TV dv[2][3],
                    const TV v[3], const TM dm[2][3], const TM vm[3], TS lam, TS pm, const Material& mat,
                    GpStress<TO>& out) {
  typedef typename Prom<TA, TB>::T TG;  // curvature-dependent geometry
  typedef typename Prom<TA, TV>::T TPV;
  typedef typename Prom<TA, TM>::T TPM;
  const TA A11 = g.A11, A12 = g.A12, A22 = g.A22, J = g.J;
  const TB b0 = b[0], b1 = b[1], b2 = b[2];
  // b^{alpha beta} = a^{..} b a^{..}  (:127)
  TG t11 = A11 * b0 + A12 * b2, t12 = A11 * b2 + A12 * b1;
  TG t21 = A12 * b0 + A22 * b2, t22 = A12 * b2 + A22 * b1;
  TG B11 = t11 * A11 + t12 * A12, B12 = t11 * A12 + t12 * A22, B22 = t21 * A12 + t22 * A22;
  TG H = 0.5 * (A11 * b0 + 2.0 * (A12 * b2) + A22 * b1);  // (:129)
  TG Kg = (b0 * b1 - b2 * b2) * g.idet;                   // (:130)

  // bending part of the in-plane stress and the moment (:150-154)
  TG bs = mat.kb * (H * H) - mat.kg * Kg;
  TG tkH = (2.0 * mat.kb) * H;
  TG sb11 = A11 * bs - tkH * B11, sb12 = A12 * bs - tkH * B12, sb22 = A22 * bs - tkH * B22;
  const double kM = mat.kb + 2.0 * mat.kg;
  TG Mt[3];  // Voigt (M11, M22, M12+M21)
  Mt[0] = A11 * H * kM - mat.kg * B11;
  Mt[1] = A22 * H * kM - mat.kg * B22;
  Mt[2] = 2.0 * (A12 * H * kM - mat.kg * B12);
  TG Q[3][3];  // J M~^k n_i : coefficient of N_{,k} (second derivatives)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    TG jm = J * Mt[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) Q[k][i] = jm * g.n[i];
  }

  // viscous stress pi^{ab} = zv (a^a . v_{,m} a^{mb} + a^b . v_{,m} a^{ma})  (:148-149)
  TPV g00 = g.up[0][0] * dv[0][0] + g.up[0][1] * dv[0][1] + g.up[0][2] * dv[0][2];
  TPV g01 = g.up[0][0] * dv[1][0] + g.up[0][1] * dv[1][1] + g.up[0][2] * dv[1][2];
  TPV g10 = g.up[1][0] * dv[0][0] + g.up[1][1] * dv[0][1] + g.up[1][2] * dv[0][2];
  TPV g11 = g.up[1][0] * dv[1][0] + g.up[1][1] * dv[1][1] + g.up[1][2] * dv[1][2];
  TPV p00 = g00 * A11 + g01 * A12, p01 = g00 * A12 + g01 * A22;
  TPV p10 = g10 * A11 + g11 * A12, p11 = g10 * A12 + g11 * A22;
  TPV pi11 = (2.0 * mat.zv) * p00, pi22 = (2.0 * mat.zv) * p11, pi12 = mat.zv * (p01 + p10);

  // total in-plane stress (:150)
  TO s11 = sb11 + A11 * lam + pi11;
  TO s22 = sb22 + A22 * lam + pi22;
  TO s12 = sb12 + A12 * lam + pi12;

  // ---- v rows (FiniteElement.jl:293-297) ----
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    TO qg0 = Q[0][i] * Gam[0][0] + Q[1][i] * Gam[1][0] + Q[2][i] * Gam[2][0];
    TO qg1 = Q[0][i] * Gam[0][1] + Q[1][i] * Gam[1][1] + Q[2][i] * Gam[2][1];
    out.Sv[CH_N1][i] = J * (s11 * a[0][i] + s12 * a[1][i]) - qg0;
    out.Sv[CH_N2][i] = J * (s12 * a[0][i] + s22 * a[1][i]) - qg1;
    out.Sv[CH_N11][i] = Q[0][i];
    out.Sv[CH_N22][i] = Q[1][i];
    out.Sv[CH_N12][i] = Q[2][i];
    out.Sv[CH_N][i] = (-mat.pn) * (J * g.n[i]);
  }
  // ---- lambda row (:298-299) ----
  out.Sl = J * (g00 + g11) - mat.kdb * lam;

  // ---- mesh rows ----
#pragma unroll
  for (int cch = 0; cch < NCH; ++cch)
#pragma unroll
    for (int i = 0; i < 3; ++i) out.Sm[cch][i] = TO(0.0);
  out.Sp = TO(0.0);
  if (MOTION == M_EUL) {  // (:300-303)
    TPV ndv = g.n[0] * v[0] + g.n[1] * v[1] + g.n[2] * v[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) out.Sm[CH_N][i] = mat.am * (J * (vm[i] - g.n[i] * ndv));
  } else if (MOTION == M_ALEV || MOTION == M_ALEVB) {  // (:304-313), sigma^m at GeoDynStress.jl:156-159
    TPM h00 = g.up[0][0] * dm[0][0] + g.up[0][1] * dm[0][1] + g.up[0][2] * dm[0][2];
    TPM h01 = g.up[0][0] * dm[1][0] + g.up[0][1] * dm[1][1] + g.up[0][2] * dm[1][2];
    TPM h10 = g.up[1][0] * dm[0][0] + g.up[1][1] * dm[0][1] + g.up[1][2] * dm[0][2];
    TPM h11 = g.up[1][0] * dm[1][0] + g.up[1][1] * dm[1][1] + g.up[1][2] * dm[1][2];
    TPM q00 = h00 * A11 + h01 * A12, q01 = h00 * A12 + h01 * A22;
    TPM q10 = h10 * A11 + h11 * A12, q11 = h10 * A12 + h11 * A22;
    TO m11 = sb11 + (2.0 * mat.zv) * q00;
    TO m22 = sb22 + (2.0 * mat.zv) * q11;
    TO m12 = sb12 + mat.zv * (q01 + q10);
    TO nd = TO(0.0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      out.Sm[CH_N1][i] = J * (m11 * a[0][i] + m12 * a[1][i]);
      out.Sm[CH_N2][i] = J * (m12 * a[0][i] + m22 * a[1][i]);
      if (MOTION == M_ALEVB) {
        TO qg0 = Q[0][i] * Gam[0][0] + Q[1][i] * Gam[1][0] + Q[2][i] * Gam[2][0];
        TO qg1 = Q[0][i] * Gam[0][1] + Q[1][i] * Gam[1][1] + Q[2][i] * Gam[2][1];
        out.Sm[CH_N1][i] = out.Sm[CH_N1][i] - qg0;
        out.Sm[CH_N2][i] = out.Sm[CH_N2][i] - qg1;
        out.Sm[CH_N11][i] = Q[0][i];
        out.Sm[CH_N22][i] = Q[1][i];
        out.Sm[CH_N12][i] = Q[2][i];
      }
      out.Sm[CH_N][i] = -(J * g.n[i]) * pm;
      nd = nd + g.n[i] * (vm[i] - v[i]);
    }
    out.Sp = -(J * nd) - mat.kdb * pm;
  }
}

// Tangent of the metric quantities in the direction  d a_gamma = dt e_j  (one Cartesian component of one tangent
// vector), in closed form from the primal ones -- what gp_geom<Dual> would produce, without its dual division and
// square root (a third of the FP64 work of a GEO_A item and its two longest dependent chains):
//   d det = 2 det dt a^gamma_j            d J = J dt a^gamma_j            d (1/J) = -(1/J) dt a^gamma_j
//   d a^{al be} = -dt (a^{al gamma} a^be_j + a^al_j a^{gamma be})
//   d n_i = -dt a^gamma_i n_j             d a^mu_i = dt (a^{mu gamma} n_i n_j - a^mu_j a^gamma_i)
MAF_HD void gp_geom_tangent(const GpGeom<double>& g, int gamma, int j, double dt, GpGeom<Dual>& gd) {
  // (gamma, j) are per-lane run-time values: selected with conditionals, never by indexing (an array indexed at run
  // time would be placed in local memory)
  const double A0g = gamma == 0 ? g.A11 : g.A12, A1g = gamma == 0 ? g.A12 : g.A22;   // a^{mu gamma}
  const double ug[3] = {gamma == 0 ? g.up[0][0] : g.up[1][0], gamma == 0 ? g.up[0][1] : g.up[1][1],
                        gamma == 0 ? g.up[0][2] : g.up[1][2]};                       // a^gamma_i
  const double u0 = dt * (j == 0 ? g.up[0][0] : (j == 1 ? g.up[0][1] : g.up[0][2]));   // dt a^1_j
  const double u1 = dt * (j == 0 ? g.up[1][0] : (j == 1 ? g.up[1][1] : g.up[1][2]));   // dt a^2_j
  const double nj = dt * (j == 0 ? g.n[0] : (j == 1 ? g.n[1] : g.n[2]));
  const double u = gamma == 0 ? u0 : u1;                                               // dt a^gamma_j
  gd.J = Dual(g.J, g.J * u);
  gd.iJ = Dual(g.iJ, -(g.iJ * u));
  gd.idet = Dual(g.idet, -2.0 * (g.idet * u));
  gd.A11 = Dual(g.A11, -2.0 * (A0g * u0));
  gd.A22 = Dual(g.A22, -2.0 * (A1g * u1));
  gd.A12 = Dual(g.A12, -(A0g * u1 + u0 * A1g));
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gd.n[i] = Dual(g.n[i], -(ug[i] * nj));
    const double nn = g.n[i] * nj;
    gd.up[0][i] = Dual(g.up[0][i], A0g * nn - u0 * ug[i]);
    gd.up[1][i] = Dual(g.up[1][i], A1g * nn - u1 * ug[i]);
  }
}

// Full evaluation from the interpolated fields: c[k][i] = x_{,11} x_{,22} x_{,12} (GeoDynStress.jl:115); the metric
// quantities g belong to a (gp_geom, or gp_geom_tangent for a seeded direction).
template <int MOTION, class TA, class TC, class TV, class TM, class TS>
MAF_HD void gp_eval_geom(const GpGeom<TA>& g, const TA a[2][3], const TC c[3][3], const TV dv[2][3], const TV v[3],
                         const TM dm[2][3], const TM vm[3], TS lam, TS pm, const Material& mat,
                         GpStress<typename Prom<typename Prom<typename Prom<TA, TC>::T, typename Prom<TV, TM>::T>::T, TS>::T>& out) {
  typedef typename Prom<TA, TC>::T TG;
  TG b[3], Gam[3][2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    b[k] = c[k][0] * g.n[0] + c[k][1] * g.n[1] + c[k][2] * g.n[2];
#pragma unroll
    for (int mu = 0; mu < 2; ++mu) Gam[k][mu] = c[k][0] * g.up[mu][0] + c[k][1] * g.up[mu][1] + c[k][2] * g.up[mu][2];
  }
  gp_core<MOTION>(g, a, b, Gam, dv, v, dm, vm, lam, pm, mat, out);
}

template <int MOTION, class TA, class TC, class TV, class TM, class TS>
MAF_HD void gp_eval(const TA a[2][3], const TC c[3][3], const TV dv[2][3], const TV v[3], const TM dm[2][3],
                    const TM vm[3], TS lam, TS pm, const Material& mat,
                    GpStress<typename Prom<typename Prom<typename Prom<TA, TC>::T, typename Prom<TV, TM>::T>::T, TS>::T>& out) {
  typedef typename Prom<TA, TC>::T TG;
  GpGeom<TA> g;
  gp_geom(a, g);
  TG b[3], Gam[3][2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    b[k] = c[k][0] * g.n[0] + c[k][1] * g.n[1] + c[k][2] * g.n[2];
#pragma unroll
    for (int mu = 0; mu < 2; ++mu) Gam[k][mu] = c[k][0] * g.up[mu][0] + c[k][1] * g.up[mu][1] + c[k][2] * g.up[mu][2];
  }
  gp_core<MOTION>(g, a, b, Gam, dv, v, dm, vm, lam, pm, mat, out);
}

// ---- Neumann boundary Gauss point (FiniteElement.jl:363-384, calc_tau_nu :431-452) -------------------------
// Sb[c][i], c in {N, N_1, N_2}: rv[a,i] += w * sum_c Sb[c][i] Phi^c_a. Only the tangent vectors enter.
template <class TA>
MAF_HD void bdry_eval(const TA a[2][3], int bdry, int ntype, double fval /* nval or Mval */, TA Sb[3][3]) {
  TA a11 = a[0][0] * a[0][0] + a[0][1] * a[0][1] + a[0][2] * a[0][2];
  TA a12 = a[0][0] * a[1][0] + a[0][1] * a[1][1] + a[0][2] * a[1][2];
  TA a22 = a[1][0] * a[1][0] + a[1][1] * a[1][1] + a[1][2] * a[1][2];
  TA det = a11 * a22 - a12 * a12;
  TA idet = 1.0 / det;
  TA A11 = a22 * idet, A22 = a11 * idet, A12 = -(a12 * idet);
  TA iJ = 1.0 / dsqrt(det);
  TA up[2][3], n[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    up[0][i] = a[0][i] * A11 + a[1][i] * A12;
    up[1][i] = a[0][i] * A12 + a[1][i] * A22;
  }
  n[0] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * iJ;
  n[1] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * iJ;
  n[2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * iJ;
  // tau = +-a_1 or +-a_2, normalised: BOTTOM +a1, RIGHT +a2, TOP -a1, LEFT -a2 (:438-448)
  const int al = (bdry == 1 || bdry == 3) ? 0 : 1;
  const double sgn = (bdry == 1 || bdry == 2) ? 1.0 : -1.0;
  // The reference normalises tau with dot(tau, tau), which CONJUGATES (FiniteElement.jl:448): its |tau| carries no
  // complex-step derivative, whereas `inorm` below is differentiated. The results coincide because every Neumann type
  // multiplies by J_Gamma = 1 / |a^alpha . tau| (:372), which is homogeneous of degree -1 in tau: the scalar |tau|
  // cancels in nu J_Gamma, tau J_Gamma and nu^alpha J_Gamma (STRETCH, SHEAR, MOMENT), exactly and for any
  // perturbation. A future Neumann type that does not carry J_Gamma would need inorm's derivative frozen here; all
  // three existing types are compared with the oracle (which conjugates like the reference) on all four sides in
  // tests/test_truth_oracle.py and tests/test_gpu_parity.py::test_shear_and_top_bottom_moment.
  TA inorm = 1.0 / dsqrt(al == 0 ? a11 : a22);
  TA tau[3], nu[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) tau[i] = (sgn * a[al][i]) * inorm;
  nu[0] = tau[1] * n[2] - tau[2] * n[1];
  nu[1] = tau[2] * n[0] - tau[0] * n[2];
  nu[2] = tau[0] * n[1] - tau[1] * n[0];
  TA t0 = up[0][0] * tau[0] + up[0][1] * tau[1] + up[0][2] * tau[2];
  TA t1 = up[1][0] * tau[0] + up[1][1] * tau[1] + up[1][2] * tau[2];
  TA JG = 1.0 / dsqrt(t0 * t0 + t1 * t1);  // (:372)
#pragma unroll
  for (int cch = 0; cch < 3; ++cch)
#pragma unroll
    for (int i = 0; i < 3; ++i) Sb[cch][i] = TA(0.0);
  if (ntype == 2 || ntype == 1) {  // STRETCH (nu) / SHEAR (tau) (:374-376)
#pragma unroll
    for (int i = 0; i < 3; ++i) Sb[0][i] = -(fval * ((ntype == 2) ? nu[i] : tau[i])) * JG;
  } else {  // MOMENT (:377-380): rv -= n (dN . nu^alpha) Mval JG w
    TA nu0 = up[0][0] * nu[0] + up[0][1] * nu[1] + up[0][2] * nu[2];
    TA nu1 = up[1][0] * nu[0] + up[1][1] * nu[1] + up[1][2] * nu[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Sb[1][i] = -(fval * n[i]) * nu0 * JG;
      Sb[2][i] = -(fval * n[i]) * nu1 * JG;
    }
  }
}

}  // namespace maf
