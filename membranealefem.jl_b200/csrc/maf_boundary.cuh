// maf_boundary.cuh -- inhomogeneous Neumann boundary elements of calc_r_K (FiniteElement.jl:151-197,
// calc_bdry_element_residual :338-400, calc_tau_nu :431-452). One warp-sized group of lanes per
// (condition, boundary element). Only velocity rows are filled (:388-392) and the only non-zero tangent columns
// are those of the dofs that move the mesh (the boundary traction depends on x alone; the reference's
// cps perturbations there are identically zero, :170-173).
#pragma once
#include "maf_element.cuh"

namespace maf {

struct BoundaryTables {
  const double* edge1;     // 2 x 10: zeta-min / zeta-max basis of direction 1 (w = 1), GpBasisFn.jl:190-196
  const double* edge2;     // 2 x 10
  const int32_t* elems;    // concatenated boundary element lists (0-based element ids), per condition
  const int32_t* offs;     // n_neu + 1 offsets into elems
  const int32_t* bdry;     // Boundary code per condition
  const int32_t* ntype;    // Neumann code per condition
  const double* nval;      // value per condition
  int n_neu;
};

// shared-memory layout per group (doubles)
enum { B_X = 0, B_PHI = 27, B_W = 108, B_S = 111, B_D = 138, B_INT = 300, B_DOUBLES = 300 + 86 };

// writes rv contributions of one boundary element. `plain` = non-atomic read-modify-write (deterministic path,
// the caller guarantees that concurrently processed boundary elements share no node).
MAF_HD void boundary_gather(int lane, int nl, const Config& cfg, const Tables& T, const BoundaryTables& BT, int bc,
                            int64_t el, const double* xms, double* sm) {
  int32_t* si = reinterpret_cast<int32_t*>(sm + B_INT);
  const int bd = BT.bdry[bc];
  const int e1 = (int)(el % T.num1el), e2 = (int)(el / T.num1el);
  for (int k = lane; k < 27; k += nl) {
    const int a = k % 9, q = k / 9;
    sm[B_X + 9 * q + a] = xms[(int64_t)T.IX[9 * el + a] + T.numnp * q];
  }
  for (int a = lane; a < 9; a += nl) {
    const int64_t node = T.IX[9 * el + a];
    si[I_NODE + a] = (int32_t)node;
    si[I_MASK + a] = T.nodemask[node];
    for (int d = 0; d < 8; ++d) si[I_EQ + 8 * a + d] = d < cfg.ndf ? T.ID[(int64_t)cfg.ndf * node + d] : -1;
    for (int b = 0; b < 9; ++b) si[I_PAIR + 9 * a + b] = T.elpair[81 * (el - T.el0) + 9 * a + b];
  }
  // boundary basis (Mesh.jl:220-225, GpBasisFn.jl:271-273): BOTTOM/TOP = line1 x edge2, RIGHT/LEFT = edge1 x line2
  for (int k = lane; k < 81; k += nl) {
    const int a = k % 9, c = (k / 9) % 3, gp = k / 27;
    const int a1 = a % 3, a2 = a / 3;
    const int o1 = c == CH_N1 ? 1 : 0, o2 = c == CH_N2 ? 1 : 0;
    const double* f1;
    const double* f2;
    if (bd == 1 || bd == 3) { f1 = T.line1 + 30 * T.uel1[e1] + 10 * gp; f2 = BT.edge2 + (bd == 3 ? 10 : 0); }
    else { f1 = BT.edge1 + (bd == 2 ? 10 : 0); f2 = T.line2 + 30 * T.uel2[e2] + 10 * gp; }
    sm[B_PHI + k] = f1[1 + 3 * o1 + a1] * f2[1 + 3 * o2 + a2];
    if (a == 0 && c == 0) sm[B_W + gp] = f1[0] * f2[0];
  }
}

MAF_HD void boundary_gauss(int lane, int nl, const Config& cfg, const BoundaryTables& BT, int bc, double fval,
                           double dt, double* sm) {
  // 3 gp x (1 primal + 6 directions)
  for (int k = lane; k < 21; k += nl) {
    const int gp = k / 7, dir = k % 7;
    double a[2][3];
    for (int al = 0; al < 2; ++al)
      for (int i = 0; i < 3; ++i) {
        double s = 0.0;
        for (int n = 0; n < 9; ++n) s += sm[B_X + 9 * i + n] * sm[B_PHI + 27 * gp + 9 * (CH_N1 + al) + n];
        a[al][i] = s;
      }
    if (dir == 0) {
      double Sb[3][3];
      bdry_eval<double>(a, BT.bdry[bc], BT.ntype[bc], fval, Sb);
      for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 3; ++i) sm[B_S + 9 * gp + 3 * c + i] = Sb[c][i];
    } else {
      const int gam = (dir - 1) / 3, j = (dir - 1) % 3;
      Dual ad[2][3];
      for (int al = 0; al < 2; ++al)
        for (int i = 0; i < 3; ++i) ad[al][i] = Dual(a[al][i], (al == gam && i == j) ? dt : 0.0);
      Dual Sb[3][3];
      bdry_eval<Dual>(ad, BT.bdry[bc], BT.ntype[bc], fval, Sb);
      for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 3; ++i) sm[B_D + 54 * gp + 18 * c + 6 * i + 3 * gam + j] = Sb[c][i].d;
    }
  }
}

MAF_HD void boundary_scatter(int lane, int nl, const Config& cfg, const Tables& T, double* sm, double* r_gl,
                             double* nzval, bool plain) {
  const int32_t* si = reinterpret_cast<const int32_t*>(sm + B_INT);
  // residual rows (a, v_i)
  for (int k = lane; k < 27; k += nl) {
    const int a = k % 9, i = k / 9;
    const int I = cfg.fdof[F_V][i];
    if (I < 0) continue;
    const int eq = si[I_EQ + 8 * a + I];
    if (eq < 0) continue;
    double s = 0.0;
    for (int gp = 0; gp < 3; ++gp) {
      double t = 0.0;
      for (int c = 0; c < 3; ++c) t += sm[B_S + 9 * gp + 3 * c + i] * sm[B_PHI + 27 * gp + 9 * c + a];
      s += sm[B_W + gp] * t;
    }
    if (plain) r_gl[eq] += s; else atomic_add(&r_gl[eq], s);
  }
  if (cfg.mesh_field < 0) return;
  // tangent: rows (a, v_i), columns (b, mesh dof j)
  for (int k = lane; k < 729; k += nl) {
    const int b = k % 9, a = (k / 9) % 9, j = (k / 81) % 3, i = k / 243;
    const int I = cfg.fdof[F_V][i], J = cfg.fdof[cfg.mesh_field][j];
    if (I < 0 || J < 0) continue;
    const int eqc = si[I_EQ + 8 * b + J];
    const unsigned m = (unsigned)si[I_MASK + a];
    if (eqc < 0 || !((m >> I) & 1u)) continue;
    double s = 0.0;
    for (int gp = 0; gp < 3; ++gp) {
      double t = 0.0;
      for (int c = 0; c < 3; ++c) {
        const double* D = sm + B_D + 54 * gp + 18 * c + 6 * i;
        t += sm[B_PHI + 27 * gp + 9 * c + a] *
             (D[j] * sm[B_PHI + 27 * gp + 9 * CH_N1 + b] + D[3 + j] * sm[B_PHI + 27 * gp + 9 * CH_N2 + b]);
      }
      s += sm[B_W + gp] * t;
    }
    const int64_t slot = T.colptr[eqc] + T.pairoff[(int64_t)si[I_PAIR + 9 * a + b] * 8 + J] +
                         popc8(m & cfg.rowmask[J] & ((1u << I) - 1u));
    if (plain) nzval[slot] += s; else atomic_add(&nzval[slot], s);
  }
}

}  // namespace maf
