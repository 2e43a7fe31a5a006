// maf_state.cuh -- the Newton update on a device-resident state: what time_step! does between two
// calls of calc_r_K (FiniteElement.jl:41-46) and the predictor of run_analysis (Analysis.jl:70), so that xms / cps
// never travel back to the host inside a time step. Entry functions are __host__ __device__ like the element
// phases: tests/emu runs the same code on the CPU.
#pragma once
#include "maf_element.cuh"

namespace maf {

// x + dt * d with the product rounded before the sum, as the reference's broadcast `xms .+= dt * cps` does
// (no fused multiply-add: the update must be bit-identical to the host loop)
MAF_HD double add_scaled(double x, double dt, double d) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(x, __dmul_rn(dt, d));
#else
  volatile double inc = dt * d;
  return x + inc;
#endif
}

// dof (0-based) whose velocity moves position component mj, or -1 (get_m_motion_order, Mesh.jl:529-542)
MAF_HD int mesh_motion_dof(const Config& cfg, int mj) {
  return cfg.mesh_field >= 0 ? cfg.fdof[cfg.mesh_field][mj] : -1;
}

// Newton update of one (node, dof) entry k = dof + ndf * node:
//   dcps[ID_inv] = du; cps += dcps; update_xms!(xms, dcps, dt)     (FiniteElement.jl:41-46, 408-423)
MAF_HD void state_update_entry(int64_t k, const Config& cfg, const Tables& T, const double* du, double dt,
                               double* xms, double* cps) {
  const int32_t eq = T.ID[k];
  if (eq < 0) return;   // Dirichlet / absent: dcps = 0
  const int64_t node = k / cfg.ndf, np = T.numnp;
  const int dof = (int)(k % cfg.ndf);
  const double d = du[eq];
  cps[node + np * dof] += d;
#pragma unroll
  for (int mj = 0; mj < 3; ++mj)
    if (mesh_motion_dof(cfg, mj) == dof) xms[node + np * mj] = add_scaled(xms[node + np * mj], dt, d);
}

// predictor: the mesh velocity at time t is the first guess at t + dt, xms += dt * cps[:, mesh dofs] (Analysis.jl:70)
MAF_HD void state_predict_entry(int64_t k /* node + numnp * mj */, const Config& cfg, const Tables& T, double dt,
                                double* xms, const double* cps) {
  const int64_t np = T.numnp, node = k % np;
  const int dof = mesh_motion_dof(cfg, (int)(k / np));
  if (dof >= 0) xms[k] = add_scaled(xms[k], dt, cps[node + np * dof]);
}

}  // namespace maf
