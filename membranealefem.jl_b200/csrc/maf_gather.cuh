// maf_gather.cuh -- deterministic scatter path ("segmented reduction"): the area kernel stages every element's
// tangent block and residual; here every nnz slot / residual row sums its segment of staged contributions in
// ASCENDING ELEMENT ID -- the summation order of the reference's element loop with one Julia thread
// (FiniteElement.jl:98,129-136) -- and is written exactly once (no atomics, bitwise reproducible run to run).
#pragma once
#include "maf_element.cuh"

namespace maf {

struct GatherTables {
  const int64_t* nbr_ptr;    // numnp+1
  const int32_t* nbr;        // npairs: sorted neighbour nodes A of every node B
  const int32_t* pair_node;  // npairs: the column node B of the pair
  const int64_t* n2e_ptr;    // numnp+1
  const int32_t* n2e;        // element ids, ascending per node
  const uint8_t* n2e_loc;    // local node index of the node in that element
  const int16_t* ij_of;      // 8 x 8: column of the (row dof I, col dof J) class in a staging row, -1 = no block
  int64_t npairs;
};

// one node pair p = (A,B): the <= 8 x 8 dof block K[(A,:),(B,:)]
MAF_HD void gather_K_pair(int64_t p, const Config& cfg, const Tables& T, const GatherTables& G, const double* kel,
                          int nij, int64_t e0, int64_t e1, double* nzval) {
  const int ndf = cfg.ndf;
  const int32_t A = G.nbr[p], B = G.pair_node[p];
  const unsigned mA = T.nodemask[A], mB = T.nodemask[B];
  if (mA == 0 || mB == 0) return;
  // elements that contain both nodes, ascending, with the local indices (a, b)
  int ne = 0;
  int64_t els[16];
  int ab[16];
  for (int64_t q = G.n2e_ptr[B]; q < G.n2e_ptr[B + 1] && ne < 16; ++q) {
    const int64_t e = G.n2e[q];
    if (e < e0 || e >= e1) continue;
    int a = -1;
    for (int k = 0; k < 9; ++k)
      if (T.IX[9 * e + k] == A) a = k;
    if (a < 0) continue;
    els[ne] = e;
    ab[ne] = 9 * a + G.n2e_loc[q];
    ++ne;
  }
  for (int J = 0; J < ndf; ++J) {
    if (!((mB >> J) & 1u)) continue;
    const unsigned rows = mA & cfg.rowmask[J];
    if (!rows) continue;
    int64_t slot = T.colptr[T.ID[(int64_t)ndf * B + J]] + T.pairoff[p * 8 + J];
    for (int I = 0; I < ndf; ++I) {
      if (!((rows >> I) & 1u)) continue;
      const int c = G.ij_of[8 * I + J];
      double acc = 0.0;
      if (c >= 0)
        for (int k = 0; k < ne; ++k) acc += kel[((size_t)81 * (els[k] - e0) + ab[k]) * nij + c];
      nzval[slot++] = acc;
    }
  }
}

// one (node, dof) residual row
MAF_HD void gather_r_row(int64_t k, const Config& cfg, const Tables& T, const GatherTables& G, const double* rel,
                         int64_t e0, int64_t e1, double* r_gl) {
  const int ndf = cfg.ndf;
  const int64_t node = k / ndf;
  const int dof = (int)(k % ndf);
  const int eq = T.ID[k];
  if (eq < 0) return;
  int u = -1;  // slot of this dof in a staged element residual: v0 v1 v2 m0 m1 m2 l p
  for (int f = 0; f < NFIELD; ++f)
    for (int i = 0; i < cfg.ncomp[f]; ++i)
      if (cfg.fdof[f][i] == dof) u = (f == F_V ? 0 : (f == F_M ? 3 : (f == F_L ? 6 : 7))) + i;
  double s = 0.0;
  for (int64_t q = G.n2e_ptr[node]; q < G.n2e_ptr[node + 1]; ++q) {
    const int64_t e = G.n2e[q];
    if (e < e0 || e >= e1) continue;
    s += rel[72 * (e - e0) + 9 * u + G.n2e_loc[q]];
  }
  r_gl[eq] = s;
}

}  // namespace maf
