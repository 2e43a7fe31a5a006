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
  uint8_t class_I[64], class_J[64];  // the classes in staging order (sorted by J, then I)
  int sym_fill;              // pattern holds rows without a class (MAF_PATTERN_SYM): they are written as zeros
  int ncls;                  // number of classes (the staging rows are padded to an even stride nij >= ncls)
  int64_t npairs;
};

// one node pair p = (A,B): the <= 8 x 8 dof block K[(A,:),(B,:)].
// A staged row (element, a, b) holds the nij (row dof, col dof) classes contiguously, ordered by (J, I) -- the order
// of the destination slots -- so the thread streams whole rows (16-byte loads) into registers and then writes the
// active entries of every column in one pass.
#define MAF_MAX_NIJ 64
MAF_HD void gather_K_pair(int64_t p, const Config& cfg, const Tables& T, const GatherTables& G, const double* kel,
                          int nij, int64_t e0, int64_t e1, double* nzval) {
  const int ndf = cfg.ndf;
  const int32_t A = G.nbr[p], B = G.pair_node[p];
  const unsigned mA = T.nodemask[A], mB = T.nodemask[B];
  if (mA == 0 || mB == 0) return;
  double acc[MAF_MAX_NIJ];
#pragma unroll
  for (int c = 0; c < MAF_MAX_NIJ; ++c) acc[c] = 0.0;
  // elements that contain both nodes, in ascending element id
  for (int64_t q = G.n2e_ptr[B]; q < G.n2e_ptr[B + 1]; ++q) {
    const int64_t e = G.n2e[q];
    if (e < e0 || e >= e1) continue;
    int a = -1;
#pragma unroll
    for (int k = 0; k < 9; ++k)
      if (T.IX[9 * e + k] == A) a = k;
    if (a < 0) continue;
    const double* row = kel + ((size_t)81 * (e - e0) + 9 * a + G.n2e_loc[q]) * nij;
    // the stride nij is even and the rows are 16-byte aligned: two classes per load (the kernel is bound by the
    // number of L1 accesses: 32 lanes read 32 different rows); a padding entry is read but never written back
#pragma unroll
    for (int c = 0; c < MAF_MAX_NIJ; c += 2)
      if (c < nij) {
        const dbl2 v = ld2(row + c);
        acc[c] += v.x;
        acc[c + 1] += v.y;
      }
  }
  // classes are sorted by (J, I): walk them once, keeping the first slot of node A's rows in the current column
  int64_t colbase = 0;
  int curJ = -1;
  bool colact = false;
#pragma unroll
  for (int c = 0; c < MAF_MAX_NIJ; ++c) {
    if (c >= G.ncls) break;
    const int I = G.class_I[c], J = G.class_J[c];
    if (J != curJ) {
      curJ = J;
      colact = (mB >> J) & 1u;
      if (colact) colbase = T.colptr[T.ID[(int64_t)ndf * B + J]] + T.pairoff[p * 8 + J];
    }
    if (colact && ((mA >> I) & 1u)) nzval[colbase + popc8(mA & cfg.rowmask[J] & ((1u << I) - 1u))] = acc[c];
  }
  // P_sym pattern: rows of dof blocks that are identically zero are part of the pattern but own no class
  if (G.sym_fill) {
    for (int J = 0; J < ndf; ++J) {
      if (!((mB >> J) & 1u)) continue;
      const unsigned rows = mA & cfg.rowmask[J];
      int64_t s0 = T.colptr[T.ID[(int64_t)ndf * B + J]] + T.pairoff[p * 8 + J];
      for (int I = 0; I < ndf; ++I) {
        if (!((rows >> I) & 1u)) continue;
        if (G.ij_of[8 * I + J] < 0) nzval[s0] = 0.0;
        ++s0;
      }
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
