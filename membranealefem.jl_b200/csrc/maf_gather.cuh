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
  // Which staged rows a node pair sums, precomputed (maf_host.h::build_pair_classes): pair p = (A, B) takes, in
  // ascending element id, row crow[9 c + q] of element eref[B] + cde[9 c + q], q < ccnt[c], c = pclass[p]. Pairs with
  // the same relative element offsets and local indices share a class (a few hundred on a structured patch). Without
  // it (pclass == NULL) the kernel scans the elements of B and looks A up in each: 99 index loads per pair against 27
  // loads of staged data, the bulk of the gather's time.
  const uint16_t* pclass;    // npairs, or NULL
  const int32_t* eref;       // numnp: first element of the node's element list
  const uint8_t* ccnt;       // classes
  const int32_t* cde;        // classes x 9
  const uint8_t* crow;       // classes x 9: 9 a + b
  int64_t ring;              // staging rows live in a ring of this many elements (bands of element rows are staged
                             // and gathered one after the other); 0: one row block per element of the range
};
MAF_HD int64_t stage_index(int64_t k, int64_t ring) { return ring > 0 ? k % ring : k; }

// one node pair p = (A,B): the <= 8 x 8 dof block K[(A,:),(B,:)], summed by MAF_GATHER_LANES threads.
// A staged row (element, a, b) holds the nij (row dof, col dof) classes contiguously, ordered by (J, I) -- the order
// of the destination slots -- padded to an even stride. Lane s of the pair owns the 16-byte chunks s, s + L, s + 2L, ...
// of the row: the L lanes of a pair read 16 L contiguous bytes per load (the kernel is bound by the number of L1
// accesses: with one thread per pair the 32 lanes of a warp read 32 different rows), keep <= 2 ceil(32 / L) sums in
// registers and write their own classes. (One whole warp per pair was slower: the per-pair index work is then done 32
// times.)
#define MAF_MAX_NIJ 64
#ifndef MAF_GATHER_LANES
#define MAF_GATHER_LANES 8
#endif
MAF_HD void gather_K_pair(int64_t p, int s, const Config& cfg, const Tables& T, const GatherTables& G,
                          const double* kel, int nij, int64_t e0, int64_t e1, double* nzval) {
  constexpr int L = MAF_GATHER_LANES, NQ = (MAF_MAX_NIJ / 2 + L - 1) / L;
  const int ndf = cfg.ndf;
  const int32_t A = G.nbr[p], B = G.pair_node[p];
  const unsigned mA = T.nodemask[A], mB = T.nodemask[B];
  if (mA == 0 || mB == 0) return;
  double acc[2 * NQ];
#pragma unroll
  for (int c = 0; c < 2 * NQ; ++c) acc[c] = 0.0;
  // elements that contain both nodes, in ascending element id
  const bool tab = G.pclass != nullptr;
  const int cls9 = tab ? 9 * (int)G.pclass[p] : 0;
  const int64_t q0 = tab ? 0 : G.n2e_ptr[B], q1 = tab ? (int64_t)G.ccnt[cls9 / 9] : G.n2e_ptr[B + 1];
  const int64_t er = tab ? (int64_t)G.eref[B] : 0;
  for (int64_t q = q0; q < q1; ++q) {
    int64_t e;
    int rowid;
    if (tab) {
      e = er + G.cde[cls9 + q];
      if (e < e0 || e >= e1) continue;
      rowid = G.crow[cls9 + q];
    } else {
      e = G.n2e[q];
      if (e < e0 || e >= e1) continue;
      int a = -1;
#pragma unroll
      for (int k = 0; k < 9; ++k)
        if (T.IX[9 * e + k] == A) a = k;
      if (a < 0) continue;
      rowid = 9 * a + G.n2e_loc[q];
    }
    const double* row = kel + ((size_t)81 * stage_index(e - e0, G.ring) + rowid) * nij;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int c = 2 * (s + L * k);   // a padding entry is read but never written back
      if (c < nij) {
        const dbl2 v = ld2(row + c);
        acc[2 * k] += v.x;
        acc[2 * k + 1] += v.y;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NQ; ++k)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 2 * (s + L * k) + h;
      if (c >= G.ncls) continue;
      const int I = G.class_I[c], J = G.class_J[c];
      if (!((mB >> J) & 1u) || !((mA >> I) & 1u)) continue;
      const int64_t colbase = T.colptr[T.ID[(int64_t)ndf * B + J]] + T.pairoff[p * 8 + J];
      nzval[colbase + popc8(mA & cfg.rowmask[J] & ((1u << I) - 1u))] = acc[2 * k + h];
    }
  // P_sym pattern: rows of dof blocks that are identically zero are part of the pattern but own no class
  if (G.sym_fill && s == 0) {
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
    s += rel[72 * stage_index(e - e0, G.ring) + 9 * u + G.n2e_loc[q]];
  }
  r_gl[eq] = s;
}

}  // namespace maf
