// maf_host.h -- validation of the C-ABI inputs and construction of everything the kernels read
// (0-based int32 index tables, kernel configuration, symbolic pattern, Dohrmann-Bochev matrices, boundary lists).
// Pure C++; used by the library (which uploads the result) and by the CPU emulation harness in tests/.
#pragma once
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/maf.h"
#include "maf_boundary.cuh"
#include "maf_config.h"
#include "maf_gather.cuh"
#include "maf_symbolic.h"

namespace maf {

struct HostModel {
  Config cfg;
  Symbolic sym;
  int64_t numel = 0, numnp = 0, nmdf = 0;
  int ndf = 0, num1el = 0, num2el = 0, nuel1 = 0, nuel2 = 0, motion = 0, scenario = 0, pattern_mode = 0;
  std::vector<int32_t> IX0, ID0, uel1, uel2;
  std::vector<double> line1, line2, edge1, edge2, tdb, utab;
  std::vector<int64_t> nodecol;      // column pointers per (node, dof) (Tables)
  std::vector<int32_t> elslot;       // per-element scatter maps, HOST copy: only built for the CPU emulation of the
  std::vector<int64_t> elbase;       // kernels (tests/emu); the library builds them on the device (build_elslot_kernel)
  std::vector<int32_t> elclass;      // host copy: one class per element (the library merges equal maps on the device)
  std::vector<int32_t> nodemask32;
  std::vector<int32_t> b_elems, b_offs, b_bdry, b_type;
  std::vector<double> b_val;
  int n_neu = 0;
};

#define MAF_REQUIRE(cond, msg)                         \
  do {                                                 \
    if (!(cond)) throw std::runtime_error(std::string(msg)); \
  } while (0)

inline void build_host_model(HostModel& M, const maf_mesh_desc* d, const maf_params* p, int nthreads) {
  MAF_REQUIRE(d && p, "null mesh descriptor or params");
  MAF_REQUIRE(d->numel > 0 && d->numnp > 0, "empty mesh");
  MAF_REQUIRE(d->ndf >= 3 && d->ndf <= 8, "ndf must be in 3..8");
  MAF_REQUIRE(d->num1el > 0 && d->num2el > 0 && d->num1el * d->num2el == d->numel, "numel != num1el*num2el");
  MAF_REQUIRE(d->numnp < ((int64_t)1 << 31) && d->numel < ((int64_t)1 << 31) / 81, "mesh too large for int32 tables");
  MAF_REQUIRE(d->IX && d->ID && d->uel_ids1 && d->uel_ids2 && d->line1 && d->line2 && d->edge1 && d->edge2,
              "null table pointer");
  MAF_REQUIRE(p->motion >= 1 && p->motion <= 5, "unknown motion code");
  MAF_REQUIRE(p->zv != 0.0, "membrane viscosity must be non-zero");
  M.numel = d->numel; M.numnp = d->numnp; M.nmdf = d->nmdf; M.ndf = (int)d->ndf;
  M.num1el = (int)d->num1el; M.num2el = (int)d->num2el; M.nuel1 = (int)d->nuel1; M.nuel2 = (int)d->nuel2;
  M.motion = p->motion; M.scenario = p->scenario; M.pattern_mode = p->pattern_mode;
  for (int u = 0; u < 8; ++u) MAF_REQUIRE(d->dofs[u] >= 0 && d->dofs[u] <= d->ndf, "dofs entry outside 0..ndf");

  M.IX0.resize((size_t)9 * M.numel);
  for (size_t k = 0; k < M.IX0.size(); ++k) {
    MAF_REQUIRE(d->IX[k] >= 1 && d->IX[k] <= M.numnp, "IX entry outside 1..numnp");
    M.IX0[k] = (int32_t)(d->IX[k] - 1);
  }
  M.ID0.resize((size_t)M.ndf * M.numnp);
  int64_t maxeq = 0;
  for (size_t k = 0; k < M.ID0.size(); ++k) {
    MAF_REQUIRE(d->ID[k] >= 0 && d->ID[k] <= M.nmdf, "ID entry outside 0..nmdf");
    M.ID0[k] = (int32_t)(d->ID[k] - 1);
    maxeq = std::max(maxeq, d->ID[k]);
  }
  MAF_REQUIRE(maxeq == M.nmdf, "nmdf != maximum(ID)");
  // node-major numbering (Mesh.jl:276-284) is what makes the CSC rows come out sorted
  {
    int64_t next = 1;
    for (size_t k = 0; k < M.ID0.size(); ++k)
      if (d->ID[k] != 0) { MAF_REQUIRE(d->ID[k] == next, "ID is not numbered node-major (Mesh.jl:276-284)"); ++next; }
  }
  if (d->LM) {
    const int nd = 9 * M.ndf;
    for (int64_t e = 0; e < M.numel; ++e)
      for (int a = 0; a < 9; ++a)
        for (int q = 0; q < M.ndf; ++q)
          MAF_REQUIRE(d->LM[(size_t)q + (size_t)M.ndf * a + (size_t)nd * e] ==
                          d->ID[(size_t)q + (size_t)M.ndf * (d->IX[(size_t)a + 9 * (size_t)e] - 1)],
                      "LM != ID[:, IX] (Mesh.jl:299)");
  }
  M.uel1.resize(M.num1el);
  M.uel2.resize(M.num2el);
  for (int k = 0; k < M.num1el; ++k) {
    MAF_REQUIRE(d->uel_ids1[k] >= 1 && d->uel_ids1[k] <= M.nuel1, "uel_ids1 entry outside 1..nuel1");
    M.uel1[k] = (int32_t)(d->uel_ids1[k] - 1);
  }
  for (int k = 0; k < M.num2el; ++k) {
    MAF_REQUIRE(d->uel_ids2[k] >= 1 && d->uel_ids2[k] <= M.nuel2, "uel_ids2 entry outside 1..nuel2");
    M.uel2[k] = (int32_t)(d->uel_ids2[k] - 1);
  }
  M.line1.assign(d->line1, d->line1 + (size_t)30 * M.nuel1);
  M.line2.assign(d->line2, d->line2 + (size_t)30 * M.nuel2);
  M.edge1.assign(d->edge1, d->edge1 + 20);
  M.edge2.assign(d->edge2, d->edge2 + 20);
  for (double v : M.line1) MAF_REQUIRE(std::isfinite(v), "non-finite basis table entry");
  for (double v : M.line2) MAF_REQUIRE(std::isfinite(v), "non-finite basis table entry");

  build_config(M.cfg, p->motion, M.ndf, d->dofs, p->kb, p->kg, p->zv, p->pn, p->adb, p->am,
               p->pattern_mode == MAF_PATTERN_SYM, nthreads);
  build_symbolic(M.sym, M.numel, M.numnp, M.ndf, M.nmdf, M.IX0.data(), M.ID0.data(), M.cfg.rowmask);
  build_tdb(M.nuel1, M.nuel2, M.line1.data(), M.line2.data(), d->xi, M.tdb);
  // column pointers by node, so that the element gather needs no ID -> colptr indirection
  M.nodecol.assign((size_t)8 * M.numnp, -1);
  M.nodemask32.assign(M.sym.nodemask.begin(), M.sym.nodemask.end());
  for (int64_t n = 0; n < M.numnp; ++n) {
    for (int J = 0; J < M.ndf; ++J) {
      const int32_t eq = M.ID0[(size_t)M.ndf * n + J];
      if (eq >= 0) M.nodecol[8 * n + J] = M.sym.colptr[eq];
    }
  }
  // basis blocks per unique element (skipped when the knot vectors are so irregular that the table would be large)
  if ((int64_t)M.nuel1 * M.nuel2 <= 4096) {
    M.utab.assign((size_t)M.nuel1 * M.nuel2 * BASIS_DOUBLES, 0.0);
    for (int u2 = 0; u2 < M.nuel2; ++u2)
      for (int u1 = 0; u1 < M.nuel1; ++u1)
        build_basis_block(0, 1, M.line1.data() + 30 * u1, M.line2.data() + 30 * u2,
                          M.tdb.data() + (size_t)81 * (u1 + (size_t)M.nuel1 * u2),
                          M.utab.data() + (size_t)BASIS_DOUBLES * (u1 + (size_t)M.nuel1 * u2));
  }

  // Neumann conditions in the reference's order (FiniteElement.jl:151-154)
  M.n_neu = d->n_neu;
  M.b_offs.assign(1, 0);
  for (int k = 0; k < d->n_neu; ++k) {
    const int bd = d->neu_bdry[k], ty = d->neu_type[k];
    MAF_REQUIRE(bd >= 1 && bd <= 4, "unknown Boundary code");
    MAF_REQUIRE(ty >= 1 && ty <= 3, "unknown Neumann code");
    // FiniteElement.jl:377-383: MOMENT is implemented for F_BEND only, anything else asserts
    MAF_REQUIRE(ty != MAF_MOMENT || p->scenario == MAF_F_BEND, "Neumann boundary condition not implemented");
    MAF_REQUIRE(d->bdry_elems[bd - 1] || d->bdry_count[bd - 1] == 0, "null boundary element list");
    for (int64_t q = 0; q < d->bdry_count[bd - 1]; ++q) {
      const int64_t e = d->bdry_elems[bd - 1][q];
      MAF_REQUIRE(e >= 1 && e <= M.numel, "boundary element id outside 1..numel");
      M.b_elems.push_back((int32_t)(e - 1));
    }
    M.b_offs.push_back((int32_t)M.b_elems.size());
    M.b_bdry.push_back(bd);
    M.b_type.push_back(ty);
    M.b_val.push_back(d->neu_val[k]);
  }
}

inline Tables host_tables(const HostModel& M) {
  Tables T;
  T.elslot = M.elslot.empty() ? nullptr : M.elslot.data();
  T.elbase = M.elbase.empty() ? nullptr : M.elbase.data();
  T.elclass = M.elclass.empty() ? nullptr : M.elclass.data();
  T.IX = M.IX0.data(); T.ID = M.ID0.data(); T.nodemask = M.sym.nodemask.data();
  T.uel1 = M.uel1.data(); T.uel2 = M.uel2.data(); T.line1 = M.line1.data(); T.line2 = M.line2.data();
  T.tdb = M.tdb.data(); T.colptr = M.sym.colptr.data(); T.elpair = M.sym.elpair.data();
  T.pairoff = M.sym.pairoff.data(); T.eq0 = M.sym.eq0.data(); T.nodecol = M.nodecol.data(); T.nodemask32 = M.nodemask32.data(); T.utab = M.utab.empty() ? nullptr : M.utab.data(); T.numnp = M.numnp; T.numel = M.numel; T.num1el = M.num1el; T.nuel1 = M.nuel1;
  T.el0 = 0;
  return T;
}
// host copy of the per-element scatter maps (CPU emulation only; needs M.sym.elpair, i.e. before maf_create frees it)
inline void build_host_elslot(HostModel& M) {
  M.elslot.assign((size_t)MAF_SLOT_INTS * M.numel, -1);
  M.elbase.assign((size_t)M.numel, 0);
  M.elclass.resize((size_t)M.numel);
  for (int64_t e = 0; e < M.numel; ++e) M.elclass[(size_t)e] = (int32_t)e;
  Tables T = host_tables(M);
  int overflow = 0;
  for (int64_t e = 0; e < M.numel; ++e)
    build_elslot(T, e, M.elslot.data() + (size_t)MAF_SLOT_INTS * e, M.elbase.data() + e, &overflow);
  if (overflow) throw std::runtime_error("scatter map offset exceeds 32 bits");
}
inline BoundaryTables host_boundary_tables(const HostModel& M) {
  BoundaryTables B;
  B.edge1 = M.edge1.data(); B.edge2 = M.edge2.data(); B.elems = M.b_elems.data(); B.offs = M.b_offs.data();
  B.bdry = M.b_bdry.data(); B.ntype = M.b_type.data(); B.nval = M.b_val.data(); B.n_neu = M.n_neu;
  return B;
}

// value multiplying the traction of a condition: nval, or Mval = nval*min(time/bend_tm, 1) (FiniteElement.jl:379)
inline double neumann_value(int ntype, double nval, double time, double bend_tm) {
  return ntype == MAF_MOMENT ? nval * std::min(time / bend_tm, 1.0) : nval;
}

// host tables of the deterministic path
struct GatherHost {
  std::vector<int32_t> pair_node;
  std::vector<int16_t> ij_of;
  uint8_t class_I[64], class_J[64];
  int nij = 0;    // stride of a staging row in doubles: the class count rounded up to even (16-byte loads in the gather)
  int ncls = 0;   // number of (row dof, col dof) classes
  int sym_fill = 0;
  // pair-contribution classes (GatherTables::pclass ...); pclass empty = not built (more than 65535 classes)
  std::vector<uint16_t> pclass;
  std::vector<int32_t> eref, cde;
  std::vector<uint8_t> ccnt, crow;
};
// For every node pair (A, B) the staged rows it sums: (element - eref[B], 9 a + b) for the elements that contain both,
// ascending. Equal lists share a class.
inline void build_pair_classes(const HostModel& M, GatherHost& GH) {
  const Symbolic& S = M.sym;
  struct Sig { uint8_t n; int32_t de[9]; uint8_t row[9]; };
  auto sig_of = [&](int64_t p, int64_t B, Sig& g) -> bool {
    const int32_t A = S.nbr[p];
    const int64_t qa = S.n2e_ptr[B], qb = S.n2e_ptr[B + 1];
    const int64_t er = qb > qa ? S.n2e[qa] : 0;
    g.n = 0;
    for (int k = 0; k < 9; ++k) { g.de[k] = 0; g.row[k] = 0; }
    for (int64_t q = qa; q < qb; ++q) {
      const int64_t e = S.n2e[q];
      int a = -1;
      for (int k = 0; k < 9; ++k)
        if (M.IX0[9 * e + k] == A) a = k;
      if (a < 0) continue;
      if (g.n >= 9) return false;   // (not a 9-node tensor-product mesh: the caller gives the classes up)
      g.de[g.n] = (int32_t)(e - er);
      g.row[g.n] = (uint8_t)(9 * a + S.n2e_loc[q]);
      g.n += 1;
    }
    return true;
  };
  auto hash_of = [](const Sig& g) {
    uint64_t h = 1469598103934665603ull ^ g.n;
    for (int k = 0; k < 9; ++k) {
      h = (h ^ (uint32_t)g.de[k]) * 1099511628211ull;
      h = (h ^ g.row[k]) * 1099511628211ull;
    }
    return h;
  };
  auto same = [](const Sig& x, const Sig& y) {
    if (x.n != y.n) return false;
    for (int k = 0; k < 9; ++k)
      if (x.de[k] != y.de[k] || x.row[k] != y.row[k]) return false;
    return true;
  };
  GH.eref.assign(M.numnp, 0);
  for (int64_t B = 0; B < M.numnp; ++B)
    if (S.n2e_ptr[B + 1] > S.n2e_ptr[B]) GH.eref[B] = S.n2e[S.n2e_ptr[B]];
  std::vector<uint64_t> hp(S.npairs);
  std::mutex mu;
  std::unordered_map<uint64_t, Sig> reps;   // one representative list per hash
  std::atomic<bool> collision{false};
  parallel_for(M.numnp, [&](int64_t lo, int64_t hi) {
    std::unordered_map<uint64_t, Sig> local;
    Sig g;
    for (int64_t B = lo; B < hi; ++B)
      for (int64_t p = S.nbr_ptr[B]; p < S.nbr_ptr[B + 1]; ++p) {
        if (!sig_of(p, B, g)) collision = true;
        const uint64_t h = hash_of(g);
        hp[p] = h;
        auto it = local.find(h);
        if (it == local.end()) local.emplace(h, g);
        else if (!same(it->second, g)) collision = true;
      }
    std::lock_guard<std::mutex> lk(mu);
    for (auto& kv : local) {
      auto it = reps.find(kv.first);
      if (it == reps.end()) reps.emplace(kv.first, kv.second);
      else if (!same(it->second, kv.second)) collision = true;
    }
  });
  GH.pclass.clear();
  GH.ccnt.clear(); GH.cde.clear(); GH.crow.clear();
  if (collision || reps.size() > 65535) return;   // the gather kernel falls back to scanning the element lists
  std::unordered_map<uint64_t, uint16_t> id;
  for (auto& kv : reps) {
    id.emplace(kv.first, (uint16_t)GH.ccnt.size());
    GH.ccnt.push_back(kv.second.n);
    for (int k = 0; k < 9; ++k) { GH.cde.push_back(kv.second.de[k]); GH.crow.push_back(kv.second.row[k]); }
  }
  GH.pclass.resize(S.npairs);
  parallel_for(S.npairs, [&](int64_t lo, int64_t hi) {
    for (int64_t p = lo; p < hi; ++p) GH.pclass[p] = id.find(hp[p])->second;
  });
}
inline void build_gather_host(const HostModel& M, GatherHost& GH) {
  const Symbolic& S = M.sym;
  GH.pair_node.resize(S.npairs);
  for (int64_t B = 0; B < M.numnp; ++B)
    for (int64_t p = S.nbr_ptr[B]; p < S.nbr_ptr[B + 1]; ++p) GH.pair_node[p] = (int32_t)B;
  // (I,J) classes = distinct (row dof, col dof) pairs that own tangent tasks, in destination order (J, then I):
  // numbered by build_config (Config::ij_of)
  GH.ij_of.assign(64, -1);
  GH.nij = 0;
  for (int J = 0; J < 8; ++J)
    for (int I = 0; I < 8; ++I)
      if (M.cfg.ij_of[8 * I + J] >= 0) {
        if (M.cfg.ij_of[8 * I + J] != GH.nij) throw std::runtime_error("internal: class numbering");
        GH.class_I[GH.nij] = (uint8_t)I;
        GH.class_J[GH.nij] = (uint8_t)J;
        GH.ij_of[8 * I + J] = (int16_t)GH.nij++;
      }
  GH.ncls = GH.nij;
  GH.nij += GH.nij & 1;
  GH.sym_fill = M.pattern_mode == MAF_PATTERN_SYM ? 1 : 0;
}
inline void fill_gather_tables(const GatherHost& GH, GatherTables& G) {
  for (int c = 0; c < 64; ++c) { G.class_I[c] = GH.class_I[c]; G.class_J[c] = GH.class_J[c]; }
  G.sym_fill = GH.sym_fill;
  G.ncls = GH.ncls;
  G.ring = 0;
  G.pclass = nullptr; G.eref = nullptr; G.ccnt = nullptr; G.cde = nullptr; G.crow = nullptr;
}

// ---- strips of element rows over several GPUs (the reference's per-task chunks, FiniteElement.jl:88-89) ----
// What the elements [e0, e1) touch: nodes [node_lo, node_hi), equations [eq_lo, eq_hi), nnz slots [slot_lo, slot_hi) --
// each contiguous because unknowns are numbered node-major (Mesh.jl:276-284).
struct TouchedRange {
  int64_t e0 = 0, e1 = 0, node_lo = 0, node_hi = 0, eq_lo = 0, eq_hi = 0, slot_lo = 0, slot_hi = 0;
};
inline TouchedRange touched_range(const HostModel& M, int64_t e0, int64_t e1) {
  TouchedRange R;
  R.e0 = e0; R.e1 = e1;
  int64_t lo = M.numnp, hi = -1;
  for (int64_t k = 9 * e0; k < 9 * e1; ++k) {
    lo = std::min<int64_t>(lo, M.IX0[k]);
    hi = std::max<int64_t>(hi, M.IX0[k]);
  }
  if (hi < 0) { lo = 0; hi = -1; }
  R.node_lo = lo; R.node_hi = hi + 1;
  int64_t eq_lo = M.nmdf, eq_hi = 0;
  for (int64_t k = lo * M.ndf; k < (hi + 1) * M.ndf; ++k)
    if (M.ID0[k] >= 0) { eq_lo = std::min<int64_t>(eq_lo, M.ID0[k]); eq_hi = std::max<int64_t>(eq_hi, M.ID0[k] + 1); }
  if (eq_hi <= eq_lo) eq_lo = eq_hi = 0;
  R.eq_lo = eq_lo; R.eq_hi = eq_hi;
  R.slot_lo = M.sym.colptr[eq_lo]; R.slot_hi = M.sym.colptr[eq_hi];
  return R;
}
// Strip `rank` of `nranks`: element rows [rank num2el / nranks, (rank + 1) num2el / nranks). Every strip needs at least
// two element rows: a quadratic strip touches node rows e2 .. e2 + 2, so with thinner strips rank k and rank k + 2
// would share a node row and the neighbour-only exchange would drop that overlap.
inline TouchedRange strip_range(const HostModel& M, int rank, int nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("strip rank outside 0..nranks-1");
  if (nranks > 1 && M.num2el / nranks < 2)
    throw std::runtime_error("the mesh has too few element rows for this many strips (each needs at least 2 rows)");
  const int64_t r0 = ((int64_t)rank * M.num2el) / nranks, r1 = ((int64_t)(rank + 1) * M.num2el) / nranks;
  return touched_range(M, r0 * M.num1el, r1 * M.num1el);
}

// processing order of the elements of a range: Z-order (Morton) over (e1, e2), so that the elements a wave of CTAs
// works on at the same time form a compact 2-D patch and the contributions to an nnz slot / the reads of a node
// arrive while the line is still in L2 (a row-major sweep re-fetches every nzval sector once per element row)
inline void build_element_order(int num1el, int64_t e0, int64_t e1, std::vector<int32_t>& order) {
  auto spread = [](uint64_t v) {
    v &= 0xffffffffull;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
  };
  std::vector<std::pair<uint64_t, int32_t>> key((size_t)(e1 - e0));
  for (int64_t e = e0; e < e1; ++e)
    key[(size_t)(e - e0)] = {spread((uint64_t)(e % num1el)) | (spread((uint64_t)(e / num1el)) << 1), (int32_t)e};
  std::sort(key.begin(), key.end());
  order.resize(key.size());
  for (size_t k = 0; k < key.size(); ++k) order[k] = key[k].second;
}

}  // namespace maf
