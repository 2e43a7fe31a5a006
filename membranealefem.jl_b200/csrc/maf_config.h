// maf_config.h -- host-side construction of the element-kernel configuration (block structure of the
// Gauss-point tangent, work lists, shared-memory layout) from the motion, the dof map and the parameters.
// Pure C++ (no CUDA) so that the CPU emulation harness in tests/ builds the very same tables.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "maf_element.cuh"

namespace maf {

inline int kind_of(int nr, int nc) {
  static const int tab[8][2] = {{1, 1}, {1, 2}, {1, 3}, {2, 1}, {2, 2}, {3, 5}, {5, 5}, {6, 5}};
  for (int k = 0; k < 8; ++k)
    if (tab[k][0] == nr && tab[k][1] == nc) return k;
  throw std::runtime_error("no block_accumulate instantiation for this block shape");
}
// Cost of one task (9 outputs) for the warp schedule: FP64 operations per Gauss point plus the fixed part
// (set-up, loop latency, scatter), in the same unit; MAF_TASK_FIXED was fitted to the measured cycles per chunk
// (profiling build, tools/profile_once.py): the short blocks cost far more than their arithmetic.
#ifndef MAF_TASK_FIXED
#define MAF_TASK_FIXED 37
#endif
inline int kind_cost(int kind, bool fused, bool tr) {
  static const int tab[8][2] = {{1, 1}, {1, 2}, {1, 3}, {2, 1}, {2, 2}, {3, 5}, {5, 5}, {6, 5}};
  if (fused) return MAF_TASK_FIXED + 107;
  if (tr) return MAF_TASK_FIXED + tab[kind][0] * tab[kind][1] + 9 * tab[kind][0];
  if (tab[kind][1] == 5) return MAF_TASK_FIXED + 5 * tab[kind][0] + 8 + 42;   // sum-factorised mesh-column blocks
  return MAF_TASK_FIXED + tab[kind][0] * tab[kind][1] + 9 * tab[kind][1];
}

// chunk plans found by tools/tune_plan.py on B200 (1001 x 1001 patch), keyed by motion and chunk count; NULL: heuristic
inline const char* tuned_plan(int motion, int nchunks, int nwarps) {
  if (nwarps != 4) return nullptr;
#ifdef MAF_NO_TUNED_PLAN
  return nullptr;
#endif
  // gains over the heuristic plan at 1001 x 1001 (gpurun_out/tune3_*.log of round 1; the plans are retuned whenever
  // the kernel changes -- the same plan lost 3 % when only the scatter map changed)
  if (motion == M_ALEVB && nchunks == 14) return "1,12,10/0,9/13,11,4,5,6/2,7,8,3";                // +4.9 % (r2 retune: +0.6 %)
  if (motion == M_ALEV && nchunks == 17) return "6,14,13,15,1/8,5,16/3,2,10,11/0,4,9,7,12";        // +6.5 % (+20 % over LPT)
  if (motion == M_EUL && nchunks == 16) return "6,7,5,10/13,3,8/9,0,2,15/14,12,4,11,1";            // +5.8 % (r2 retune: +0.9 %)
  if (motion == M_LAG && nchunks == 6) return "1/2/0,4/3,5";                                       // +0.6 %
  return nullptr;
}

// dofs8: column (1-based) of vx vy vz vmx vmy vmz lambda pm, or 0 (Mesh.dofs, Bc.jl:414-431)
inline void build_config(Config& cfg, int motion, int ndf, const int32_t dofs8[8], double kb, double kg, double zv,
                         double pn, double adb, double am, int pattern_sym, int nthreads) {
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.motion = motion;
  cfg.ndf = ndf;
  cfg.nthreads = nthreads;
  cfg.mat = Material{kb, kg, zv, pn, adb, am, adb / zv};
  cfg.dbscale = adb / zv;
  cfg.ncomp[F_V] = 3; cfg.ncomp[F_M] = 3; cfg.ncomp[F_L] = 1; cfg.ncomp[F_P] = 1;
  for (int f = 0; f < NFIELD; ++f)
    for (int i = 0; i < 3; ++i) cfg.fdof[f][i] = -1;
  for (int i = 0; i < 3; ++i) { cfg.fdof[F_V][i] = dofs8[i] - 1; cfg.fdof[F_M][i] = dofs8[3 + i] - 1; }
  cfg.fdof[F_L][0] = dofs8[6] - 1;
  cfg.fdof[F_P][0] = dofs8[7] - 1;
  if (cfg.fdof[F_L][0] < 0) throw std::runtime_error("the surface tension must be a degree of freedom");
  const bool has_m = dofs8[3] || dofs8[4] || dofs8[5];
  const bool has_p = dofs8[7] != 0;
  // get_m_motion_order (Mesh.jl:529-542)
  cfg.mesh_field = motion == M_STATIC ? -1 : (motion == M_LAG ? F_V : F_M);
  if (motion == M_LAG && has_m) throw std::runtime_error("LAG motion carries no mesh-velocity dofs");
  if ((motion == M_EUL || motion == M_ALEV || motion == M_ALEVB) && !has_m)
    throw std::runtime_error("EUL/ALE motion needs mesh-velocity dofs");
  if ((motion == M_ALEV || motion == M_ALEVB) != has_p)
    throw std::runtime_error("the mesh pressure is a dof exactly for ALEV/ALEVB");

  const int vr0 = pn != 0.0 ? 0 : 1, vnr = pn != 0.0 ? 6 : 5;
  // ALEVB: the mesh equations are the membrane equations with the viscous stress of the mesh velocity, no surface
  // tension and the mesh pressure as normal load (GeoDynStress.jl:156-159, FiniteElement.jl:304-313): the bending and
  // moment tangent of the v rows and of the vm rows w.r.t. the mesh dofs coincide. Without a normal pressure on the
  // v rows both blocks are produced by one fused task from A_m and the 2 x 2 corner difference d(S_v - S_m).
  cfg.fused_vm = (motion == M_ALEVB && pn == 0.0) ? 1 : 0;
  struct B { int f, g, c0, nr, d0, nc, db, notask, fused; };
  std::vector<B> bl;
  switch (motion) {
    case M_STATIC:
      bl = {{F_V, F_V, 1, 2, 1, 2, 0}, {F_V, F_L, 1, 2, 0, 1, 0}, {F_L, F_V, 0, 1, 1, 2, 0}, {F_L, F_L, 0, 1, 0, 1, 1}};
      break;
    case M_LAG:
      bl = {{F_V, F_V, vr0, vnr, 1, 5, 0}, {F_V, F_L, 1, 2, 0, 1, 0}, {F_L, F_V, 0, 1, 1, 2, 0}, {F_L, F_L, 0, 1, 0, 1, 1}};
      break;
    case M_EUL:
      bl = {{F_V, F_V, 1, 2, 1, 2, 0}, {F_V, F_M, vr0, vnr, 1, 5, 0}, {F_V, F_L, 1, 2, 0, 1, 0},
            {F_M, F_V, 0, 1, 0, 1, 0}, {F_M, F_M, 0, 1, 0, 3, 0},
            {F_L, F_V, 0, 1, 1, 2, 0}, {F_L, F_M, 0, 1, 1, 2, 0}, {F_L, F_L, 0, 1, 0, 1, 1}};
      break;
    case M_ALEV:
    case M_ALEVB:
      bl = {{F_V, F_V, 1, 2, 1, 2, 0}, {F_V, F_M, vr0, vnr, 1, 5, 0}, {F_V, F_L, 1, 2, 0, 1, 0},
            {F_M, F_M, 0, motion == M_ALEVB ? 6 : 3, 1, 5, 0}, {F_M, F_P, 0, 1, 0, 1, 0},
            {F_L, F_V, 0, 1, 1, 2, 0}, {F_L, F_M, 0, 1, 1, 2, 0}, {F_L, F_L, 0, 1, 0, 1, 1},
            {F_P, F_V, 0, 1, 0, 1, 0}, {F_P, F_M, 0, 1, 0, 3, 0}, {F_P, F_P, 0, 1, 0, 1, 1}};
      break;
    default: throw std::runtime_error("unknown motion code");
  }
  if (cfg.fused_vm) {
    bl[1] = B{F_V, F_M, 1, 2, 1, 2, 0, 1, 0};   // corner difference only, consumed by the fused block
    bl[3].fused = 1;
  }
  // row channel range per field = union over its blocks; column ranges are per block
  int rlo[NFIELD], rhi[NFIELD];
  for (int f = 0; f < NFIELD; ++f) { rlo[f] = 99; rhi[f] = -1; }
  for (int f = 0; f < NFIELD; ++f)
    for (int g = 0; g < NFIELD; ++g) cfg.coloff[f][g] = -1;
  auto is_mesh = [&](const B& b) { return b.nc == 5 && b.d0 == 1 && b.g == cfg.mesh_field; };
  for (const B& b : bl) {
    rlo[b.f] = std::min(rlo[b.f], b.c0);
    rhi[b.f] = std::max(rhi[b.f], b.c0 + b.nr);
    cfg.cd0[b.f][b.g] = b.d0;
    cfg.cnc[b.f][b.g] = is_mesh(b) ? 2 : b.nc;   // mesh blocks store N1,N2 per dof j and 3 shared b-direction columns
    if (b.nc == 5 && !is_mesh(b)) throw std::runtime_error("a 5-channel block must be a mesh-column block");
  }
  int off = 0;
  for (int f = 0; f < NFIELD; ++f) {
    if (rhi[f] < 0) { cfg.rc0[f] = 0; cfg.rnc[f] = 0; cfg.aoff[f] = off; cfg.ald[f] = 0; cfg.bcol[f] = -1; continue; }
    cfg.rc0[f] = rlo[f];
    cfg.rnc[f] = rhi[f] - rlo[f];
    int ld = 0;
    cfg.bcol[f] = -1;
    for (int g = 0; g < NFIELD; ++g)
      if (cfg.cnc[f][g] > 0) { cfg.coloff[f][g] = ld; ld += cfg.ncomp[g] * cfg.cnc[f][g]; }
    for (const B& b : bl)
      if (b.f == f && is_mesh(b)) { ld += ld & 1; cfg.bcol[f] = ld; ld += 3; }   // 16-byte aligned for LDS.128
    ld += ld & 1;
    cfg.ald[f] = ld;
    cfg.aoff[f] = off;
    off += cfg.ncomp[f] * cfg.rnc[f] * ld;
  }
  // per-Gauss-point stride of A: even (16-byte loads) and = 2 (mod 4), so that the stores of the Gauss-phase lanes,
  // which write the same offset of different Gauss points, spread over 8 bank groups (328 doubles put them on 2,
  // EUL's 320 on one: 9-way conflicts on every store of the phase)
  cfg.asize = off + (off & 1);
  while (cfg.asize % 4 != 2) cfg.asize += 2;

  cfg.nblocks = (int)bl.size();
  for (int k = 0; k < cfg.nblocks; ++k) {
    const B& b = bl[k];
    const bool q = is_mesh(b) && (b.f == F_V || (b.f == F_M && motion == M_ALEVB));
    cfg.blocks[k] = Block{(int8_t)b.f, (int8_t)b.g, (int8_t)b.c0, (int8_t)b.nr, (int8_t)b.d0, (int8_t)b.nc,
                          (int8_t)kind_of(b.nr, b.nc), (int8_t)b.db, (int8_t)(is_mesh(b) ? 1 : 0), (int8_t)(q ? 1 : 0),
                          (int8_t)b.notask, (int8_t)b.fused, (int8_t)((!is_mesh(b) && b.nr < b.nc) ? 1 : 0)};
  }
  // present components of every field; (row dof, col dof) classes of the deterministic staging rows,
  // in destination order (J, then I)
  for (int f = 0; f < NFIELD; ++f) {
    cfg.npc[f] = 0;
    for (int i = 0; i < cfg.ncomp[f]; ++i)
      if (cfg.fdof[f][i] >= 0) cfg.pcomp[f][cfg.npc[f]++] = (int8_t)i;
  }
  {
    bool has[64] = {false};
    for (int k = 0; k < cfg.nblocks; ++k)
      for (int ii = 0; ii < cfg.npc[bl[k].f]; ++ii)
        for (int jj = 0; jj < cfg.npc[bl[k].g]; ++jj)
          has[8 * cfg.fdof[bl[k].f][cfg.pcomp[bl[k].f][ii]] + cfg.fdof[bl[k].g][cfg.pcomp[bl[k].g][jj]]] = true;
    int nij = 0;
    for (int J = 0; J < 8; ++J)
      for (int I = 0; I < 8; ++I) cfg.ij_of[8 * I + J] = has[8 * I + J] ? (int8_t)nij++ : (int8_t)-1;
  }
  // tangent chunks: <= 32 consecutive tasks of one block
  struct Ch { int blk, first, count, cost; };
  std::vector<Ch> chs;
  cfg.ntasks = 0;
  for (int k = 0; k < cfg.nblocks; ++k) {
    const Block& b = cfg.blocks[k];
    if (b.notask) continue;
    const int nt = 9 * cfg.npc[b.f] * cfg.npc[b.g];
    if (nt == 0) continue;
    if (nt > 255) throw std::runtime_error("task index overflow");
    cfg.ntasks += nt;
    const int parts = (nt + 31) / 32, per = (nt + parts - 1) / parts;
    for (int first = 0; first < nt; first += per)
      chs.push_back(Ch{k, first, std::min(per, nt - first), kind_cost(b.kind, b.fused != 0, b.tr != 0)});
  }
  if ((int)chs.size() > MAF_MAX_CHUNKS) throw std::runtime_error("chunk table overflow");

  // Gauss-point work items: GEO_A (6 per gp) when the mesh moves, GEO_B (3 per gp: one per b-direction), LIN (1 per gp)
  std::vector<Item> items;
  if (cfg.mesh_field >= 0) {
    for (int gp = 0; gp < 9; ++gp)
      for (int gam = 0; gam < 2; ++gam)
        for (int j = 0; j < 3; ++j) {
          if (cfg.fdof[cfg.mesh_field][j] < 0) continue;
          items.push_back(Item{IT_GEO_A, (uint8_t)gp, (uint8_t)gam, (uint8_t)j});
        }
    for (int gp = 0; gp < 9; ++gp)
      for (int k = 0; k < 3; ++k) items.push_back(Item{IT_GEO_B, (uint8_t)gp, (uint8_t)k, 0});
  }
  for (int gp = 0; gp < 9; ++gp) items.push_back(Item{IT_LIN, (uint8_t)gp, 0, 0});
  // closed-form columns of the components of v / vm as their own items (maf_element.cuh::lin_split)
  if (lin_split(motion))
    for (int gp = 0; gp < 9; ++gp)
      for (int j = 0; j < 3; ++j) items.push_back(Item{IT_LIN_C, (uint8_t)gp, 0, (uint8_t)j});
  if ((int)items.size() > MAF_MAX_ITEMS) throw std::runtime_error("item table overflow");
  cfg.nitems = (int)items.size();
  for (int k = 0; k < cfg.nitems; ++k) cfg.items[k] = items[k];

  // thread -> work maps. Work of one class (item type / block kind) shares a code path, so every class is cut
  // into warp-sized chunks and each chunk is given to the least-loaded warp (longest-processing-time first):
  // the lanes of a warp never diverge on the class, and the warps of the CTA finish at about the same time.
  const int nwarps = nthreads / 32;
  // `pin_cls` (if >= 0) is placed on the last warp and nothing else is (used for the Gauss phase, where the last
  // warp first gathers the next element and then only runs the cheap LIN items)
  auto schedule = [&](const std::vector<int>& cls, const std::vector<int>& cost, const std::vector<int>& init_load,
                      int pin_cls, int16_t* slot, int& rounds) {
    struct Chunk { int cost, cls; std::vector<int> ids; };
    std::vector<Chunk> chunks;
    for (size_t k = 0; k < cls.size();) {
      size_t e = k;
      while (e < cls.size() && cls[e] == cls[k]) ++e;
      const int n = (int)(e - k), parts = (n + 31) / 32, per = (n + parts - 1) / parts;
      for (int q = 0; q < parts; ++q) {
        Chunk ch;
        ch.cost = cost[k];
        ch.cls = cls[k];
        for (size_t t = k + (size_t)q * per; t < std::min(e, k + (size_t)(q + 1) * per); ++t) ch.ids.push_back((int)t);
        chunks.push_back(ch);
      }
      k = e;
    }
    std::stable_sort(chunks.begin(), chunks.end(), [](const Chunk& x, const Chunk& y) { return x.cost > y.cost; });
    std::vector<int> load(init_load.begin(), init_load.end());
    load.resize(nwarps, 0);
    std::vector<std::vector<const Chunk*>> plan(nwarps);
    for (const Chunk& ch : chunks) {
      int wbest = 0;
      const int wlim = pin_cls >= 0 ? nwarps - 1 : nwarps;
      for (int w = 1; w < wlim; ++w)
        if (load[w] < load[wbest]) wbest = w;
      if (pin_cls >= 0 && ch.cls == pin_cls) wbest = nwarps - 1;
      plan[wbest].push_back(&ch);
      load[wbest] += ch.cost;
    }
    rounds = 0;
    for (int w = 0; w < nwarps; ++w) rounds = std::max(rounds, (int)plan[w].size());
    if (rounds * nthreads > MAF_MAX_SLOTS) throw std::runtime_error("slot table overflow");
    for (int s2 = 0; s2 < rounds * nthreads; ++s2) slot[s2] = -1;
    for (int w = 0; w < nwarps; ++w)
      for (size_t r = 0; r < plan[w].size(); ++r)
        for (size_t l = 0; l < plan[w][r]->ids.size(); ++l)
          slot[r * nthreads + w * 32 + l] = (int16_t)plan[w][r]->ids[l];
  };
  {
    std::vector<int> cls(cfg.nitems), cost(cfg.nitems);
    for (int k = 0; k < cfg.nitems; ++k) {
      cls[k] = cfg.items[k].type;
      const int ty = cfg.items[k].type;
      if (lin_split(motion)) cost[k] = ty == IT_GEO_A ? 10 : (ty == IT_GEO_B ? 5 : (ty == IT_LIN ? 4 : 2));
      else cost[k] = ty == IT_LIN ? 6 : 10;
    }
    schedule(cls, cost, {}, -1, cfg.item_slot, cfg.item_rounds);
  }
  {
    // chunks of tangent tasks: longest-processing-time first over the warps. During the tangent phase the first
    // warps also form the element residual (72 rows over the first 72 threads): account for that work when
    // balancing.
    std::stable_sort(chs.begin(), chs.end(), [](const Ch& x, const Ch& y) { return x.cost > y.cost; });
    cfg.nchunks = (int)chs.size();
    for (int k = 0; k < cfg.nchunks; ++k)
      cfg.chunks[k] = Chunk{(uint8_t)chs[k].blk, (uint8_t)chs[k].first, (uint8_t)chs[k].count, 0};
    std::vector<int> load(nwarps, 0);
    for (int w = 0; w < nwarps && w < 3; ++w) load[w] = 30;
    std::vector<std::vector<int>> plan(nwarps);
    for (int k = 0; k < cfg.nchunks; ++k) {
      int wbest = 0;
      for (int w = 1; w < nwarps; ++w)
        if (load[w] < load[wbest]) wbest = w;
      plan[wbest].push_back(k);
      load[wbest] += chs[k].cost;
    }
    // A tuned plan replaces the heuristic one: which chunks run side by side on the SM (shared-memory pipe against
    // FP64 pipe) matters more than the balance of the warps, and no cost model captured that (tools/tune_plan.py
    // searches the plans on the GPU). "c,c,c/c,c/..." = chunk ids per warp in execution order; MAF_PLAN overrides.
    {
      const char* text = std::getenv("MAF_PLAN");
      if (!text || !*text) text = tuned_plan(motion, cfg.nchunks, nwarps);
      if (text && *text) {
        std::vector<std::vector<int>> tp(1);
        std::vector<int> seen(cfg.nchunks, 0);
        bool ok = true;
        int v = -1;
        for (const char* c = text;; ++c) {
          if (*c >= '0' && *c <= '9') { v = (v < 0 ? 0 : v * 10) + (*c - '0'); continue; }
          if (v >= 0) {
            if (v >= cfg.nchunks || seen[v]++) ok = false;
            else tp.back().push_back(v);
            v = -1;
          }
          if (*c == '/') tp.emplace_back();
          else if (*c != ',' && *c != 0) ok = false;
          if (*c == 0) break;
        }
        for (int k = 0; k < cfg.nchunks; ++k) ok = ok && seen[k] == 1;
        if (!ok || (int)tp.size() > nwarps) throw std::runtime_error(std::string("invalid chunk plan: ") + text);
        tp.resize(nwarps);
        plan = tp;
      }
    }
    cfg.task_rounds = 0;
    for (int w = 0; w < nwarps; ++w) cfg.task_rounds = std::max(cfg.task_rounds, (int)plan[w].size());
    if (cfg.task_rounds > MAF_MAX_ROUNDS || nwarps > 8) throw std::runtime_error("chunk slot table overflow");
    for (int q = 0; q < MAF_MAX_ROUNDS * 8; ++q) cfg.chunk_slot[q] = -1;
    for (int w = 0; w < nwarps; ++w)
      for (size_t r = 0; r < plan[w].size(); ++r) cfg.chunk_slot[r * nwarps + w] = (int8_t)plan[w][r];
  }

  // rows present in the pattern of a column of dof J
  for (int J = 0; J < 8; ++J) cfg.rowmask[J] = 0;
  for (int g = 0; g < NFIELD; ++g)
    for (int j = 0; j < cfg.ncomp[g]; ++j) {
      const int J = cfg.fdof[g][j];
      if (J < 0) continue;
      unsigned m = 0;
      for (int f = 0; f < NFIELD; ++f)
        for (int i = 0; i < cfg.ncomp[f]; ++i) {
          if (cfg.fdof[f][i] < 0) continue;
          if (pattern_sym || cfg.coloff[f][g] >= 0) m |= 1u << cfg.fdof[f][i];
        }
      cfg.rowmask[J] = (uint8_t)m;
    }

  // shared-memory layout (doubles). Front block (what the gather phase writes; double buffered so that the
  // next element is gathered while the current one is contracted): offsets relative to the front block.
  int o = 0;
  cfg.o_x = o; o += 27;
  cfg.o_cv = o; o += 27;
  cfg.o_cm = o; o += 27;
  cfg.o_cl = o; o += 9;
  cfg.o_cp = o; o += 9;
  o += o & 1;
  cfg.o_phi = o; o += PHI_DOUBLES;      // basis block: Phi | FG | w | tdb, contiguous (BASIS_DOUBLES)
  cfg.o_FG = o; o += 9 * FG_STRIDE;
  cfg.o_w = o; o += 10;
  cfg.o_tdb = o; o += 82;
  cfg.o_int = o; o += (I_PAIR + 1) / 2;   // node ids, equation numbers, active-dof masks
  o += o & 1;
  cfg.o_slot = o; o += MAF_SLOT_INTS / 2;  // scatter map of the element (int32, build_elslot)
  cfg.o_po = o; o += 2;                    // its base (int64)
  o += o & 1;
  cfg.front_doubles = o;
  // back block (offsets relative to sm + 2 * front_doubles + 2 * MAF_IDS_DOUBLES: two ids buffers in between)
  o = 0;
  cfg.o_E = o; o += 9 * E_STRIDE;
  cfg.o_S = o; o += 9 * S_STRIDE;
  o += o & 1;
  cfg.o_G = o; o += 9 * G_STRIDE;
  o += o & 1;
  cfg.o_A = o; o += 9 * cfg.asize;
  cfg.o_ctr = o; o += 2;
  cfg.smem_doubles = 2 * cfg.front_doubles + 2 * MAF_IDS_DOUBLES + o;
  // interpolation table: E[gp][q] = sum_a Phi^{ch(q)}_a(gp) * (nodal values at o_src(q) + a); layout of E in
  // maf_element.cuh (E_A, E_C, E_DV, E_V, E_DM, E_VM, E_LAM, E_PM)
  for (int q = 0; q < 35; ++q) {
    int src, ch;
    if (q < 6) { ch = CH_N1 + q / 3; src = cfg.o_x + 9 * (q % 3); }
    else if (q < 15) { ch = CH_N11 + (q - 6) / 3; src = cfg.o_x + 9 * ((q - 6) % 3); }
    else if (q < 21) { ch = CH_N1 + (q - 15) / 3; src = cfg.o_cv + 9 * ((q - 15) % 3); }
    else if (q < 24) { ch = CH_N; src = cfg.o_cv + 9 * (q - 21); }
    else if (q < 30) { ch = CH_N1 + (q - 24) / 3; src = cfg.o_cm + 9 * ((q - 24) % 3); }
    else if (q < 33) { ch = CH_N; src = cfg.o_cm + 9 * (q - 30); }
    else if (q == 33) { ch = CH_N; src = cfg.o_cl; }
    else { ch = CH_N; src = cfg.o_cp; }
    cfg.interp_src[q] = (int16_t)src;
    cfg.interp_fo[q] = (int8_t)ch_fo(ch);
    cfg.interp_go[q] = (int8_t)ch_go(ch);
  }

  // flattened task descriptors (need the final storage layout and rowmask)
  for (int k = 0; k < cfg.nblocks; ++k) {
    const Block& b = cfg.blocks[k];
    TaskDesc& d = cfg.td[k];
    std::memset(&d, 0, sizeof(d));
    const int f = b.f, g = b.g;
    d.a0 = (int16_t)a_index(cfg, f, 0, b.c0, g, 0, b.d0);
    d.si = (int16_t)(cfg.rnc[f] * cfg.ald[f]);
    d.sj = (int16_t)cfg.cnc[f][g];
    d.ald = (int16_t)cfg.ald[f];
    d.boff0 = (int16_t)(b.mesh ? cfg.bcol[f] - cfg.coloff[f][g] : 0);
    if (b.fused) {
      d.av0 = (int16_t)a_index(cfg, F_V, 0, CH_N1, g, 0, CH_N1);
      d.svi = (int16_t)(cfg.rnc[F_V] * cfg.ald[F_V]);
      d.svj = (int16_t)cfg.cnc[F_V][g];
      d.aldv = (int16_t)cfg.ald[F_V];
    }
    d.c0 = (uint8_t)b.c0; d.d0 = (uint8_t)b.d0; d.kind = (uint8_t)b.kind; d.npcf = (uint8_t)cfg.npc[f];
    d.mesh = (uint8_t)b.mesh; d.qterm = (uint8_t)b.qterm; d.fused = (uint8_t)b.fused; d.tr = (uint8_t)b.tr; d.db = (uint8_t)b.db;
    for (int q = 0; q < 3; ++q) {
      d.ic[q] = q < cfg.npc[f] ? (uint8_t)cfg.pcomp[f][q] : 0;
      d.jc[q] = q < cfg.npc[g] ? (uint8_t)cfg.pcomp[g][q] : 0;
      d.I[q] = (q < cfg.ncomp[f] && cfg.fdof[f][q] >= 0) ? (uint8_t)cfg.fdof[f][q] : 0;
      d.J[q] = (q < cfg.ncomp[g] && cfg.fdof[g][q] >= 0) ? (uint8_t)cfg.fdof[g][q] : 0;
      d.Iv[q] = cfg.fdof[F_V][q] >= 0 ? (uint8_t)cfg.fdof[F_V][q] : 0;
      d.rm[q] = (q < cfg.ncomp[g] && cfg.fdof[g][q] >= 0) ? cfg.rowmask[cfg.fdof[g][q]] : 0;
    }
  }
  // everything that is read with 16-byte loads must sit on an even double offset
  bool ok = !(cfg.o_A & 1) && !(cfg.o_phi & 1) && !(cfg.o_FG & 1) && !(cfg.asize & 1) &&
            !(cfg.front_doubles & 1) && !(cfg.o_po & 1) && !(cfg.o_slot & 1);
  for (int f = 0; f < NFIELD; ++f) {
    ok = ok && !(cfg.aoff[f] & 1) && !(cfg.ald[f] & 1) && (cfg.bcol[f] < 0 || !(cfg.bcol[f] & 1));
    for (int g = 0; g < NFIELD; ++g)
      if (cfg.bcol[f] >= 0 && g == cfg.mesh_field && cfg.coloff[f][g] >= 0) ok = ok && !(cfg.coloff[f][g] & 1);
  }
  if (!ok) throw std::runtime_error("internal: misaligned shared-memory layout");
}

// Dohrmann-Bochev matrices tmpDB = G^T H^-1 G of every unique element (FiniteElement.jl:279-281, 315-323):
// G = sum_gp NDB N^T w, H = sum_gp NDB NDB^T w with NDB = (xi[gp1], xi[gp2], 1). State independent.
// line tables: [uel][gp][10] = (w, N[3], dN[3], ddN[3]). out: (nuel1*nuel2) x 81, [a][b].
inline void build_tdb(int nuel1, int nuel2, const double* line1, const double* line2, const double xi[3],
                      std::vector<double>& out) {
  out.assign((size_t)nuel1 * nuel2 * 81, 0.0);
  for (int u2 = 0; u2 < nuel2; ++u2)
    for (int u1 = 0; u1 < nuel1; ++u1) {
      double G[3][9] = {{0}}, H[3][3] = {{0}};
      for (int gp = 0; gp < 9; ++gp) {
        const double* f1 = line1 + 30 * u1 + 10 * (gp % 3);
        const double* f2 = line2 + 30 * u2 + 10 * (gp / 3);
        const double w = f1[0] * f2[0];
        const double ndb[3] = {xi[gp % 3], xi[gp / 3], 1.0};
        for (int i = 0; i < 3; ++i) {
          for (int a = 0; a < 9; ++a) G[i][a] += ndb[i] * (f1[1 + a % 3] * f2[1 + a / 3]) * w;
          for (int j = 0; j < 3; ++j) H[i][j] += ndb[i] * ndb[j] * w;
        }
      }
      const double a = H[0][0], b = H[0][1], c = H[0][2], d = H[1][0], e = H[1][1], f = H[1][2], g = H[2][0],
                   h = H[2][1], i = H[2][2];
      const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
      const double id = 1.0 / det;
      const double Hi[3][3] = {{(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id},
                               {(f * g - d * i) * id, (a * i - c * g) * id, (c * d - a * f) * id},
                               {(d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id}};
      double* T = out.data() + (size_t)81 * (u1 + (size_t)nuel1 * u2);
      for (int p = 0; p < 9; ++p) {
        double t1[3];
        for (int j = 0; j < 3; ++j) t1[j] = G[0][p] * Hi[0][j] + G[1][p] * Hi[1][j] + G[2][p] * Hi[2][j];
        for (int q = 0; q < 9; ++q) T[9 * p + q] = t1[0] * G[0][q] + t1[1] * G[1][q] + t1[2] * G[2][q];
      }
    }
}

}  // namespace maf
