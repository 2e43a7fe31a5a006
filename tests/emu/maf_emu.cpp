// tests/emu/maf_emu.cpp -- TEST INFRASTRUCTURE. Runs the library's element / boundary phase functions
// (membranealefem.jl_b200/csrc/maf_element.cuh, maf_boundary.cuh -- the exact code the CUDA kernels execute)
// on the CPU by looping the "threads" of a CTA between barriers. Lets the no-GPU container validate the kernel
// logic, the symbolic phase and the scatter maps against the oracle. Never loaded by the product.
#include <cstring>
#include <string>
#include <vector>

#include "../../membranealefem.jl_b200/csrc/maf_host.h"
#include "../../membranealefem.jl_b200/csrc/maf_state.cuh"

using namespace maf;

static std::string g_err;

template <int MOTION>
static void run_area(const HostModel& M, const Tables& T, const double* xms, const double* cps, double dt, double* r,
                     double* nzval, double* kel, double* rel, const GatherHost* GH, int64_t e0, int64_t e1) {
  const Config& cfg = M.cfg;
  const int nt = cfg.nthreads;
  if (nt != MAF_NT) throw std::runtime_error("the element phases are written for MAF_NT threads");
  std::vector<double> smem_all(cfg.smem_doubles);
  std::vector<int32_t> order;
  build_element_order(M.num1el, e0, e1, order);
  const int nfront = cfg.front_doubles;
  int32_t* ids0 = reinterpret_cast<int32_t*>(smem_all.data() + 2 * nfront);
  double* sm = smem_all.data() + 2 * nfront + 2 * MAF_IDS_DOUBLES;
  // same software pipeline as the kernel (the asynchronous copies complete at once here)
  std::fill(smem_all.begin(), smem_all.end(), std::nan(""));  // any read of an unwritten slot poisons the result
  for (int t = 0; t < nt; ++t) gather_init(t, cfg, smem_all.data());
  for (int t = 0; t < nt; ++t) gather_init(t, cfg, smem_all.data() + nfront);
  if (!order.empty()) {
    for (int t = 0; t < nt; ++t) gather_ids_async(t, T, order[0], ids0);
    for (int t = 0; t < nt; ++t) gather_data_async(t, cfg, T, ids0, xms, cps, smem_all.data());
    if (order.size() > 1)
      for (int t = 0; t < nt; ++t) gather_ids_async(t, T, order[1], ids0 + MAF_IDS_INTS);
  }
  int cur = 0;
  for (size_t k = 0; k < order.size(); ++k, cur ^= 1) {
    const int64_t el = order[k];
    const double* fr = smem_all.data() + cur * nfront;
    std::fill(sm, smem_all.data() + smem_all.size(), std::nan(""));
    if (k + 1 < order.size()) {
      for (int t = 0; t < nt; ++t)
        gather_data_async(t, cfg, T, ids0 + (cur ^ 1) * MAF_IDS_INTS, xms, cps, smem_all.data() + (cur ^ 1) * nfront);
      if (k + 2 < order.size())
        for (int t = 0; t < nt; ++t) gather_ids_async(t, T, order[k + 2], ids0 + cur * MAF_IDS_INTS);
    }
    for (int t = 0; t < nt; ++t) phase_interp(t, nt, cfg, fr, sm);
    for (int t = 0; t < nt; ++t) phase_gauss<MOTION>(t, cfg, dt, fr, sm);
    KSink sink{nzval, nullptr, 0};
    if (kel) sink = KSink{nullptr, kel + (size_t)81 * GH->nij * (el - e0), GH->nij};
    for (int t = 0; t < nt; ++t) phase_residual(t, nt, cfg, fr, sm, r, rel ? rel + 72 * (el - e0) : nullptr);
    for (int t = 0; t < nt; ++t) phase_tangent<MOTION>(t, cfg, fr, sm, sink);
  }
}

// elem_residual_kernel (maf_api.cu): staged residual of the listed elements (0-based), rv[27 k + comp + 3 a]
template <int MOTION>
static void run_elem_residual(const HostModel& M, const Tables& T, const double* xms, const double* cps, int32_t el,
                              double* rel) {
  const Config& cfg = M.cfg;
  const int nt = cfg.nthreads;
  std::vector<double> smem_all(cfg.smem_doubles, std::nan(""));
  int32_t* ids = reinterpret_cast<int32_t*>(smem_all.data() + 2 * cfg.front_doubles);
  double* sm = smem_all.data() + 2 * cfg.front_doubles + 2 * MAF_IDS_DOUBLES;
  double* fr = smem_all.data();
  for (int t = 0; t < nt; ++t) gather_init(t, cfg, fr);
  for (int t = 0; t < nt; ++t) gather_ids_async(t, T, el, ids);
  for (int t = 0; t < nt; ++t) gather_data_async(t, cfg, T, ids, xms, cps, fr);
  for (int t = 0; t < nt; ++t) phase_interp(t, nt, cfg, fr, sm);
  for (int t = 0; t < nt; ++t) phase_gauss<MOTION>(t, cfg, 0.0, fr, sm);
  for (int t = 0; t < nt; ++t) phase_residual(t, nt, cfg, fr, sm, nullptr, rel, true);
}

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

struct emu_model {
  HostModel M;
};

void* emu_create(const maf_mesh_desc* d, const maf_params* p, int nthreads) {
  try {
    emu_model* m = new emu_model();
    build_host_model(m->M, d, p, nthreads);
    build_host_elslot(m->M);
    return m;
  } catch (std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void emu_destroy(void* h) { delete (emu_model*)h; }
int64_t emu_nnz(void* h) { return ((emu_model*)h)->M.sym.nnz; }
void emu_pattern(void* h, int64_t* colptr1, int64_t* rowval1) {
  HostModel& M = ((emu_model*)h)->M;
  for (int64_t c = 0; c <= M.nmdf; ++c) colptr1[c] = M.sym.colptr[c] + 1;
  build_rowval(M.sym, M.ID0.data(), M.cfg.rowmask, rowval1);
}
// info: [asize, smem_doubles, ntasks, nitems, item_rounds, task_rounds, nblocks]
void emu_info(void* h, int64_t* o) {
  const Config& c = ((emu_model*)h)->M.cfg;
  o[0] = c.asize; o[1] = c.smem_doubles; o[2] = c.ntasks; o[3] = c.nitems; o[4] = c.item_rounds;
  o[5] = c.task_rounds; o[6] = c.nblocks;
}

// strip of element rows `rank` of `nranks` as the library cuts it (maf_host.h::strip_range), 1-based inclusive:
// [elements, rows of r touched, entries of nzval touched]; returns 1 when the partition is refused
int emu_strip_range(void* h, int rank, int nranks, int64_t* o6) {
  try {
    const TouchedRange R = strip_range(((emu_model*)h)->M, rank, nranks);
    o6[0] = R.e0 + 1; o6[1] = R.e1; o6[2] = R.eq_lo + 1; o6[3] = R.eq_hi; o6[4] = R.slot_lo + 1; o6[5] = R.slot_hi;
    return 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// tangent schedule: per chunk [f, g, kind, fused, first, count], then chunk_slot[task_rounds * nwarps]
int emu_chunks(void* h, int32_t* chunks6, int32_t* slots) {
  const Config& c = ((emu_model*)h)->M.cfg;
  for (int k = 0; k < c.nchunks; ++k) {
    const Block& b = c.blocks[c.chunks[k].blk];
    const int32_t v[6] = {b.f, b.g, b.kind, b.fused, c.chunks[k].first, c.chunks[k].count};
    for (int q = 0; q < 6; ++q) chunks6[6 * k + q] = v[q];
  }
  for (int q = 0; q < c.task_rounds * (c.nthreads / 32); ++q) slots[q] = c.chunk_slot[q];
  return c.nchunks;
}

// the resident-state kernels (maf_state.cuh) entry by entry, in place
void emu_state_update(void* h, const double* du, double dt, double* xms, double* cps) {
  HostModel& M = ((emu_model*)h)->M;
  Tables T = host_tables(M);
  for (int64_t k = 0; k < M.numnp * M.ndf; ++k) state_update_entry(k, M.cfg, T, du, dt, xms, cps);
}
void emu_state_predict(void* h, double dt, double* xms, const double* cps) {
  HostModel& M = ((emu_model*)h)->M;
  Tables T = host_tables(M);
  for (int64_t k = 0; k < M.numnp * 3; ++k) state_predict_entry(k, M.cfg, T, dt, xms, cps);
}

int emu_elem_v_residuals(void* h, const double* xms, const double* cps, const int64_t* el_ids, int64_t n, double* rv) {
  try {
    HostModel& M = ((emu_model*)h)->M;
    Tables T = host_tables(M);
    for (int64_t k = 0; k < n; ++k) {
      double rel[72];
      const int32_t el = (int32_t)(el_ids[k] - 1);
      switch (M.motion) {
        case M_STATIC: run_elem_residual<M_STATIC>(M, T, xms, cps, el, rel); break;
        case M_EUL: run_elem_residual<M_EUL>(M, T, xms, cps, el, rel); break;
        case M_LAG: run_elem_residual<M_LAG>(M, T, xms, cps, el, rel); break;
        case M_ALEV: run_elem_residual<M_ALEV>(M, T, xms, cps, el, rel); break;
        default: run_elem_residual<M_ALEVB>(M, T, xms, cps, el, rel); break;
      }
      for (int a = 0; a < 9; ++a)
        for (int i = 0; i < 3; ++i) rv[27 * k + 3 * a + i] = rel[9 * i + a];
    }
    return 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

int emu_assemble(void* h, const double* xms, const double* cps, double time, double dt, double bend_tm, int mode,
                 int64_t el_first, int64_t el_last, double* r, double* nzval) {
  try {
    HostModel& M = ((emu_model*)h)->M;
    Tables T = host_tables(M);
    BoundaryTables BT = host_boundary_tables(M);
    std::memset(r, 0, sizeof(double) * (size_t)M.nmdf);
    std::memset(nzval, 0, sizeof(double) * (size_t)M.sym.nnz);
    const int64_t e0 = el_first - 1, e1 = el_last;
    GatherHost GH;
    std::vector<double> kel, rel;
    if (mode == 1) {
      build_gather_host(M, GH);
      kel.assign((size_t)81 * GH.nij * (e1 - e0), std::nan(""));
      rel.assign((size_t)72 * (e1 - e0), std::nan(""));
      std::fill(r, r + M.nmdf, std::nan(""));
      std::fill(nzval, nzval + M.sym.nnz, std::nan(""));
    }
    double* kp = mode == 1 ? kel.data() : nullptr;
    double* rp = mode == 1 ? rel.data() : nullptr;
    switch (M.motion) {
      case M_STATIC: run_area<M_STATIC>(M, T, xms, cps, dt, r, nzval, kp, rp, &GH, e0, e1); break;
      case M_EUL: run_area<M_EUL>(M, T, xms, cps, dt, r, nzval, kp, rp, &GH, e0, e1); break;
      case M_LAG: run_area<M_LAG>(M, T, xms, cps, dt, r, nzval, kp, rp, &GH, e0, e1); break;
      case M_ALEV: run_area<M_ALEV>(M, T, xms, cps, dt, r, nzval, kp, rp, &GH, e0, e1); break;
      default: run_area<M_ALEVB>(M, T, xms, cps, dt, r, nzval, kp, rp, &GH, e0, e1); break;
    }
    if (mode == 1) {
      GatherTables G;
      G.nbr_ptr = M.sym.nbr_ptr.data(); G.nbr = M.sym.nbr.data(); G.pair_node = GH.pair_node.data();
      G.n2e_ptr = M.sym.n2e_ptr.data(); G.n2e = M.sym.n2e.data(); G.n2e_loc = M.sym.n2e_loc.data();
      G.ij_of = GH.ij_of.data(); G.npairs = M.sym.npairs;
      fill_gather_tables(GH, G);
      // the product gathers through the precomputed pair-contribution classes: the emulation does both and insists
      // on identical sums (the scan over the element lists is the definition, the table the fast form)
      std::vector<double> nz_scan(nzval, nzval + M.sym.nnz);
      for (int64_t p = 0; p < G.npairs; ++p)
        for (int s = 0; s < MAF_GATHER_LANES; ++s)
          gather_K_pair(p, s, M.cfg, T, G, kel.data(), GH.nij, e0, e1, nz_scan.data());
      build_pair_classes(M, GH);
      if (GH.pclass.empty()) throw std::runtime_error("emulation: pair-contribution classes were not built");
      G.pclass = GH.pclass.data(); G.eref = GH.eref.data(); G.ccnt = GH.ccnt.data(); G.cde = GH.cde.data();
      G.crow = GH.crow.data();
      for (int64_t p = 0; p < G.npairs; ++p)
        for (int s = 0; s < MAF_GATHER_LANES; ++s) gather_K_pair(p, s, M.cfg, T, G, kel.data(), GH.nij, e0, e1, nzval);
      if (std::memcmp(nz_scan.data(), nzval, sizeof(double) * (size_t)M.sym.nnz) != 0)
        throw std::runtime_error("emulation: gather through the pair classes differs from the scan");
      for (int64_t k = 0; k < M.numnp * M.ndf; ++k) gather_r_row(k, M.cfg, T, G, rel.data(), e0, e1, r);
    }
    std::vector<double> sm(B_DOUBLES);
    for (int bc = 0; bc < BT.n_neu; ++bc) {
      const double fval = neumann_value(BT.ntype[bc], BT.nval[bc], time, bend_tm);
      for (int q = BT.offs[bc]; q < BT.offs[bc + 1]; ++q) {
        const int64_t el = BT.elems[q];
        if (el < e0 || el >= e1) continue;
        std::fill(sm.begin(), sm.end(), std::nan(""));
        for (int l = 0; l < 32; ++l) boundary_gather(l, 32, M.cfg, T, BT, bc, el, xms, sm.data());
        for (int l = 0; l < 32; ++l) boundary_gauss(l, 32, M.cfg, BT, bc, fval, dt, sm.data());
        for (int l = 0; l < 32; ++l) boundary_scatter(l, 32, M.cfg, T, sm.data(), r, nzval, mode == 1);
      }
    }
    return 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return 1;
  }
}
}
