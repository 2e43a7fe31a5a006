// maf_symbolic.h -- one-time symbolic phase: sparsity pattern of K and the element -> nnz-slot maps.
//
// The pattern is  P = U_e (active LM rows of e) x (active LM cols of e)  (FiniteElement.jl:107,129-136), optionally
// minus the dof blocks that are identically zero by the equations (Config::rowmask). Because the reference numbers
// unknowns node-major (Mesh.jl:276-284), column (B,J) of the CSC holds, for every node A adjacent to B in ascending
// order, the active row dofs of A in ascending order. That lets a slot be computed from three small tables:
//     slot(A,I ; B,J) = colptr[ID[J,B]] + pairoff[pair(A,B)][J] + popcount(nodemask[A] & rowmask[J] & ((1<<I)-1))
// Pure C++ (threads), shared by the library and by the CPU emulation harness in tests/.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <thread>
#include <vector>

namespace maf {

struct Symbolic {
  int64_t numel = 0, numnp = 0, nmdf = 0, nnz = 0, npairs = 0;
  int ndf = 0;
  std::vector<int64_t> nbr_ptr;    // numnp+1
  std::vector<int32_t> nbr;        // npairs: sorted neighbour nodes of every node
  std::vector<uint8_t> nodemask;   // numnp
  std::vector<uint8_t> pairoff;    // npairs x 8
  std::vector<int64_t> colptr;     // nmdf+1 (0-based)
  std::vector<int32_t> elpair;     // numel x 81, [a][b] -> index of A in nbr(B), global pair index
  std::vector<int64_t> n2e_ptr;    // numnp+1   node -> elements (ascending element id)
  std::vector<int32_t> n2e;        // 9 numel entries: element id
  std::vector<uint8_t> n2e_loc;    // local node index of the node in that element
  std::vector<int32_t> eq0;        // numnp: first active equation at or after the node (nmdf if none): colptr[eq0[n]]
                                   // is a lower bound of every column pointer of nodes >= n
};

template <class F> inline void parallel_for(int64_t n, F&& fn) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
  if (n < 4096) nt = 1;
  std::vector<std::thread> th;
  const int64_t chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int64_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([=, &fn]() { fn(lo, hi); });
  }
  for (auto& x : th) x.join();
}

// IX0: 9 x numel 0-based node ids; ID0: ndf x numnp 0-based equation number or -1
inline void build_symbolic(Symbolic& S, int64_t numel, int64_t numnp, int ndf, int64_t nmdf, const int32_t* IX0,
                           const int32_t* ID0, const uint8_t rowmask[8]) {
  S.numel = numel; S.numnp = numnp; S.ndf = ndf; S.nmdf = nmdf;
  // node -> elements
  S.n2e_ptr.assign(numnp + 1, 0);
  for (int64_t k = 0; k < 9 * numel; ++k) {
    if (IX0[k] < 0 || IX0[k] >= numnp) throw std::runtime_error("IX holds a node id outside 1..numnp");
    S.n2e_ptr[IX0[k] + 1] += 1;
  }
  for (int64_t n = 0; n < numnp; ++n) S.n2e_ptr[n + 1] += S.n2e_ptr[n];
  S.n2e.resize(9 * numel);
  S.n2e_loc.resize(9 * numel);
  {
    std::vector<int64_t> cur(S.n2e_ptr.begin(), S.n2e_ptr.end() - 1);
    for (int64_t e = 0; e < numel; ++e)
      for (int a = 0; a < 9; ++a) {
        const int64_t p = cur[IX0[9 * e + a]]++;
        S.n2e[p] = (int32_t)e;
        S.n2e_loc[p] = (uint8_t)a;
      }
  }
  // node adjacency (two passes: degrees, then fill)
  auto neighbours = [&](int64_t B, int32_t* buf) -> int {
    int n = 0;
    for (int64_t q = S.n2e_ptr[B]; q < S.n2e_ptr[B + 1]; ++q) {
      const int32_t* ix = IX0 + 9 * (int64_t)S.n2e[q];
      for (int a = 0; a < 9; ++a) buf[n++] = ix[a];
      if (n > 9 * 28) throw std::runtime_error("node belongs to too many elements");
    }
    std::sort(buf, buf + n);
    return (int)(std::unique(buf, buf + n) - buf);
  };
  S.nbr_ptr.assign(numnp + 1, 0);
  parallel_for(numnp, [&](int64_t lo, int64_t hi) {
    int32_t buf[9 * 32];
    for (int64_t B = lo; B < hi; ++B) S.nbr_ptr[B + 1] = neighbours(B, buf);
  });
  for (int64_t n = 0; n < numnp; ++n) S.nbr_ptr[n + 1] += S.nbr_ptr[n];
  S.npairs = S.nbr_ptr[numnp];
  if (S.npairs >= (int64_t)1 << 31) throw std::runtime_error("node-pair count exceeds int32");
  S.nbr.resize(S.npairs);
  S.nodemask.assign(numnp, 0);
  for (int64_t n = 0; n < numnp; ++n) {
    unsigned m = 0;
    for (int d = 0; d < ndf; ++d)
      if (ID0[(int64_t)ndf * n + d] >= 0) m |= 1u << d;
    S.nodemask[n] = (uint8_t)m;
  }
  S.eq0.assign(numnp, (int32_t)nmdf);
  {
    int32_t next = (int32_t)nmdf;
    for (int64_t n = numnp - 1; n >= 0; --n) {
      for (int d = ndf - 1; d >= 0; --d)
        if (ID0[(int64_t)ndf * n + d] >= 0) next = ID0[(int64_t)ndf * n + d];
      S.eq0[n] = next;
    }
  }
  S.pairoff.assign((size_t)S.npairs * 8, 0);
  std::vector<int64_t> colcount(nmdf + 1, 0);
  bool overflow = false;
  parallel_for(numnp, [&](int64_t lo, int64_t hi) {
    int32_t buf[9 * 32];
    for (int64_t B = lo; B < hi; ++B) {
      const int n = neighbours(B, buf);
      const int64_t p0 = S.nbr_ptr[B];
      int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int q = 0; q < n; ++q) {
        S.nbr[p0 + q] = buf[q];
        const unsigned m = S.nodemask[buf[q]];
        for (int J = 0; J < ndf; ++J) {
          if (cnt[J] > 255) overflow = true;
          S.pairoff[(size_t)(p0 + q) * 8 + J] = (uint8_t)cnt[J];
          cnt[J] += __builtin_popcount(m & rowmask[J]);
        }
      }
      for (int J = 0; J < ndf; ++J) {
        const int32_t eq = ID0[(int64_t)ndf * B + J];
        if (eq >= 0) colcount[eq + 1] = cnt[J];
      }
    }
  });
  if (overflow) throw std::runtime_error("a column of K has more than 255 rows before a node block (pairoff is uint8)");
  S.colptr.assign(nmdf + 1, 0);
  for (int64_t c = 0; c < nmdf; ++c) S.colptr[c + 1] = S.colptr[c] + colcount[c + 1];
  S.nnz = S.colptr[nmdf];
  // element -> pair index
  S.elpair.resize((size_t)81 * numel);
  parallel_for(numel, [&](int64_t lo, int64_t hi) {
    for (int64_t e = lo; e < hi; ++e) {
      const int32_t* ix = IX0 + 9 * e;
      for (int b = 0; b < 9; ++b) {
        const int64_t p0 = S.nbr_ptr[ix[b]], p1 = S.nbr_ptr[ix[b] + 1];
        for (int a = 0; a < 9; ++a) {
          const int32_t* it = std::lower_bound(S.nbr.data() + p0, S.nbr.data() + p1, ix[a]);
          S.elpair[(size_t)81 * e + 9 * a + b] = (int32_t)(it - S.nbr.data());
        }
      }
    }
  });
}

// rows of the CSC pattern, 1-based (SparseMatrixCSC rowval); colptr1 = colptr + 1
inline void build_rowval(const Symbolic& S, const int32_t* ID0, const uint8_t rowmask[8], int64_t* rowval1) {
  const int ndf = S.ndf;
  parallel_for(S.numnp, [&](int64_t lo, int64_t hi) {
    for (int64_t B = lo; B < hi; ++B)
      for (int J = 0; J < ndf; ++J) {
        const int32_t eq = ID0[(int64_t)ndf * B + J];
        if (eq < 0) continue;
        int64_t k = S.colptr[eq];
        for (int64_t p = S.nbr_ptr[B]; p < S.nbr_ptr[B + 1]; ++p) {
          const int32_t A = S.nbr[p];
          const unsigned m = S.nodemask[A] & rowmask[J];
          for (int I = 0; I < ndf; ++I)
            if ((m >> I) & 1u) rowval1[k++] = (int64_t)ID0[(int64_t)ndf * A + I] + 1;
        }
      }
  });
}

// rows of the columns [c_lo, c_hi) only (0-based column range), written from rowval1[0]
inline void build_rowval_columns(const Symbolic& S, const int32_t* ID0, const uint8_t rowmask[8], int64_t c_lo,
                                 int64_t c_hi, int64_t* rowval1) {
  const int ndf = S.ndf;
  const int64_t base = S.colptr[c_lo];
  // unknowns are numbered node-major: eq0 (first active equation at or after a node) is monotone, so the nodes that
  // own the columns [c_lo, c_hi) are a contiguous range
  int64_t n_lo = std::lower_bound(S.eq0.begin(), S.eq0.end(), (int32_t)c_lo) - S.eq0.begin();
  int64_t n_hi = std::lower_bound(S.eq0.begin(), S.eq0.end(), (int32_t)c_hi) - S.eq0.begin();
  n_lo = std::max<int64_t>(0, n_lo - 1);
  n_hi = std::min<int64_t>(S.numnp, n_hi + 1);
  parallel_for(n_hi - n_lo, [&](int64_t lo, int64_t hi) {
    for (int64_t B = n_lo + lo; B < n_lo + hi; ++B)
      for (int J = 0; J < ndf; ++J) {
        const int32_t eq = ID0[(int64_t)ndf * B + J];
        if (eq < c_lo || eq >= c_hi) continue;
        int64_t k = S.colptr[eq] - base;
        for (int64_t p = S.nbr_ptr[B]; p < S.nbr_ptr[B + 1]; ++p) {
          const int32_t A = S.nbr[p];
          const unsigned m = S.nodemask[A] & rowmask[J];
          for (int I = 0; I < ndf; ++I)
            if ((m >> I) & 1u) rowval1[k++] = (int64_t)ID0[(int64_t)ndf * A + I] + 1;
        }
      }
  });
}

}  // namespace maf
