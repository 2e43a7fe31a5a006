// prints the tangent-chunk plan of a motion: g++ -O1 -std=c++17 -I membranealefem.jl_b200/csrc tools/print_plan.cpp
#include <cstdio>
#include "maf_config.h"
using namespace maf;
int main(int argc, char** argv) {
  const int motion = argc > 1 ? atoi(argv[1]) : M_ALEVB;
  int32_t dofs[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  int ndf = 8;
  if (motion == M_LAG || motion == M_STATIC) { dofs[3] = dofs[4] = dofs[5] = 0; dofs[6] = 4; dofs[7] = 0; ndf = 4; }
  if (motion == M_EUL) { dofs[7] = 0; ndf = 7; }
  static Config cfg;
  build_config(cfg, motion, ndf, dofs, 1.0, -0.5, 1.0, 0.0, 4096.0, 1.0, 0, MAF_NT);
  printf("asize %d smem %d doubles ntasks %d nchunks %d rounds %d items %d item_rounds %d\n", cfg.asize, cfg.smem_doubles,
         cfg.ntasks, cfg.nchunks, cfg.task_rounds, cfg.nitems, cfg.item_rounds);
  for (int k = 0; k < cfg.nchunks; ++k) {
    const Chunk& c = cfg.chunks[k];
    const Block& b = cfg.blocks[c.blk];
    int w = -1, r = -1;
    for (int q = 0; q < MAF_MAX_ROUNDS * 8; ++q)
      if (cfg.chunk_slot[q] == k) { r = q / (MAF_NT / 32); w = q % (MAF_NT / 32); }
    printf("chunk %2d blk %2d (f%d,g%d) nr%d nc%d kind %d fused %d tr %d first %3d count %2d cost %3d -> warp %d round %d\n", k, c.blk,
           b.f, b.g, b.nr, b.nc, b.kind, b.fused, b.tr, c.first, c.count, kind_cost(b.kind, b.fused, b.tr), w, r);
  }
}
