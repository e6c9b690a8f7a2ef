// Host-side check of the persistent kernels' work walk (umma_gemm.cuh: TileWalk): for every
// geometry, the workers together must visit every (row unit, column block, K split) exactly once,
// in the raster order the kernel documents.  Built and run by tests/test_host_logic.py (no GPU).
#include <cstdio>
#include <vector>

#include "../../clip_event_b200/csrc/umma_gemm.cuh"

int main() {
  int checked = 0;
  for (int nmu : {1, 2, 3, 16, 37, 144})
    for (int nnb : {1, 2, 5, 18, 144})
      for (int ksplits : {1, 2, 7})
        for (int W : {1, 3, 74, 148})
          for (int raster_n : {0, 1}) {
            const int num_tiles = nmu * nnb, num_items = num_tiles * ksplits;
            std::vector<int> seen(num_items, 0);
            for (int w = 0; w < W; ++w) {
              ce::TileWalk tw;
              tw.init(w, nmu, nnb, W, raster_n);
              int prev = -1;
              for (; tw.item < num_items; tw.next()) {
                if (tw.item <= prev || tw.item % W != w) { printf("bad item order\n"); return 1; }
                prev = tw.item;
                if (tw.mu < 0 || tw.mu >= nmu || tw.nb < 0 || tw.nb >= nnb || tw.ksp < 0 || tw.ksp >= ksplits) {
                  printf("out of range: nmu %d nnb %d ks %d W %d r %d item %d -> mu %d nb %d ksp %d\n", nmu, nnb,
                         ksplits, W, raster_n, tw.item, tw.mu, tw.nb, tw.ksp);
                  return 1;
                }
                const int tile = tw.item % num_tiles, ksp = tw.item / num_tiles;
                const int want_mu = raster_n ? tile / nnb : tile % nmu;
                const int want_nb = raster_n ? tile % nnb : tile / nmu;
                if (tw.mu != want_mu || tw.nb != want_nb || tw.ksp != ksp) {
                  printf("wrong decode: nmu %d nnb %d ks %d W %d r %d item %d -> (%d,%d,%d) want (%d,%d,%d)\n", nmu,
                         nnb, ksplits, W, raster_n, tw.item, tw.mu, tw.nb, tw.ksp, want_mu, want_nb, ksp);
                  return 1;
                }
                ++seen[(ksp * nmu + tw.mu) * nnb + tw.nb];
              }
            }
            for (int v : seen)
              if (v != 1) { printf("coverage: nmu %d nnb %d ks %d W %d r %d\n", nmu, nnb, ksplits, W, raster_n); return 1; }
            ++checked;
          }
  printf("tilewalk ok: %d geometries\n", checked);
  return 0;
}
