// Host-only driver for the integer index maps of csrc/remap.cuh (compiled by tests/test_remap_cpu.py with nvcc, run on
// the CPU): prints, for one stage geometry, the slot -> token table of every window as the attention kernel's TMA boxes
// deliver it, next to key_slot_377 / slot_token_377, so that the Python test can compare them with the oracle's tables.
#include <cstdio>
#include <cstdlib>

#include "../../vqa-lrce-kbs-2023_b200/csrc/remap.cuh"

using namespace lrce;

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  StageGeom g;
  g.D = 3; g.H = atoi(argv[1]); g.W = atoi(argv[2]);
  g.wd = 3; g.wh = 7; g.ww = 7; g.sd = 0; g.sh = atoi(argv[3]); g.sw = atoi(argv[4]);
  const int nh = g.H / 7, nw = g.W / 7;
  for (int t = 0; t < 147; ++t) printf("%d ", key_slot_377(t));
  printf("\n");
  for (int s = 0; s < 160; ++s) printf("%d ", slot_token_377(s));
  printf("\n");
  for (int hW = 0; hW < nh; ++hW)
    for (int wW = 0; wW < nw; ++wW) {
      const WindowBoxes377 b = window_boxes_377(g, hW, wW);
      // no box may cross the frame border: TMA boxes do not wrap
      if (b.xa + 4 > g.W || b.xb + 3 > g.W || b.ya + 4 > g.H || b.yb + 3 > g.H) return 3;
      for (int s = 0; s < 160; ++s) printf("%d ", box_slot_token_377(g, b, s));
      printf("\n");
      for (int t = 0; t < 147; ++t) printf("%d ", window_source_token_377(g, hW, wW, t));
      printf("\n");
      for (int t = 0; t < 147; ++t) printf("%d ", shift_region_id(g, hW * nw + wW, t));
      printf("\n");
    }
  return 0;
}
