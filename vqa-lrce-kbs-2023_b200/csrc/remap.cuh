// remap.cuh — integer index maps of the Video Swin path, shared by host and device so that the standalone gather
// kernels, the attention kernel and lrce_remap_index() all use the SAME arithmetic (bit-exact contract).
//
//   window_source_token : window_partition(roll(x, -shift))  (video_swin_ori.py:262, :60-72) and, read as a scatter,
//                         roll(window_reverse(y), +shift)     (video_swin_ori.py:75-88, :276)
//   shift_region_id     : region image of compute_mask        (video_swin_ori.py:346-356) for the shifted axes H, W
//   rel_pos_offset      : f(t) with relative_position_index[i][j] = f(i) - f(j) + const (video_swin_ori.py:134-147)
//   merge_source_token  : PatchMerging 2x2 gather             (video_swin_ori.py:333-337)
#pragma once
#include <cuda_runtime.h>

namespace lrce {

struct StageGeom {
  int D, H, W;     // token grid of one segment
  int wd, wh, ww;  // effective (clamped) window, e.g. (3,7,7)
  int sd, sh, sw;  // cyclic shift of this block (0 when not shifted / clamped)
};

__host__ __device__ inline int window_tokens(const StageGeom& g) { return g.wd * g.wh * g.ww; }
__host__ __device__ inline int windows_per_segment(const StageGeom& g) {
  return (g.D / g.wd) * (g.H / g.wh) * (g.W / g.ww);
}

// flat (d*H + h)*W + w index, inside one segment, of token `tok` of window `win`
__host__ __device__ inline int window_source_token(const StageGeom& g, int win, int tok) {
  const int nh = g.H / g.wh, nw = g.W / g.ww;
  const int wW = win % nw, hW = (win / nw) % nh, dW = win / (nw * nh);
  const int w = tok % g.ww, h = (tok / g.ww) % g.wh, d = tok / (g.ww * g.wh);
  const int sd_ = (dW * g.wd + d + g.sd) % g.D;
  const int sh_ = (hW * g.wh + h + g.sh) % g.H;
  const int sw_ = (wW * g.ww + w + g.sw) % g.W;
  return (sd_ * g.H + sh_) * g.W + sw_;
}

// The same map specialised for LRCE's clamped (3,7,7) window on a D == 3 grid (every Swin stage of the path): constant
// divisors and a conditional subtract instead of ~10 integer divisions per token. (hW, wW) are the window's coordinates
// in the window grid: wW = win % (W/7), hW = (win / (W/7)) % (H/7). lrce_remap_index() and the standalone remap kernel
// use this function whenever the geometry allows, so the bit-exact tests pin exactly what the attention kernel executes.
__host__ __device__ inline bool is_window_377(const StageGeom& g) {
  return g.D == 3 && g.wd == 3 && g.wh == 7 && g.ww == 7 && g.sd == 0;
}
__host__ __device__ inline int window_source_token_377(const StageGeom& g, int hW, int wW, int tok) {
  const int d = tok / 49, rem = tok - d * 49, h = rem / 7, w = rem - h * 7;
  int y = hW * 7 + h + g.sh, x = wW * 7 + w + g.sw;  // shift < window <= extent: one wrap at most
  if (y >= g.H) y -= g.H;
  if (x >= g.W) x -= g.W;
  return (d * g.H + y) * g.W + x;
}

// Key order inside the attention kernel's score tile for the (3,7,7) window. Keys are grouped by their shift-mask class
// k = 2 [h >= 4] + [w >= 4] (the seam of a shift-3 border window, see shift_region_id) so that the mask is constant over
// every 8-column group of the tile: class 0 -> columns [0, 48), class 1 -> [48, 84) (+4 pads), class 2 -> [88, 124)
// (+4 pads), class 3 -> [128, 155) (+5 pads). Softmax and P v are invariant to the key order; only K/V staging rows and
// the dense bias table's columns use it.
__host__ __device__ inline int key_class_377(int tok) {
  const int rem = tok % 49, h = rem / 7, w = rem - h * 7;
  return (h >= 4 ? 2 : 0) + (w >= 4 ? 1 : 0);
}
__host__ __device__ inline int key_slot_377(int tok) {
  const int d = tok / 49, rem = tok - d * 49, h = rem / 7, w = rem - h * 7;
  const bool ch = h >= 4, cw = w >= 4;
  const int nh_c = ch ? 3 : 4, nw_c = cw ? 3 : 4;
  const int idx = (d * nh_c + (ch ? h - 4 : h)) * nw_c + (cw ? w - 4 : w);
  const int base = ch ? (cw ? 128 : 88) : (cw ? 48 : 0);
  return base + idx;
}
// inverse of key_slot_377: window token of row / key slot `slot` (0..159), -1 for the 13 pad slots. Inside a class the slots
// run (d, h', w') with w' fastest, which is exactly the order in which a TMA box (32 channels, w-part, h-part, 3 frames)
// lands in shared memory: the attention kernel loads each class with ONE box per operand.
__host__ __device__ inline int slot_token_377(int slot) {
  const int cls = (slot >= 48) + (slot >= 88) + (slot >= 128);
  const bool ch = cls >= 2, cw = (cls & 1) != 0;
  const int base = ch ? (cw ? 128 : 88) : (cw ? 48 : 0);
  const int nh_c = ch ? 3 : 4, nw_c = cw ? 3 : 4;
  const int idx = slot - base;
  if (idx >= 3 * nh_c * nw_c) return -1;
  const int w = idx % nw_c, h = (idx / nw_c) % nh_c, d = idx / (nw_c * nh_c);
  return d * 49 + (h + (ch ? 4 : 0)) * 7 + w + (cw ? 4 : 0);
}
// The four TMA boxes of window (hW, wW): origins of the parts before (a) / behind (b) the shift seam along h and w in the
// UNROLLED frame. Part a is 4 long, part b 3 long; only part b can wrap around the frame border, and it wraps as a whole.
// Mask class k = 2 [h part b] + [w part b] is box (w part, h part) = (k & 1, k >> 1) with 3 frames, landing at slot base
// 0 / 48 / 88 / 128.
struct WindowBoxes377 {
  int xa, xb, ya, yb;
};
__host__ __device__ inline WindowBoxes377 window_boxes_377(const StageGeom& g, int hW, int wW) {
  WindowBoxes377 b;
  b.ya = hW * 7 + g.sh;
  b.xa = wW * 7 + g.sw;
  b.yb = b.ya + 4;
  b.xb = b.xa + 4;
  if (b.yb >= g.H) b.yb -= g.H;
  if (b.xb >= g.W) b.xb -= g.W;
  return b;
}
// flat (d*H + h)*W + w token of row / key slot `slot` as the boxes deliver it (-1 for a pad slot): box traversal order is
// w fastest, then h, then frame
__host__ __device__ inline int box_slot_token_377(const StageGeom& g, const WindowBoxes377& b, int slot) {
  const int cls = (slot >= 48) + (slot >= 88) + (slot >= 128);
  const bool ch = cls >= 2, cw = (cls & 1) != 0;
  const int base = ch ? (cw ? 128 : 88) : (cw ? 48 : 0);
  const int nh_c = ch ? 3 : 4, nw_c = cw ? 3 : 4;
  const int idx = slot - base;
  if (idx >= 3 * nh_c * nw_c) return -1;
  const int w = idx % nw_c, h = (idx / nw_c) % nh_c, d = idx / (nw_c * nh_c);
  return (d * g.H + (ch ? b.yb : b.ya) + h) * g.W + (cw ? b.xb : b.xa) + w;
}

// mask class of the 8-column group `grp` (0..19) of the score tile
__host__ __device__ inline int key_group_class_377(int grp) { return (grp >= 6) + (grp >= 11) + (grp >= 16); }

// region id (0..8) of token `tok` of window `win` in shifted coordinates; tokens attend each other iff ids are equal.
// Along a shifted axis of length L the slices [0, L-win), [L-win, L-shift), [L-shift, L) carry ids 0, 1, 2.
__host__ __device__ inline int shift_region_id(const StageGeom& g, int win, int tok) {
  const int nh = g.H / g.wh, nw = g.W / g.ww;
  const int wW = win % nw, hW = (win / nw) % nh;
  const int w = tok % g.ww, h = (tok / g.ww) % g.wh;
  const int ph = hW * g.wh + h, pw = wW * g.ww + w;  // position in the rolled frame
  const int rh = g.sh ? ((ph >= g.H - g.wh) + (ph >= g.H - g.sh)) : 0;
  const int rw = g.sw ? ((pw >= g.W - g.ww) + (pw >= g.W - g.sw)) : 0;
  return rh * 3 + rw;
}

// f(t) = 169 d + 13 h + w for the configured (8,7,7) window: index(i, j) = f(i) - f(j) + (7*169 + 6*13 + 6)
__host__ __device__ inline int rel_pos_offset(const StageGeom& g, int tok) {
  const int w = tok % g.ww, h = (tok / g.ww) % g.wh, d = tok / (g.ww * g.wh);
  return d * 169 + h * 13 + w;
}
constexpr int REL_POS_CENTER = 7 * 169 + 6 * 13 + 6;

// source token (flat index in the (D,H,W) grid) of part k in {0,1,2,3} of merged token `out_tok` in (D,H/2,W/2)
__host__ __device__ inline int merge_source_token(int D, int H, int W, int out_tok, int k) {
  const int W2 = W / 2, H2 = H / 2;
  const int j = out_tok % W2, i = (out_tok / W2) % H2, d = out_tok / (W2 * H2);
  return (d * H + 2 * i + (k & 1)) * W + 2 * j + (k >> 1);
}

}  // namespace lrce
