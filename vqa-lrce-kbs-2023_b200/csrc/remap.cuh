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
