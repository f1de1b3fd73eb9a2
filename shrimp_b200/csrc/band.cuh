// Band geometry of the full Smith-Waterman kernels (common/anchors.c) shared by sw_full.cu and
// sw_full_cs.cu.
#pragma once
#include "stages.cuh"

namespace shrimp {

#define NEG_HALF (-1073741823)  // -INT_MAX/2, init_cell sw-full-ls.c:66-81

struct Rect {
  long long x, y;
  int length, width;
};

__device__ __forceinline__ void rect_x_range(const Rect &a, int x_len, int y, int &x_min, int &x_max) {
  // anchor_get_x_range, anchors.c:66-95
  if (y < a.y) x_min = 0;
  else if (y <= a.y + (a.length - 1)) x_min = (int)(a.x + (y - a.y));
  else x_min = (int)(a.x + a.length);
  if (x_min < 0) x_min = 0;
  if (x_min >= x_len) x_min = x_len - 1;
  if (y < a.y - (a.width - 1)) x_max = (int)(a.x + (a.width - 1) - 1);
  else if (y <= a.y - (a.width - 1) + (a.length - 1)) x_max = (int)(a.x + (a.width - 1) + (y - (a.y - (a.width - 1))));
  else x_max = x_len - 1;
  if (x_max < 0) x_max = 0;
  if (x_max >= x_len) x_max = x_len - 1;
}

__device__ __forceinline__ Rect rect_join2(long long x0, long long y0, int l0, int w0, long long x1, long long y1,
                                           int l1, int w1) {
  // anchor_join, anchors.c:9-54
  long long nw0 = x0 + y0, sw0 = x0 - y0, ne0 = sw0 + 2 * (w0 - 1), se0 = nw0 + 2 * (l0 - 1);
  long long nw1 = x1 + y1, sw1 = x1 - y1, ne1 = sw1 + 2 * (w1 - 1), se1 = nw1 + 2 * (l1 - 1);
  long long nw_min = nw0 < nw1 ? nw0 : nw1, sw_min = sw0 < sw1 ? sw0 : sw1;
  long long ne_max = ne0 > ne1 ? ne0 : ne1, se_max = se0 > se1 ? se0 : se1;
  Rect r;
  if ((nw_min + sw_min) % 2 != 0) nw_min--;
  r.x = (nw_min + sw_min) / 2;
  r.y = nw_min - r.x;
  if ((ne_max - sw_min) % 2 != 0) ne_max++;
  r.width = (int)((ne_max - sw_min) / 2 + 1);
  if ((se_max - nw_min) % 2 != 0) se_max++;
  r.length = (int)((se_max - nw_min) / 2 + 1);
  return r;
}


// the band rectangle of one task: joined anchor widened by anchor_width, or the threshold band
// (sw-full-ls.c:175-191, sw-full-cs.c:285-303)
__device__ __forceinline__ Rect task_rect(const FullTask &T, int anchor_width, int match, bool use_anchor) {
  Rect rect;
  if (use_anchor && anchor_width >= 0) {
    rect.x = T.ax;
    rect.y = T.ay;
    rect.length = T.alen;
    rect.width = T.awidth;
    rect.x -= anchor_width / 2;  // anchor_widen, anchors.c:57-63
    rect.y += anchor_width / 2;
    rect.width += anchor_width;
  } else {
    long long y0 = (T.rlen * match - T.thresh) / match;
    rect = rect_join2(0, y0, 1, 1, T.glen - 1, T.rlen - 1 - y0, 1, 1);
  }
  return rect;
}

}  // namespace shrimp
