// Window geometry and sampling-coordinate math shared by the RVSA forward and backward kernels.
// SURVEY.md Appendix A.1; [V]:298-342 (padding, reference grid, base coordinates), [V]:372-385 (scale, rotate, shift).
#pragma once
#include <cuda_runtime.h>

namespace mtp {

constexpr int WS = 7;            // window size ([V]:629 hard-wires (7, 7))
constexpr int NTOK = WS * WS;    // 49 tokens per window
constexpr int HD = 64;           // head dim of ViT-B/L (768/12 = 1024/16)

struct RvsaGeom {
  int B, h, w, C, nH;
  int pt, pl;        // top / left zero padding ([V]:300-303: pad//2 on top/left, the rest bottom/right)
  int Hq, Wq;        // padded grid
  int nh, nw;        // windows per column / row
};

inline RvsaGeom make_rvsa_geom(int B, int h, int w, int C, int nH) {
  RvsaGeom g;
  g.B = B; g.h = h; g.w = w; g.C = C; g.nH = nH;
  const int pd_h = (WS - h % WS) % WS, pd_w = (WS - w % WS) % WS;
  g.pt = pd_h / 2; g.pl = pd_w / 2;
  g.Hq = h + pd_h; g.Wq = w + pd_w;
  g.nh = g.Hq / WS; g.nw = g.Wq / WS;
  return g;
}

// Pixel coordinates (in the PADDED grid) sampled by window token (iy, ix) of window (wy, wx) for one head.
__device__ __forceinline__ void rvsa_sample_coord(const RvsaGeom& g, int wy, int wx, int iy, int ix, float ox, float oy, float sx,
                                                  float sy, float th, float& px, float& py) {
  const float inv_w = 2.0f / (float)(g.Wq - 1), inv_h = 2.0f / (float)(g.Hq - 1);
  const float refx = -1.0f + (float)(wx * WS + WS / 2) * inv_w;      // mean of linspace(-1,1,W')[7wx:7wx+7]
  const float refy = -1.0f + (float)(wy * WS + WS / 2) * inv_h;
  const float X = (1.0f + sx) * ((float)(ix - WS / 2) * inv_w);
  const float Y = (1.0f + sy) * ((float)(iy - WS / 2) * inv_h);
  float s, c;
  sincosf(th, &s, &c);
  const float cx = refx + X * c - Y * s + ox;
  const float cy = refy + Y * c + X * s + oy;
  px = (cx + 1.0f) * 0.5f * (float)(g.Wq - 1);
  py = (cy + 1.0f) * 0.5f * (float)(g.Hq - 1);
}

}  // namespace mtp
