// Weight operand transforms shared by the packing kernels of the bf16 path (conv.cu) and the parity path
// (conv_parity.cu): element (tap, co, ci) of the GEMM operand as a function of the nn.Conv2d weight.
#pragma once

namespace ryolo {

// mode 0: plain            packed[tap][co][ci] = w[co][ci][kh][kw]
// mode 1: dgrad of mode 0  packed[tap][co][ci] = w[ci][co][k-1-kh][k-1-kw]      (w = the FORWARD weight [cin_d, cout_d, k, k])
// mode 2: space-to-depth   packed[(qy,qx)][co][(py*2+px)*C + c] = w[co][c][kh(qy,py)][kw(qx,px)] or 0   (w = [cout, C, 3, 3])
// mode 3: dgrad of mode 2  packed[(ty,tx)][(py*2+px)*C + c][ci] = w[ci][c][kh(1-ty,py)][kw(1-tx,px)] or 0
__device__ __forceinline__ int s2d_k(int q, int p) { return q == 0 ? (p == 1 ? 0 : -1) : (p == 0 ? 1 : 2); }

// caller guarantees ci < cin, co < cout; ks = taps per side of the PACKED operand (2 for the s2d modes)
__device__ __forceinline__ float pack_value(const float* __restrict__ w, int cout, int cin, int ks, int mode, int tap, int co,
                                            int ci) {
  const int ty = tap / ks, tx = tap % ks;
  if (mode == 0) return w[(((size_t)co * cin + ci) * ks + ty) * ks + tx];
  if (mode == 1) return w[(((size_t)ci * cout + co) * ks + (ks - 1 - ty)) * ks + (ks - 1 - tx)];
  if (mode == 2) {
    const int C = cin >> 2, ph = ci / C, c = ci - ph * C;
    const int kh = s2d_k(ty, ph >> 1), kw = s2d_k(tx, ph & 1);
    return (kh >= 0 && kw >= 0) ? w[(((size_t)co * C + c) * 3 + kh) * 3 + kw] : 0.f;
  }
  const int C = cout >> 2, ph = co / C, c = co - ph * C;
  const int kh = s2d_k(1 - ty, ph >> 1), kw = s2d_k(1 - tx, ph & 1);
  return (kh >= 0 && kw >= 0) ? w[(((size_t)ci * C + c) * 3 + kh) * 3 + kw] : 0.f;
}

}  // namespace ryolo
