// Dense part of compute_loss (reference model/loss.py:340-348): the objectness term
// BCEWithLogitsLoss(pos_weight)(pi[..., 5], tobj) runs over EVERY cell of the three head maps (64 x 72 x (19^2 + 38^2 +
// 76^2) = 35 M cells per step).  Written as framework element-wise ops it costs ~20 passes over strided views of the head
// tensors (forward + autograd) plus a permute copy each way -- ~4 ms of a 66 ms step.  Here: one kernel for the loss sum,
// one kernel that writes the WHOLE head cotangent (objectness channel = d(loss)/dx, every other channel zero; the few
// matched rows are scattered on top by the caller), both reading the head tensor through its strides so that the NCHW
// buffer the head convolution wrote is consumed in place (no permute().contiguous() copy), and a transposing
// NCHW fp32 -> padded-NHWC bf16 pass for the cotangent on its way into the head convolutions' dgrad / wgrad.
//
// BCE with logits, PyTorch's formulation (aten/native/Loss.cpp binary_cross_entropy_with_logits):
//   lw = 1 + (pos_weight - 1) t;   l = (1 - t) x + lw (log1p(exp(-|x|)) + max(-x, 0));   dl/dx = (1 - t) + lw (sigmoid(x) - 1)
#include <cuda_bf16.h>

#include "common.cuh"

namespace ryolo {

struct HeadView {            // logical [B, na, ny, nx, no] tensor with arbitrary element strides
  long long sb, sa, sy, sx, sc;
  int B, na, ny, nx, no;
};

__device__ __forceinline__ size_t cell_offset(const HeadView& v, size_t cell) {
  const int x = (int)(cell % v.nx);
  size_t r = cell / v.nx;
  const int y = (int)(r % v.ny);
  r /= v.ny;
  const int a = (int)(r % v.na);
  const int b = (int)(r / v.na);
  return (size_t)(b * v.sb + a * v.sa + y * v.sy + x * v.sx);
}

__global__ void __launch_bounds__(256) obj_bce_fwd_kernel(const float* __restrict__ x, HeadView v, int ch,
                                                          const float* __restrict__ tobj, float pw, double* __restrict__ out) {
  const size_t n = (size_t)v.B * v.na * v.ny * v.nx;
  double acc = 0.0;
  for (size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cell < n; cell += (size_t)gridDim.x * blockDim.x) {
    const float xv = x[cell_offset(v, cell) + (size_t)ch * v.sc];
    const float t = tobj[cell];
    const float lw = 1.f + (pw - 1.f) * t;
    acc += (double)((1.f - t) * xv + lw * (log1pf(expf(-fabsf(xv))) + fmaxf(-xv, 0.f)));
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += part[w];
    atomicAdd(out, s);
  }
}

// g[cell, k] = k == ch ? scale * dl/dx : 0, for every channel k (g has the strides of the head tensor)
__global__ void __launch_bounds__(256) obj_bce_bwd_kernel(const float* __restrict__ x, HeadView v, int ch,
                                                          const float* __restrict__ tobj, float pw,
                                                          const float* __restrict__ scale_dev, float* __restrict__ g) {
  const size_t n = (size_t)v.B * v.na * v.ny * v.nx;
  const float scale = __ldg(scale_dev);
  for (size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cell < n; cell += (size_t)gridDim.x * blockDim.x) {
    const size_t off = cell_offset(v, cell);
    const float xv = x[off + (size_t)ch * v.sc];
    const float t = tobj[cell];
    const float lw = 1.f + (pw - 1.f) * t;
    const float sg = 1.f / (1.f + expf(-xv));
    const float d = scale * ((1.f - t) + lw * (sg - 1.f));
    for (int k = 0; k < v.no; k++) g[off + (size_t)k * v.sc] = k == ch ? d : 0.f;
  }
}

// NCHW fp32 [B, C, ny, nx] -> padded-NHWC bf16 (channel stride dcs).  One block = 32 consecutive x of one image row, all
// channels: 128-byte coalesced reads per channel, transposed through shared memory, 16-byte stores per pixel.
__device__ __forceinline__ size_t pad_off(int b, int y, int x, int h, int w, int cs) {
  return (((size_t)b * (h + 2) + y + 1) * (w + 2) + x + 1) * cs;
}
__global__ void __launch_bounds__(256) head_grad_nchw_kernel(const float* __restrict__ g, int c, int ny, int nx,
                                                             __nv_bfloat16* __restrict__ dst, int dcs) {
  extern __shared__ __align__(16) unsigned char hg_smem[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(hg_smem);   // [32][cpad]
  const int cpad = ((c + 7) & ~7) + 8;                                 // + 8: rows 16 B apart modulo the bank cycle
  const int x0 = blockIdx.x * 32, y = blockIdx.y, b = blockIdx.z;
  const int npx = min(32, nx - x0);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int ch = wid; ch < c; ch += 8)
    if (lane < npx) tile[lane * cpad + ch] = __float2bfloat16_rn(__ldg(g + (((size_t)b * c + ch) * ny + y) * nx + x0 + lane));
  __syncthreads();
  const int chunks = c >> 3;
  for (int idx = threadIdx.x; idx < npx * chunks; idx += 256) {
    const int px = idx / chunks, k = idx - px * chunks;
    *reinterpret_cast<uint4*>(dst + pad_off(b, y, x0 + px, ny, nx, dcs) + k * 8) =
        *reinterpret_cast<const uint4*>(tile + px * cpad + k * 8);
  }
  const int tail = c & 7;
  for (int idx = threadIdx.x; idx < npx * tail; idx += 256) {
    const int px = idx / tail, e = (c & ~7) + idx % tail;
    dst[pad_off(b, y, x0 + px, ny, nx, dcs) + e] = tile[px * cpad + e];
  }
}

static inline int loss_grid(size_t n) {
  const size_t want = (n + 255) / 256;
  const size_t cap = (size_t)device_sm_count() * 8;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace ryolo

using namespace ryolo;

static inline bool mk_view(const long long* strides, int batch, int na, int ny, int nx, int no, HeadView* v) {
  if (!strides || batch <= 0 || na <= 0 || ny <= 0 || nx <= 0 || no <= 0) return false;
  for (int i = 0; i < 5; i++)
    if (strides[i] < 0) return false;
  v->sb = strides[0]; v->sa = strides[1]; v->sy = strides[2]; v->sx = strides[3]; v->sc = strides[4];
  v->B = batch; v->na = na; v->ny = ny; v->nx = nx; v->no = no;
  return true;
}

extern "C" int ryolo_obj_bce_fwd(const float* x, const long long* strides, int batch, int na, int ny, int nx, int no, int ch,
                                 const float* tobj, float pos_weight, double* sum_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HeadView v;
  RYOLO_ARG_CHECK(x && tobj && sum_out && mk_view(strides, batch, na, ny, nx, no, &v) && ch >= 0 && ch < no);
  const size_t n = (size_t)batch * na * ny * nx;
  obj_bce_fwd_kernel<<<loss_grid(n), 256, 0, stream>>>(x, v, ch, tobj, pos_weight, sum_out);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_obj_bce_bwd(const float* x, const long long* strides, int batch, int na, int ny, int nx, int no, int ch,
                                 const float* tobj, float pos_weight, const float* scale_dev, float* grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HeadView v;
  RYOLO_ARG_CHECK(x && tobj && scale_dev && grad && mk_view(strides, batch, na, ny, nx, no, &v) && ch >= 0 && ch < no);
  const size_t n = (size_t)batch * na * ny * nx;
  obj_bce_bwd_kernel<<<loss_grid(n), 256, 0, stream>>>(x, v, ch, tobj, pos_weight, scale_dev, grad);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_head_grad_nchw_to_padded(const float* g, int batch, int c, int ny, int nx, void* dst, int dst_cstride,
                                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(g && dst && batch > 0 && c > 0 && ny > 0 && nx > 0);
  RYOLO_ARG_CHECK(dst_cstride >= c && dst_cstride % 8 == 0 && batch <= 65535 && ny <= 65535);
  const int cpad = ((c + 7) & ~7) + 8;
  const size_t smem = (size_t)32 * cpad * 2;
  RYOLO_ARG_CHECK(smem <= 160 * 1024);
  RYOLO_SMEM_OPT_IN(head_grad_nchw_kernel, 160 * 1024);
  dim3 grid((unsigned)((nx + 31) / 32), (unsigned)ny, (unsigned)batch);
  head_grad_nchw_kernel<<<grid, 256, smem, stream>>>(g, c, ny, nx, static_cast<__nv_bfloat16*>(dst), dst_cstride);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
