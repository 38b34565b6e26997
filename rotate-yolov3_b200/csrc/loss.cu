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

// NCHW fp32 [B, C, ny, nx] -> padded-NHWC bf16 (channel stride dcs), plus the per-channel sums of g (the head convolution's
// bias gradient) on the way.  One block = one image row: chunks of 32 consecutive x, 8 channels per warp and step (eight
// 128-byte coalesced loads in flight per warp), transposed through shared memory with 16-byte stores whose row pitch is an
// odd number of 16-byte units (conflict-free), 16-byte global stores of each pixel's contiguous channels.
__device__ __forceinline__ size_t pad_off(int b, int y, int x, int h, int w, int cs) {
  return (((size_t)b * (h + 2) + y + 1) * (w + 2) + x + 1) * cs;
}
// sum over the 32 lanes of 8 per-lane values in 9 shuffles (halving exchange instead of 8 x 5 butterflies): on return
// lanes 0, 4, .., 28 hold the total of value index lane / 4
__device__ __forceinline__ float warp_sum8(const float* v, int lane) {
  float a[4], b2[2], c1;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float send = h16 ? v[e] : v[e + 4], keep = h16 ? v[e + 4] : v[e];
    a[e] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const float send = h8 ? a[e] : a[e + 2], keep = h8 ? a[e + 2] : a[e];
    b2[e] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float send = h4 ? b2[0] : b2[1], keep = h4 ? b2[1] : b2[0];
    c1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  c1 += __shfl_xor_sync(0xffffffffu, c1, 2);
  c1 += __shfl_xor_sync(0xffffffffu, c1, 1);
  return c1;   // value index = 4 * (lane bit 4) + 2 * (lane bit 3) + (lane bit 2)
}

__global__ void __launch_bounds__(256) head_grad_nchw_kernel(const float* __restrict__ g, int c, int ny, int nx,
                                                             __nv_bfloat16* __restrict__ dst, int dcs,
                                                             float* __restrict__ bias_grad, int rows_per_block) {
  extern __shared__ __align__(16) unsigned char hg_smem[];
  const int c8 = (c + 7) & ~7;
  const int cpad = c8 + ((c8 >> 3) & 1 ? 0 : 8);                      // pitch = odd multiple of 16 bytes
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(hg_smem);   // [32][cpad]
  float* s_bias = reinterpret_cast<float*>(hg_smem + (size_t)32 * cpad * 2);   // [c8]: channel ch is owned by one warp
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < c8; i += 256) s_bias[i] = 0.f;
  __syncthreads();
  const size_t plane = (size_t)ny * nx;
  const int y_end = min(ny, (int)(blockIdx.x + 1) * rows_per_block);
  for (int y = blockIdx.x * rows_per_block; y < y_end; y++) {
    const float* grow = g + ((size_t)b * c * ny + y) * nx;
    for (int x0 = 0; x0 < nx; x0 += 32) {
      const int npx = min(32, nx - x0);
#pragma unroll 1
      for (int ch0 = wid * 8; ch0 < c8; ch0 += 64) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = (lane < npx && ch0 + e < c) ? __ldg(grow + (size_t)(ch0 + e) * plane + x0 + lane) : 0.f;
        __nv_bfloat162 h[4];
#pragma unroll
        for (int e = 0; e < 4; e++) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(tile + lane * cpad + ch0) = *reinterpret_cast<uint4*>(h);
        if (bias_grad) {
          const float t = warp_sum8(v, lane);
          if ((lane & 3) == 0) s_bias[ch0 + (lane >> 2)] += t;
        }
      }
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
      __syncthreads();
    }
  }
  if (bias_grad)
    for (int i = threadIdx.x; i < c; i += 256) atomicAdd(bias_grad + i, s_bias[i]);
}

// The matched (anchor, target) rows of compute_loss: gather of the head values at (b, a, gj, gi), the objectness targets,
// and the scatter of the rows' gradients -- three tiny kernels instead of framework advanced indexing (each index op there
// launches ~6 kernels: bounds reductions, asserts, the op itself).  Rows whose indices fall outside the map (a target on
// the right / bottom border, where the reference raises an IndexError) and rows with mask == 0 read as 0 / write nothing.
struct RowIdx {
  const long long *b, *a, *gj, *gi;
  const unsigned char* mask;
  int rows;
};
__device__ __forceinline__ bool row_cell(const RowIdx& r, const HeadView& v, int i, size_t* off, size_t* cell) {
  if (!r.mask[i]) return false;
  const long long b = r.b[i], a = r.a[i], y = r.gj[i], x = r.gi[i];
  if (b < 0 || b >= v.B || a < 0 || a >= v.na || y < 0 || y >= v.ny || x < 0 || x >= v.nx) return false;
  *off = (size_t)(b * v.sb + a * v.sa + y * v.sy + x * v.sx);
  *cell = (((size_t)b * v.na + a) * v.ny + y) * v.nx + x;
  return true;
}
__global__ void rows_gather_kernel(const float* __restrict__ x, HeadView v, RowIdx r, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= r.rows * v.no) return;
  const int i = t / v.no, k = t - i * v.no;
  size_t off, cell;
  out[t] = row_cell(r, v, i, &off, &cell) ? x[off + (size_t)k * v.sc] : 0.f;
}
__global__ void rows_scatter_add_kernel(float* __restrict__ g, HeadView v, RowIdx r, const float* __restrict__ vals,
                                        const float* __restrict__ scale_dev) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= r.rows * v.no) return;
  const int i = t / v.no, k = t - i * v.no;
  size_t off, cell;
  if (row_cell(r, v, i, &off, &cell)) atomicAdd(g + off + (size_t)k * v.sc, vals[t] * __ldg(scale_dev));
}
__global__ void rows_set_tobj_kernel(float* __restrict__ tobj, HeadView v, RowIdx r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r.rows) return;
  size_t off, cell;
  if (row_cell(r, v, i, &off, &cell)) tobj[cell] = 1.f;
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
                                              float* bias_grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(g && dst && batch > 0 && c > 0 && ny > 0 && nx > 0);
  RYOLO_ARG_CHECK(dst_cstride >= c && dst_cstride % 8 == 0 && batch <= 65535 && c <= 2048);
  const int c8 = (c + 7) & ~7;
  const int cpad = c8 + ((c8 >> 3) & 1 ? 0 : 8);
  const size_t smem = (size_t)32 * cpad * 2 + (size_t)c8 * sizeof(float);
  RYOLO_SMEM_OPT_IN(head_grad_nchw_kernel, 160 * 1024);
  // rows per block: keep ~2k blocks in the grid (one set of c atomics per block for the bias gradient)
  int rpb = (int)(((long long)ny * batch + 2047) / 2048);
  if (rpb < 1) rpb = 1;
  dim3 grid((unsigned)((ny + rpb - 1) / rpb), (unsigned)batch);
  head_grad_nchw_kernel<<<grid, 256, smem, stream>>>(g, c, ny, nx, static_cast<__nv_bfloat16*>(dst), dst_cstride, bias_grad,
                                                      rpb);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

static inline RowIdx mk_rows(const long long* b, const long long* a, const long long* gj, const long long* gi,
                             const unsigned char* mask, int rows) {
  RowIdx r;
  r.b = b; r.a = a; r.gj = gj; r.gi = gi; r.mask = mask; r.rows = rows;
  return r;
}

extern "C" int ryolo_loss_rows_gather(const float* x, const long long* strides, int batch, int na, int ny, int nx, int no,
                                      const long long* b, const long long* a, const long long* gj, const long long* gi,
                                      const unsigned char* mask, int rows, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HeadView v;
  RYOLO_ARG_CHECK(rows >= 0 && mk_view(strides, batch, na, ny, nx, no, &v));
  if (rows == 0) return RYOLO_OK;
  RYOLO_ARG_CHECK(x && b && a && gj && gi && mask && out);
  rows_gather_kernel<<<(rows * no + 255) / 256, 256, 0, stream>>>(x, v, mk_rows(b, a, gj, gi, mask, rows), out);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_loss_rows_scatter_add(float* grad, const long long* strides, int batch, int na, int ny, int nx, int no,
                                           const long long* b, const long long* a, const long long* gj, const long long* gi,
                                           const unsigned char* mask, int rows, const float* vals, const float* scale_dev,
                                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HeadView v;
  RYOLO_ARG_CHECK(rows >= 0 && mk_view(strides, batch, na, ny, nx, no, &v));
  if (rows == 0) return RYOLO_OK;
  RYOLO_ARG_CHECK(grad && b && a && gj && gi && mask && vals && scale_dev);
  rows_scatter_add_kernel<<<(rows * no + 255) / 256, 256, 0, stream>>>(grad, v, mk_rows(b, a, gj, gi, mask, rows), vals,
                                                                        scale_dev);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_loss_rows_set_tobj(float* tobj, int batch, int na, int ny, int nx, const long long* b, const long long* a,
                                        const long long* gj, const long long* gi, const unsigned char* mask, int rows,
                                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(rows >= 0 && batch > 0 && na > 0 && ny > 0 && nx > 0);
  if (rows == 0) return RYOLO_OK;
  RYOLO_ARG_CHECK(tobj && b && a && gj && gi && mask);
  HeadView v;
  v.sb = v.sa = v.sy = v.sx = v.sc = 0;
  v.B = batch; v.na = na; v.ny = ny; v.nx = nx; v.no = 1;
  rows_set_tobj_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(tobj, v, mk_rows(b, a, gj, gi, mask, rows));
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
