// fp32-FMA direct convolution on CUDA cores: the parity-mode conv (bit-faithful fp32 accumulation,
// only the summation order differs from the oracle) and the path for layers too small/odd for the
// tensor-core kernel (2->2 ConvTranspose flow upsamplers, first layers when running fp32 storage).
// Same argument contract as vps_conv2d_tc; weights are f32 [kh][kw][cin][cout].
#include "common.cuh"

namespace {

struct ConvSimtParams {
  const void* x; int x_n, x_h, x_w, x_cs;
  void* y; int y_h, y_w, y_cs;
  const void* res; int res_cs; int res_after_act;
  const float* w; const float* bias;
  int kh, kw, sh, sw, ph, pw, oh, ow;
  int oy_mul, oy_off, ox_mul, ox_off;
  int cin, cout, act;
  float slope, out_scale;
  int64_t total_pix;
};

constexpr int TP = 64;   // pixels per block
constexpr int TC = 64;   // couts per block
constexpr int KC = 16;   // cin chunk

template <typename TI, typename TO, typename TR>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvSimtParams p) {
  __shared__ float xs[KC][TP + 1];
  __shared__ float ws[KC][TC];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t pix0 = (int64_t)blockIdx.x * TP;
  const int co0 = blockIdx.y * TC;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader assignments: x tile: thread -> (pixel lp = tid % 64, cin rows tid/64 + 4*i)
  const int lp = threadIdx.x & 63, lc = threadIdx.x >> 6;
  const int64_t lpix = pix0 + lp;
  int ln = 0, loy = 0, lox = 0;
  const bool lvalid = lpix < p.total_pix;
  if (lvalid) {
    ln = (int)(lpix / ((int64_t)p.oh * p.ow));
    const int rem = (int)(lpix - (int64_t)ln * p.oh * p.ow);
    loy = rem / p.ow; lox = rem - loy * p.ow;
  }
  const TI* xin = (const TI*)p.x;

  for (int r = 0; r < p.kh; ++r) {
    for (int s = 0; s < p.kw; ++s) {
      const int iy = loy * p.sh - p.ph + r, ix = lox * p.sw - p.pw + s;
      const bool inb = lvalid && iy >= 0 && iy < p.x_h && ix >= 0 && ix < p.x_w;
      const int64_t xoff = inb ? (((int64_t)ln * p.x_h + iy) * p.x_w + ix) * p.x_cs : 0;
      const float* wrs = p.w + (int64_t)(r * p.kw + s) * p.cin * p.cout;
      for (int c0 = 0; c0 < p.cin; c0 += KC) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ci = c0 + lc + 4 * i;
          xs[lc + 4 * i][lp] = (inb && ci < p.cin) ? vps::ldf<TI>(xin + xoff + ci) : 0.f;
        }
        // w tile: thread -> (cout = tid % 64, cin rows tid/64 + 4*i)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ci = c0 + lc + 4 * i, co = co0 + lp;
          ws[lc + 4 * i][lp] = (ci < p.cin && co < p.cout) ? __ldg(wrs + (int64_t)ci * p.cout + co) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          float xv[4], wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = xs[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) wv[j] = ws[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }

  TO* yout = (TO*)p.y;
  const TR* rin = (const TR*)p.res;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pix = pix0 + ty * 4 + i;
    if (pix >= p.total_pix) continue;
    const int n = (int)(pix / ((int64_t)p.oh * p.ow));
    const int rem = (int)(pix - (int64_t)n * p.oh * p.ow);
    const int oy = rem / p.ow, ox = rem - oy * p.ow;
    const int64_t opix = ((int64_t)n * p.y_h + oy * p.oy_mul + p.oy_off) * p.y_w + ox * p.ox_mul + p.ox_off;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + co);
      float rv = 0.f;
      if (rin) rv = vps::ldf<TR>(rin + opix * p.res_cs + co);
      if (rin && !p.res_after_act) v += rv;
      v = vps::apply_act(v, p.act, p.slope) * p.out_scale;
      if (rin && p.res_after_act) v += rv;
      vps::stf<TO>(yout + opix * p.y_cs + co, v);
    }
  }
}

__global__ void pack_weights_simt_kernel(const float* __restrict__ src, const float* __restrict__ scale,
                                         float* __restrict__ dst, int cout, int cin, int kh, int kw,
                                         int transposed) {
  const int64_t total = (int64_t)cout * cin * kh * kw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    int64_t t = i / cout;
    const int ci = (int)(t % cin); t /= cin;
    const int s = (int)(t % kw); t /= kw;
    const int r = (int)t;
    const int64_t si = transposed ? ((((int64_t)ci * cout + co) * kh + r) * kw + s)
                                  : ((((int64_t)co * cin + ci) * kh + r) * kw + s);
    float v = src[si];
    if (scale) v *= scale[co];
    dst[i] = v;
  }
}

template <typename TI, typename TO>
int launch_simt(const vps_conv_args* a, const ConvSimtParams& p, dim3 grid, cudaStream_t st) {
  if (a->res.ptr && a->res.dtype == VPS_BF16)
    conv_simt_kernel<TI, TO, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else
    conv_simt_kernel<TI, TO, float><<<grid, 256, 0, st>>>(p);
  VPS_CUDA_LAST("conv_simt_kernel");
  return VPS_OK;
}

}  // namespace

extern "C" int vps_pack_weights_simt(const float* w, const float* scale, float* dst, int cout, int cin, int kh,
                                     int kw, int transposed, void* stream) {
  const int64_t total = (int64_t)cout * cin * kh * kw;
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  pack_weights_simt_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, scale, dst, cout, cin, kh, kw, transposed);
  VPS_CUDA_LAST("pack_weights_simt");
  return VPS_OK;
}

extern "C" int vps_conv2d_simt(const vps_conv_args* a, void* stream) {
  VPS_CHECK_ARG(a->cin <= a->x.c, "conv2d_simt: cin %d > x.c %d", a->cin, a->x.c);
  VPS_CHECK_ARG((a->oh - 1) * a->oy_mul + a->oy_off < a->y.h && (a->ow - 1) * a->ox_mul + a->ox_off < a->y.w,
                "conv2d_simt: output mapping out of range");
  ConvSimtParams p;
  p.x = a->x.ptr; p.x_n = a->x.n; p.x_h = a->x.h; p.x_w = a->x.w; p.x_cs = a->x.cs;
  p.y = a->y.ptr; p.y_h = a->y.h; p.y_w = a->y.w; p.y_cs = a->y.cs;
  p.res = a->res.ptr; p.res_cs = a->res.cs; p.res_after_act = a->res_after_act;
  p.w = (const float*)a->w; p.bias = a->bias;
  p.kh = a->kh; p.kw = a->kw; p.sh = a->sh; p.sw = a->sw; p.ph = a->ph; p.pw = a->pw; p.oh = a->oh; p.ow = a->ow;
  p.oy_mul = a->oy_mul; p.oy_off = a->oy_off; p.ox_mul = a->ox_mul; p.ox_off = a->ox_off;
  p.cin = a->cin; p.cout = a->cout; p.act = a->act; p.slope = a->slope; p.out_scale = a->out_scale;
  p.total_pix = (int64_t)a->x.n * a->oh * a->ow;
  if (p.total_pix == 0) return VPS_OK;
  dim3 grid((unsigned)((p.total_pix + TP - 1) / TP), (unsigned)((a->cout + TC - 1) / TC));
  cudaStream_t st = (cudaStream_t)stream;
  if (a->x.dtype == VPS_F32 && a->y.dtype == VPS_F32) return launch_simt<float, float>(a, p, grid, st);
  if (a->x.dtype == VPS_BF16 && a->y.dtype == VPS_BF16) return launch_simt<__nv_bfloat16, __nv_bfloat16>(a, p, grid, st);
  if (a->x.dtype == VPS_BF16 && a->y.dtype == VPS_F32) return launch_simt<__nv_bfloat16, float>(a, p, grid, st);
  return launch_simt<float, __nv_bfloat16>(a, p, grid, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Thin-output 3x3 convolution (cout <= 4): FlowNet2 `predict_flow` (C -> 2, submodules.py:30-31).  These layers are
// pure input streaming (2*9*C MAC per pixel): one warp per output pixel, lanes split the channels with 16-byte
// loads, fp32 accumulation, warp-shuffle reduction.  Weights f32 [kh][kw][cin][cout] (the simt packing).
namespace {
template <typename TI, typename TO, int V, int CO>
__global__ void __launch_bounds__(256) conv3x3_thin_kernel(vps::TV<const TI> x, vps::TV<TO> y, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int act, float slope,
                                                           float out_scale, int64_t npix, int G) {
  // G (power of two <= 32) lanes cooperate on one output pixel, 32/G pixels per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % G, slot = lane / G, ppw = 32 / G;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int C = x.c;
  for (int64_t base = warp * ppw; base < npix; base += nwarps * ppw) {
    const int64_t pix = base + slot;
    const bool pv = pix < npix;
    const int ox = pv ? (int)(pix % y.w) : 0;
    const int64_t t = pv ? pix / y.w : 0;
    const int oy = (int)(t % y.h), n = (int)(t / y.h);
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
    if (pv) {
      for (int r = 0; r < 3; ++r) {
        const int iy = oy - 1 + r;
        if (iy < 0 || iy >= x.h) continue;
        for (int s = 0; s < 3; ++s) {
          const int ix = ox - 1 + s;
          if (ix < 0 || ix >= x.w) continue;
          const TI* xp = x.p + x.off(n, iy, ix);
          const float* wp = w + (int64_t)(r * 3 + s) * C * CO;
          for (int c = sub * V; c < C; c += G * V) {
            float v[V];
            if (V > 1 && c + V <= C) {
              vps::ldv<TI, V>(xp + c, v);
            } else {
#pragma unroll
              for (int j = 0; j < V; ++j) v[j] = (c + j < C) ? vps::ldf<TI>(xp + c + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {
              if (c + j < C) {
#pragma unroll
                for (int o = 0; o < CO; ++o) acc[o] = fmaf(v[j], __ldg(wp + (int64_t)(c + j) * CO + o), acc[o]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < CO; ++o)
      for (int sft = G >> 1; sft > 0; sft >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], sft);
    if (pv && sub == 0) {
      TO* yp = y.p + y.off(n, oy, ox);
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        float v = acc[o] + (bias ? bias[o] : 0.f);
        vps::stf<TO>(yp + o, vps::apply_act(v, act, slope) * out_scale);
      }
    }
  }
}
}  // namespace

extern "C" int vps_conv3x3_thin(const vps_conv_args* a, void* stream) {
  VPS_CHECK_ARG(a->kh == 3 && a->kw == 3 && a->sh == 1 && a->sw == 1 && a->ph == 1 && a->pw == 1, "conv3x3_thin: 3x3 s1 p1 only");
  VPS_CHECK_ARG(a->cout == 2 && a->cin == a->x.c && a->res.ptr == nullptr, "conv3x3_thin: cout must be 2, no residual");
  VPS_CHECK_ARG(a->oy_mul == 1 && a->ox_mul == 1 && a->oy_off == 0 && a->ox_off == 0 && a->oh == a->x.h && a->ow == a->x.w &&
                    a->y.h == a->x.h && a->y.w == a->x.w, "conv3x3_thin: identity output mapping only");
  const int64_t npix = (int64_t)a->x.n * a->oh * a->ow;
  if (!npix) return VPS_OK;
  const int Vw = a->x.dtype == VPS_F32 ? 4 : 8;
  const bool vec = a->x.cs % Vw == 0 && ((uintptr_t)a->x.ptr & 15) == 0;
  int G = 1;
  while (G < 32 && G * (vec ? Vw : 1) < a->x.c) G <<= 1;       // lanes per pixel
  int64_t blocks = (npix * G + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  const float* w = (const float*)a->w;
#define THIN(TI, TO, V) conv3x3_thin_kernel<TI, TO, V, 2><<<(int)blocks, 256, 0, st>>>(vps::tv<const TI>(a->x), vps::tv<TO>(a->y), w, a->bias, a->act, a->slope, a->out_scale, npix, G)
  if (a->x.dtype == VPS_F32) {
    if (a->y.dtype == VPS_F32) { if (vec) THIN(float, float, 4); else THIN(float, float, 1); }
    else { if (vec) THIN(float, __nv_bfloat16, 4); else THIN(float, __nv_bfloat16, 1); }
  } else {
    if (a->y.dtype == VPS_F32) { if (vec) THIN(__nv_bfloat16, float, 8); else THIN(__nv_bfloat16, float, 1); }
    else { if (vec) THIN(__nv_bfloat16, __nv_bfloat16, 8); else THIN(__nv_bfloat16, __nv_bfloat16, 1); }
  }
#undef THIN
  VPS_CUDA_LAST("conv3x3_thin");
  return VPS_OK;
}
