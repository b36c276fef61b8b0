// HBM-bound layout / pointwise / resampling kernels (NHWC, f32 or bf16 storage, fp32 math).
// One thread per output element with the channel index fastest => coalesced, 128-bit where the
// channel count allows.  Torch semantics are reproduced exactly where index arithmetic matters
// (align_corners=False bilinear, floor nearest, adaptive pooling windows).
#include "common.cuh"

namespace {

inline int grid_for(int64_t total, int threads = 256) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148 * 32;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

#define GRID_STRIDE(i, total) \
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (total); i += (int64_t)gridDim.x * blockDim.x)

// decompose flat index over [n,h,w,c] with c fastest
#define DECOMP_NHWC(i, T_, n_, y_, x_, c_)      \
  const int c_ = (int)((i) % (T_).c);           \
  int64_t t__ = (i) / (T_).c;                   \
  const int x_ = (int)(t__ % (T_).w);           \
  t__ /= (T_).w;                                \
  const int y_ = (int)(t__ % (T_).h);           \
  const int n_ = (int)(t__ / (T_).h)

template <typename TO>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, vps::TV<TO> dst, int64_t total) {
  GRID_STRIDE(i, total) {
    DECOMP_NHWC(i, dst, n, y, x, c);
    vps::stf<TO>(dst.p + dst.off(n, y, x) + c, src[(((int64_t)n * dst.c + c) * dst.h + y) * dst.w + x]);
  }
}
template <typename TI>
__global__ void nhwc_to_nchw_kernel(vps::TV<const TI> src, float* __restrict__ dst, int64_t total) {
  GRID_STRIDE(i, total) {
    // iterate in NCHW order for coalesced writes
    const int x = (int)(i % src.w);
    int64_t t = i / src.w;
    const int y = (int)(t % src.h); t /= src.h;
    const int c = (int)(t % src.c);
    const int n = (int)(t / src.c);
    dst[i] = vps::ldf<TI>(src.p + src.off(n, y, x) + c);
  }
}

template <typename TI, typename TO>
__global__ void axpby_kernel(vps::TV<const TI> a, vps::TV<const TI> b, int has_b, vps::TV<TO> out, float alpha,
                             float beta, int64_t total) {
  GRID_STRIDE(i, total) {
    DECOMP_NHWC(i, out, n, y, x, c);
    float v = alpha * vps::ldf<TI>(a.p + a.off(n, y, x) + c);
    if (has_b) v += beta * vps::ldf<TI>(b.p + b.off(n, y, x) + c);
    vps::stf<TO>(out.p + out.off(n, y, x) + c, v);
  }
}

template <typename TI, typename TO>
__global__ void resize_bilinear_kernel(vps::TV<const TI> src, vps::TV<TO> out, float sy, float sx, float mul,
                                       int64_t total) {
  GRID_STRIDE(i, total) {
    DECOMP_NHWC(i, out, n, y, x, c);
    // area_pixel_compute_source_index(align_corners=False): max(scale*(dst+0.5)-0.5, 0)
    float fy = fmaxf(sy * ((float)y + 0.5f) - 0.5f, 0.f);
    float fx = fmaxf(sx * ((float)x + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < src.h - 1 ? 1 : 0), x1 = x0 + (x0 < src.w - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float v00 = vps::ldf<TI>(src.p + src.off(n, y0, x0) + c);
    const float v01 = vps::ldf<TI>(src.p + src.off(n, y0, x1) + c);
    const float v10 = vps::ldf<TI>(src.p + src.off(n, y1, x0) + c);
    const float v11 = vps::ldf<TI>(src.p + src.off(n, y1, x1) + c);
    const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
    vps::stf<TO>(out.p + out.off(n, y, x) + c, v * mul);
  }
}

template <typename TI, typename TO>
__global__ void resize_nearest_kernel(vps::TV<const TI> src, vps::TV<TO> out, float sy, float sx, float mul,
                                      int accumulate, int64_t total) {
  GRID_STRIDE(i, total) {
    DECOMP_NHWC(i, out, n, y, x, c);
    const int ys = min((int)floorf((float)y * sy), src.h - 1);
    const int xs = min((int)floorf((float)x * sx), src.w - 1);
    float v = vps::ldf<TI>(src.p + src.off(n, ys, xs) + c) * mul;
    TO* op = out.p + out.off(n, y, x) + c;
    if (accumulate) v += vps::ldf<TO>(op);
    vps::stf<TO>(op, v);
  }
}

template <typename TI, typename TO>
__global__ void pool2d_kernel(vps::TV<const TI> src, vps::TV<TO> out, int k, int s, int p, int is_avg,
                              int64_t total) {
  GRID_STRIDE(i, total) {
    DECOMP_NHWC(i, out, n, y, x, c);
    const int ys = y * s - p, xs = x * s - p;
    float acc = is_avg ? 0.f : -INFINITY;
    for (int r = 0; r < k; ++r) {
      const int yy = ys + r;
      if (yy < 0 || yy >= src.h) continue;
      for (int q = 0; q < k; ++q) {
        const int xx = xs + q;
        if (xx < 0 || xx >= src.w) continue;
        const float v = vps::ldf<TI>(src.p + src.off(n, yy, xx) + c);
        acc = is_avg ? acc + v : fmaxf(acc, v);
      }
    }
    if (is_avg) acc /= (float)(k * k);  // count_include_pad=True (torch default, tcea_modules.py:28)
    vps::stf<TO>(out.p + out.off(n, y, x) + c, acc);
  }
}

// ---- GroupNorm: pass 1 = per-(n,group) sum / sumsq in double via block partials; pass 2 = apply
template <typename TI>
__global__ void gn_stats_kernel(vps::TV<const TI> x, int groups, double* __restrict__ stats) {
  // grid: (chunks, groups, n)
  const int g = blockIdx.y, n = blockIdx.z;
  const int cg = x.c / groups;
  const int64_t npix = (int64_t)x.h * x.w;
  const int64_t total = npix * cg;
  double s = 0.0, ss = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg);
    const int64_t pix = i / cg;
    const float v = vps::ldf<TI>(x.p + ((int64_t)n * npix + pix) * x.cs + g * cg + c);
    s += v; ss += (double)v * v;
  }
  __shared__ double sh[2][32];
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); ss += __shfl_down_sync(0xffffffffu, ss, o); }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = s; sh[1][w] = ss; }
  __syncthreads();
  if (w == 0) {
    s = lane < (blockDim.x >> 5) ? sh[0][lane] : 0.0;
    ss = lane < (blockDim.x >> 5) ? sh[1][lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); ss += __shfl_down_sync(0xffffffffu, ss, o); }
    if (lane == 0) {
      atomicAdd(stats + ((int64_t)n * groups + g) * 2, s);
      atomicAdd(stats + ((int64_t)n * groups + g) * 2 + 1, ss);
    }
  }
}
template <typename TI, typename TO>
__global__ void gn_apply_kernel(vps::TV<const TI> x, vps::TV<TO> y, const double* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int groups,
                                float eps, int relu, int64_t total) {
  const int cg = x.c / groups;
  const double cnt = (double)x.h * x.w * cg;
  GRID_STRIDE(i, total) {
    DECOMP_NHWC(i, y, n, yy, xx, c);
    const int g = c / cg;
    const double m = stats[((int64_t)n * groups + g) * 2] / cnt;
    const double var = stats[((int64_t)n * groups + g) * 2 + 1] / cnt - m * m;
    const float rstd = rsqrtf((float)var + eps);
    float v = (vps::ldf<TI>(x.p + x.off(n, yy, xx) + c) - (float)m) * rstd * gamma[c] + beta[c];
    if (relu) v = fmaxf(v, 0.f);
    vps::stf<TO>(y.p + y.off(n, yy, xx) + c, v);
  }
}

template <typename TI, typename TO>
__global__ void im2col_kernel(vps::TV<const TI> x, vps::TV<TO> cols, int kh, int kw, int sh, int sw, int ph, int pw,
                              int64_t total) {
  const int kk = kh * kw * x.c;
  GRID_STRIDE(i, total) {
    DECOMP_NHWC(i, cols, n, oy, ox, k);
    float v = 0.f;
    if (k < kk) {
      const int ci = k % x.c;
      const int rs = k / x.c;
      const int s = rs % kw, r = rs / kw;
      const int iy = oy * sh - ph + r, ix = ox * sw - pw + s;
      if (iy >= 0 && iy < x.h && ix >= 0 && ix < x.w) v = vps::ldf<TI>(x.p + x.off(n, iy, ix) + ci);
    }
    vps::stf<TO>(cols.p + cols.off(n, oy, ox) + k, v);
  }
}

template <typename TI>
__global__ void sigmoid_flat_kernel_t(vps::TV<const TI> src, float* __restrict__ dst, int64_t total) {
  GRID_STRIDE(i, total) {
    DECOMP_NHWC(i, src, n, y, x, c);
    const float v = vps::ldf<TI>(src.p + src.off(n, y, x) + c);
    dst[i] = 1.f / (1.f + expf(-v));
  }
}

}  // namespace

#define DISPATCH_IO(in_dt, out_dt, TI, TO, ...)                                                  \
  do {                                                                                          \
    if ((in_dt) == VPS_F32 && (out_dt) == VPS_F32) { using TI = float; using TO = float; __VA_ARGS__; }                 \
    else if ((in_dt) == VPS_F32) { using TI = float; using TO = __nv_bfloat16; __VA_ARGS__; }                           \
    else if ((out_dt) == VPS_F32) { using TI = __nv_bfloat16; using TO = float; __VA_ARGS__; }                          \
    else { using TI = __nv_bfloat16; using TO = __nv_bfloat16; __VA_ARGS__; }                                           \
  } while (0)

extern "C" int vps_nchw_to_nhwc(const float* src, const vps_tensor* dst, void* stream) {
  const int64_t total = (int64_t)dst->n * dst->h * dst->w * dst->c;
  if (!total) return VPS_OK;
  VPS_DISPATCH_T(dst->dtype, TO,
                 (nchw_to_nhwc_kernel<TO><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(src, vps::tv<TO>(*dst), total)));
  VPS_CUDA_LAST("nchw_to_nhwc");
  return VPS_OK;
}
extern "C" int vps_nhwc_to_nchw(const vps_tensor* src, float* dst, void* stream) {
  const int64_t total = (int64_t)src->n * src->h * src->w * src->c;
  if (!total) return VPS_OK;
  VPS_DISPATCH_T(src->dtype, TI,
                 (nhwc_to_nchw_kernel<TI><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(vps::tv<const TI>(*src), dst, total)));
  VPS_CUDA_LAST("nhwc_to_nchw");
  return VPS_OK;
}
extern "C" int vps_axpby(const vps_tensor* a, const vps_tensor* b, const vps_tensor* out, float alpha, float beta,
                         void* stream) {
  VPS_CHECK_ARG(a->h == out->h && a->w == out->w && a->c >= out->c && a->n == out->n, "axpby: shape");
  if (b) VPS_CHECK_ARG(b->dtype == a->dtype && b->h == out->h && b->w == out->w && b->c >= out->c, "axpby: b");
  const int64_t total = (int64_t)out->n * out->h * out->w * out->c;
  if (!total) return VPS_OK;
  vps_tensor bb = b ? *b : *a;
  DISPATCH_IO(a->dtype, out->dtype, TI, TO,
              (axpby_kernel<TI, TO><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                  vps::tv<const TI>(*a), vps::tv<const TI>(bb), b != nullptr, vps::tv<TO>(*out), alpha, beta, total)));
  VPS_CUDA_LAST("axpby");
  return VPS_OK;
}
extern "C" int vps_copy_scale(const vps_tensor* src, const vps_tensor* dst, float alpha, void* stream) {
  return vps_axpby(src, nullptr, dst, alpha, 0.f, stream);
}
extern "C" int vps_resize_bilinear(const vps_tensor* src, const vps_tensor* out, float mul, void* stream) {
  VPS_CHECK_ARG(src->c >= out->c && src->n == out->n, "resize_bilinear: shape");
  const int64_t total = (int64_t)out->n * out->h * out->w * out->c;
  if (!total) return VPS_OK;
  const float sy = (float)src->h / (float)out->h, sx = (float)src->w / (float)out->w;
  DISPATCH_IO(src->dtype, out->dtype, TI, TO,
              (resize_bilinear_kernel<TI, TO><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                  vps::tv<const TI>(*src), vps::tv<TO>(*out), sy, sx, mul, total)));
  VPS_CUDA_LAST("resize_bilinear");
  return VPS_OK;
}
extern "C" int vps_resize_nearest(const vps_tensor* src, const vps_tensor* out, float mul, int accumulate,
                                  void* stream) {
  VPS_CHECK_ARG(src->c >= out->c && src->n == out->n, "resize_nearest: shape");
  const int64_t total = (int64_t)out->n * out->h * out->w * out->c;
  if (!total) return VPS_OK;
  const float sy = (float)src->h / (float)out->h, sx = (float)src->w / (float)out->w;
  DISPATCH_IO(src->dtype, out->dtype, TI, TO,
              (resize_nearest_kernel<TI, TO><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                  vps::tv<const TI>(*src), vps::tv<TO>(*out), sy, sx, mul, accumulate, total)));
  VPS_CUDA_LAST("resize_nearest");
  return VPS_OK;
}
extern "C" int vps_pool2d(const vps_tensor* src, const vps_tensor* out, int k, int s, int p, int is_avg,
                          void* stream) {
  VPS_CHECK_ARG(src->c >= out->c && src->n == out->n, "pool2d: shape");
  VPS_CHECK_ARG(out->h == (src->h + 2 * p - k) / s + 1 && out->w == (src->w + 2 * p - k) / s + 1, "pool2d: out size");
  const int64_t total = (int64_t)out->n * out->h * out->w * out->c;
  if (!total) return VPS_OK;
  DISPATCH_IO(src->dtype, out->dtype, TI, TO,
              (pool2d_kernel<TI, TO><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                  vps::tv<const TI>(*src), vps::tv<TO>(*out), k, s, p, is_avg, total)));
  VPS_CUDA_LAST("pool2d");
  return VPS_OK;
}

namespace { double* g_gn_stats = nullptr; int64_t g_gn_cap = 0; }

extern "C" int vps_groupnorm(const vps_tensor* x, const vps_tensor* y, const float* gamma, const float* beta,
                             int groups, float eps, int relu, void* stream) {
  VPS_CHECK_ARG(x->c % groups == 0 && x->c == y->c && x->h == y->h && x->w == y->w, "groupnorm: shape");
  const int64_t total = (int64_t)x->n * x->h * x->w * x->c;
  if (!total) return VPS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t need = (int64_t)x->n * groups * 2;
  if (g_gn_cap < need) {
    if (g_gn_stats) cudaFree(g_gn_stats);
    if (cudaMalloc(&g_gn_stats, need * sizeof(double) * 4) != cudaSuccess) { vps::set_error("groupnorm: malloc"); return VPS_E_CUDA; }
    g_gn_cap = need * 4;
  }
  cudaMemsetAsync(g_gn_stats, 0, need * sizeof(double), st);
  const int64_t per_group = (int64_t)x->h * x->w * (x->c / groups);
  int chunks = (int)((per_group + 256 * 32 - 1) / (256 * 32));
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, groups, x->n);
  VPS_DISPATCH_T(x->dtype, TI, (gn_stats_kernel<TI><<<grid, 256, 0, st>>>(vps::tv<const TI>(*x), groups, g_gn_stats)));
  VPS_CUDA_LAST("gn_stats");
  DISPATCH_IO(x->dtype, y->dtype, TI, TO,
              (gn_apply_kernel<TI, TO><<<grid_for(total), 256, 0, st>>>(vps::tv<const TI>(*x), vps::tv<TO>(*y), g_gn_stats,
                                                                       gamma, beta, groups, eps, relu, total)));
  VPS_CUDA_LAST("gn_apply");
  return VPS_OK;
}

extern "C" int vps_im2col(const vps_tensor* x, const vps_tensor* cols, int kh, int kw, int sh, int sw, int ph, int pw,
                          void* stream) {
  VPS_CHECK_ARG(cols->c >= kh * kw * x->c, "im2col: cols.c %d < %d", cols->c, kh * kw * x->c);
  const int64_t total = (int64_t)cols->n * cols->h * cols->w * cols->c;
  if (!total) return VPS_OK;
  DISPATCH_IO(x->dtype, cols->dtype, TI, TO,
              (im2col_kernel<TI, TO><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                  vps::tv<const TI>(*x), vps::tv<TO>(*cols), kh, kw, sh, sw, ph, pw, total)));
  VPS_CUDA_LAST("im2col");
  return VPS_OK;
}

// RPN objectness: sigmoid of an NHWC score map flattened to the reference's (h, w, anchor) order
// (rpn_head.py:69-72).
extern "C" int vps_sigmoid_flat(const vps_tensor* t, float* dst, void* stream) {
  const int64_t total = (int64_t)t->n * t->h * t->w * t->c;
  if (!total) return VPS_OK;
  VPS_DISPATCH_T(t->dtype, TI,
                 (sigmoid_flat_kernel_t<TI><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(vps::tv<const TI>(*t), dst, total)));
  VPS_CUDA_LAST("sigmoid_flat");
  return VPS_OK;
}
