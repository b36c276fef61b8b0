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

// All kernels below use the (x*chunks, y, n) pixel grid of vps::pix_grid: no 64-bit index division, and V
// consecutive channels (16 bytes) per thread when every tensor involved allows it (V = 1 otherwise).
template <typename TI, typename TO, int V>
__global__ void axpby_kernel(vps::TV<const TI> a, vps::TV<const TI> b, int has_b, vps::TV<TO> out, float alpha, float beta) {
  VPS_PIX_COORDS(out, V, c, x, y, n);
  float va[V], vb[V];
  vps::ldv<TI, V>(a.p + a.off(n, y, x) + c, va);
  if (has_b) vps::ldv<TI, V>(b.p + b.off(n, y, x) + c, vb);
#pragma unroll
  for (int j = 0; j < V; ++j) va[j] = alpha * va[j] + (has_b ? beta * vb[j] : 0.f);
  vps::stv<TO, V>(out.p + out.off(n, y, x) + c, va);
}

template <typename TI, typename TO, int V>
__global__ void resize_bilinear_kernel(vps::TV<const TI> src, vps::TV<TO> out, float sy, float sx, float mul) {
  VPS_PIX_COORDS(out, V, c, x, y, n);
  // area_pixel_compute_source_index(align_corners=False): max(scale*(dst+0.5)-0.5, 0)
  const float fy = fmaxf(sy * ((float)y + 0.5f) - 0.5f, 0.f);
  const float fx = fmaxf(sx * ((float)x + 0.5f) - 0.5f, 0.f);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < src.h - 1 ? 1 : 0), x1 = x0 + (x0 < src.w - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  float v00[V], v01[V], v10[V], v11[V];
  vps::ldv<TI, V>(src.p + src.off(n, y0, x0) + c, v00);
  vps::ldv<TI, V>(src.p + src.off(n, y0, x1) + c, v01);
  vps::ldv<TI, V>(src.p + src.off(n, y1, x0) + c, v10);
  vps::ldv<TI, V>(src.p + src.off(n, y1, x1) + c, v11);
#pragma unroll
  for (int j = 0; j < V; ++j) v00[j] = vps_bilerp(v00[j], v01[j], v10[j], v11[j], hx, lx, hy, ly) * mul;
  vps::stv<TO, V>(out.p + out.off(n, y, x) + c, v00);
}

template <typename TI, typename TO, int V>
__global__ void resize_nearest_kernel(vps::TV<const TI> src, vps::TV<TO> out, float sy, float sx, float mul, int accumulate) {
  VPS_PIX_COORDS(out, V, c, x, y, n);
  const int ys = min((int)floorf((float)y * sy), src.h - 1);
  const int xs = min((int)floorf((float)x * sx), src.w - 1);
  float v[V], o[V];
  vps::ldv<TI, V>(src.p + src.off(n, ys, xs) + c, v);
  TO* op = out.p + out.off(n, y, x) + c;
  if (accumulate) vps::ldv<TO, V>(op, o);
#pragma unroll
  for (int j = 0; j < V; ++j) v[j] = v[j] * mul + (accumulate ? o[j] : 0.f);
  vps::stv<TO, V>(op, v);
}

template <typename TI, typename TO, int V>
__global__ void pool2d_kernel(vps::TV<const TI> src, vps::TV<TO> out, int k, int s, int p, int is_avg) {
  VPS_PIX_COORDS(out, V, c, x, y, n);
  const int ys = y * s - p, xs = x * s - p;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = is_avg ? 0.f : -INFINITY;
  for (int r = 0; r < k; ++r) {
    const int yy = ys + r;
    if (yy < 0 || yy >= src.h) continue;
    for (int q = 0; q < k; ++q) {
      const int xx = xs + q;
      if (xx < 0 || xx >= src.w) continue;
      float v[V];
      vps::ldv<TI, V>(src.p + src.off(n, yy, xx) + c, v);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = is_avg ? acc[j] + v[j] : fmaxf(acc[j], v[j]);
    }
  }
  if (is_avg) {
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] /= (float)(k * k);   // count_include_pad=True (torch default, tcea_modules.py:28)
  }
  vps::stv<TO, V>(out.p + out.off(n, y, x) + c, acc);
}

// space-to-depth (block 2): y[n, Y, X, (dy*2+dx)*C + c] = x[n, 2Y+dy, 2X+dx, c]; out-of-range reads (odd sizes) are 0.
// Turns a stride-2 convolution over C channels into a stride-1 convolution over 4C channels (first layers).
template <typename TI, typename TO>
__global__ void space_to_depth2_kernel(vps::TV<const TI> x, vps::TV<TO> y) {
  VPS_PIX_COORDS(y, 1, k, X, Y, n);
  const int C = x.c;
  const int c = k % C, q = k / C;
  const int iy = 2 * Y + (q >> 1), ix = 2 * X + (q & 1);
  float v = 0.f;
  if (q < 4 && iy < x.h && ix < x.w) v = vps::ldf<TI>(x.p + x.off(n, iy, ix) + c);
  vps::stf<TO>(y.p + y.off(n, Y, X) + k, v);
}

// bf16 fast path: one thread assembles 8 consecutive output channels (scalar, L1-resident reads of the four source
// pixels) and writes them with one 16-byte store; the channel padding up to the pixel stride is written as zeros.
__global__ void space_to_depth2_bf16x8_kernel(vps::TV<const __nv_bfloat16> x, vps::TV<__nv_bfloat16> y) {
  const int chunks = y.cs / 8;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= y.w * chunks) return;
  const int k0 = (t % chunks) * 8, X = t / chunks, Y = blockIdx.y, n = blockIdx.z;
  const int C = x.c;
  int c = k0 % C, q = k0 / C;
  uint32_t pk[4];
#pragma unroll
  for (int j2 = 0; j2 < 4; ++j2) {
    unsigned short v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int iy = 2 * Y + (q >> 1), ix = 2 * X + (q & 1);
      v[e] = 0;
      if (q < 4 && iy < x.h && ix < x.w) v[e] = *reinterpret_cast<const unsigned short*>(x.p + x.off(n, iy, ix) + c);
      if (++c == C) { c = 0; ++q; }
    }
    pk[j2] = (uint32_t)v[0] | ((uint32_t)v[1] << 16);
  }
  *reinterpret_cast<uint4*>(y.p + y.off(n, Y, X) + k0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// ---- GroupNorm: pass 1 = per-(n,group) sum / sumsq in double via block partials; pass 2 = apply
template <typename TI>
__global__ void gn_stats_kernel(vps::TV<const TI> x, int groups, double* __restrict__ stats) {
  // scalar fallback. grid: (chunks, groups, n)
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

// vector path: a thread owns one 16-byte channel chunk (V channels, <= 2 groups) and strides over pixels, so a warp
// reads whole pixels (fully coalesced); per-chunk partials are combined across the block's pixel rows in shared memory.
// block = (C/V chunks) x (256*V/C pixel rows); grid: (pixel slices, 1, n)
template <typename TI, int V>
__global__ void __launch_bounds__(256) gn_stats_vec_kernel(vps::TV<const TI> x, int groups, double* __restrict__ stats) {
  const int chunks = x.c / V;
  const int rows = blockDim.x / chunks;
  const int ch = threadIdx.x % chunks, row = threadIdx.x / chunks;
  const int n = blockIdx.z;
  const int cg = x.c / groups;
  const int c0 = ch * V;
  const int64_t npix = (int64_t)x.h * x.w;
  float s[2] = {0.f, 0.f}, ss[2] = {0.f, 0.f};       // fp32 partials over <= a few thousand pixels, combined in double
  double ds[2] = {0.0, 0.0}, dss[2] = {0.0, 0.0};
  int cnt = 0;
  if (row < rows) {
    // 4 independent 16-byte loads in flight per thread (one was latency bound: 1.7 TB/s on an L2/HBM-resident map)
    constexpr int U = 4;
    const int64_t stride = (int64_t)gridDim.x * rows;
    for (int64_t pix = (int64_t)blockIdx.x * rows + row; pix < npix; pix += U * stride) {
      float v[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t q = pix + u * stride;
        if (q < npix) vps::ldv<TI, V>(x.p + ((int64_t)n * npix + q) * x.cs + c0, v[u]);
        else {
#pragma unroll
          for (int j = 0; j < V; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int gi = (V > 1 && j >= cg) ? 1 : 0;          // cg < V only when V == 2 * cg (no runtime division)
          s[gi] += v[u][j];
          ss[gi] += v[u][j] * v[u][j];
        }
      }
      if (++cnt == 16) {
        ds[0] += s[0]; ds[1] += s[1]; dss[0] += ss[0]; dss[1] += ss[1];
        s[0] = s[1] = ss[0] = ss[1] = 0.f; cnt = 0;
      }
    }
    ds[0] += s[0]; ds[1] += s[1]; dss[0] += ss[0]; dss[1] += ss[1];
  }
  __shared__ double sh[256][4];
  sh[threadIdx.x][0] = ds[0]; sh[threadIdx.x][1] = dss[0]; sh[threadIdx.x][2] = ds[1]; sh[threadIdx.x][3] = dss[1];
  __syncthreads();
  if (row == 0) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int r = 0; r < rows; ++r) {
      const double* p = sh[r * chunks + ch];
      a0 += p[0]; a1 += p[1]; a2 += p[2]; a3 += p[3];
    }
    const int g0 = c0 / cg;
    atomicAdd(stats + ((int64_t)n * groups + g0) * 2, a0);
    atomicAdd(stats + ((int64_t)n * groups + g0) * 2 + 1, a1);
    if (cg < V) {
      atomicAdd(stats + ((int64_t)n * groups + g0 + 1) * 2, a2);
      atomicAdd(stats + ((int64_t)n * groups + g0 + 1) * 2 + 1, a3);
    }
  }
}

constexpr int GN_MAX_C = 1024;
// mean / rstd of every (n, group) are finalised ONCE per block into shared memory (fp64 divides per element made this
// kernel ALU-bound); the per-element expression (v - mean) * rstd * gamma + beta is unchanged.
template <typename TI, typename TO, int V>
__global__ void gn_apply_kernel(vps::TV<const TI> x, vps::TV<TO> y, const double* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int groups, float eps,
                                int relu) {
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float s_gamma[GN_MAX_C], s_beta[GN_MAX_C];      // per-element global loads of gamma/beta made this LSU-bound
  const int cg = x.c / groups;
  for (int c = threadIdx.x; c < x.c; c += blockDim.x) { s_gamma[c] = gamma[c]; s_beta[c] = beta[c]; }
  {
    const int n_blk = blockIdx.z;                    // pix_grid: z = image index
    const double cnt = (double)x.h * x.w * cg;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      const double m = stats[((int64_t)n_blk * groups + g) * 2] / cnt;
      const double var = stats[((int64_t)n_blk * groups + g) * 2 + 1] / cnt - m * m;
      s_mean[g] = (float)m;
      s_rstd[g] = rsqrtf((float)var + eps);
    }
  }
  __syncthreads();
  VPS_PIX_COORDS(y, V, c, xx, yy, n);
  float v[V];
  vps::ldv<TI, V>(x.p + x.off(n, yy, xx) + c, v);
  const int g0 = c / cg;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int g = (V > 1 && cg < V) ? g0 + (j >= cg ? 1 : 0) : (cg % V == 0 ? g0 : (c + j) / cg);
    float o = (v[j] - s_mean[g]) * s_rstd[g] * s_gamma[c + j] + s_beta[c + j];
    v[j] = relu ? fmaxf(o, 0.f) : o;
  }
  vps::stv<TO, V>(y.p + y.off(n, yy, xx) + c, v);
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

// bf16 fast path: one thread = one output pixel x 8 consecutive k (one 16-byte store); (r,s,ci) is decomposed once
// with 32-bit arithmetic and then incremented.
__global__ void im2col_bf16x8_kernel(vps::TV<const __nv_bfloat16> x, vps::TV<__nv_bfloat16> cols, int kh, int kw, int sh,
                                     int sw, int ph, int pw, int chunks, int64_t total) {
  const int kk = kh * kw * x.c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    const int64_t pix = i / chunks;
    const int ox = (int)(pix % cols.w);
    const int64_t t = pix / cols.w;
    const int oy = (int)(t % cols.h), n = (int)(t / cols.h);
    int k = ch * 8;
    int ci = k % x.c;
    int rs = k / x.c;
    int s = rs % kw, r = rs / kw;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f = 0.f;
      if (k < kk) {
        const int iy = oy * sh - ph + r, ix = ox * sw - pw + s;
        if (iy >= 0 && iy < x.h && ix >= 0 && ix < x.w) f = __bfloat162float(x.p[x.off(n, iy, ix) + ci]);
      }
      v[j] = __float2bfloat16_rn(f);
      ++k;
      if (++ci == x.c) { ci = 0; if (++s == kw) { s = 0; ++r; } }
    }
    *(uint4*)(cols.p + pix * cols.cs + ch * 8) = *(const uint4*)v;
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
  if (!((int64_t)out->n * out->h * out->w * out->c)) return VPS_OK;
  const vps_tensor bb = b ? *b : *a;
  const bool vc = vps::vec_ok(*a, out->c) && vps::vec_ok(bb, out->c) && vps::vec_ok(*out, out->c);
  const bool vec = a->dtype == out->dtype && vc;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && out->dtype == VPS_F32)
    axpby_kernel<float, float, 4><<<vps::pix_grid(out->w, out->c / 4, out->h, out->n), 256, 0, st>>>(
        vps::tv<const float>(*a), vps::tv<const float>(bb), b != nullptr, vps::tv<float>(*out), alpha, beta);
  else if (vec)
    axpby_kernel<__nv_bfloat16, __nv_bfloat16, 8><<<vps::pix_grid(out->w, out->c / 8, out->h, out->n), 256, 0, st>>>(
        vps::tv<const __nv_bfloat16>(*a), vps::tv<const __nv_bfloat16>(bb), b != nullptr, vps::tv<__nv_bfloat16>(*out), alpha, beta);
  else
    DISPATCH_IO(a->dtype, out->dtype, TI, TO,
                (axpby_kernel<TI, TO, 1><<<vps::pix_grid(out->w, out->c, out->h, out->n), 256, 0, st>>>(
                    vps::tv<const TI>(*a), vps::tv<const TI>(bb), b != nullptr, vps::tv<TO>(*out), alpha, beta)));
  VPS_CUDA_LAST("axpby");
  return VPS_OK;
}
extern "C" int vps_copy_scale(const vps_tensor* src, const vps_tensor* dst, float alpha, void* stream) {
  return vps_axpby(src, nullptr, dst, alpha, 0.f, stream);
}
extern "C" int vps_resize_bilinear(const vps_tensor* src, const vps_tensor* out, float mul, void* stream) {
  VPS_CHECK_ARG(src->c >= out->c && src->n == out->n, "resize_bilinear: shape");
  if (!((int64_t)out->n * out->h * out->w * out->c)) return VPS_OK;
  const float sy = (float)src->h / (float)out->h, sx = (float)src->w / (float)out->w;
  const bool vec = src->dtype == out->dtype && vps::vec_ok(*src, out->c) && vps::vec_ok(*out, out->c);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && out->dtype == VPS_F32)
    resize_bilinear_kernel<float, float, 4><<<vps::pix_grid(out->w, out->c / 4, out->h, out->n), 256, 0, st>>>(
        vps::tv<const float>(*src), vps::tv<float>(*out), sy, sx, mul);
  else if (vec)
    resize_bilinear_kernel<__nv_bfloat16, __nv_bfloat16, 8><<<vps::pix_grid(out->w, out->c / 8, out->h, out->n), 256, 0, st>>>(
        vps::tv<const __nv_bfloat16>(*src), vps::tv<__nv_bfloat16>(*out), sy, sx, mul);
  else
    DISPATCH_IO(src->dtype, out->dtype, TI, TO,
                (resize_bilinear_kernel<TI, TO, 1><<<vps::pix_grid(out->w, out->c, out->h, out->n), 256, 0, st>>>(
                    vps::tv<const TI>(*src), vps::tv<TO>(*out), sy, sx, mul)));
  VPS_CUDA_LAST("resize_bilinear");
  return VPS_OK;
}
extern "C" int vps_resize_nearest(const vps_tensor* src, const vps_tensor* out, float mul, int accumulate,
                                  void* stream) {
  VPS_CHECK_ARG(src->c >= out->c && src->n == out->n, "resize_nearest: shape");
  if (!((int64_t)out->n * out->h * out->w * out->c)) return VPS_OK;
  const float sy = (float)src->h / (float)out->h, sx = (float)src->w / (float)out->w;
  const bool vec = src->dtype == out->dtype && vps::vec_ok(*src, out->c) && vps::vec_ok(*out, out->c);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && out->dtype == VPS_F32)
    resize_nearest_kernel<float, float, 4><<<vps::pix_grid(out->w, out->c / 4, out->h, out->n), 256, 0, st>>>(
        vps::tv<const float>(*src), vps::tv<float>(*out), sy, sx, mul, accumulate);
  else if (vec)
    resize_nearest_kernel<__nv_bfloat16, __nv_bfloat16, 8><<<vps::pix_grid(out->w, out->c / 8, out->h, out->n), 256, 0, st>>>(
        vps::tv<const __nv_bfloat16>(*src), vps::tv<__nv_bfloat16>(*out), sy, sx, mul, accumulate);
  else
    DISPATCH_IO(src->dtype, out->dtype, TI, TO,
                (resize_nearest_kernel<TI, TO, 1><<<vps::pix_grid(out->w, out->c, out->h, out->n), 256, 0, st>>>(
                    vps::tv<const TI>(*src), vps::tv<TO>(*out), sy, sx, mul, accumulate)));
  VPS_CUDA_LAST("resize_nearest");
  return VPS_OK;
}
extern "C" int vps_pool2d(const vps_tensor* src, const vps_tensor* out, int k, int s, int p, int is_avg,
                          void* stream) {
  VPS_CHECK_ARG(src->c >= out->c && src->n == out->n, "pool2d: shape");
  VPS_CHECK_ARG(out->h == (src->h + 2 * p - k) / s + 1 && out->w == (src->w + 2 * p - k) / s + 1, "pool2d: out size");
  if (!((int64_t)out->n * out->h * out->w * out->c)) return VPS_OK;
  const bool vec = src->dtype == out->dtype && vps::vec_ok(*src, out->c) && vps::vec_ok(*out, out->c);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && out->dtype == VPS_F32)
    pool2d_kernel<float, float, 4><<<vps::pix_grid(out->w, out->c / 4, out->h, out->n), 256, 0, st>>>(
        vps::tv<const float>(*src), vps::tv<float>(*out), k, s, p, is_avg);
  else if (vec)
    pool2d_kernel<__nv_bfloat16, __nv_bfloat16, 8><<<vps::pix_grid(out->w, out->c / 8, out->h, out->n), 256, 0, st>>>(
        vps::tv<const __nv_bfloat16>(*src), vps::tv<__nv_bfloat16>(*out), k, s, p, is_avg);
  else
    DISPATCH_IO(src->dtype, out->dtype, TI, TO,
                (pool2d_kernel<TI, TO, 1><<<vps::pix_grid(out->w, out->c, out->h, out->n), 256, 0, st>>>(
                    vps::tv<const TI>(*src), vps::tv<TO>(*out), k, s, p, is_avg)));
  VPS_CUDA_LAST("pool2d");
  return VPS_OK;
}

// statistics scratch: a ring of slots, one per call, so GroupNorm calls on parallel streams / graph branches never share one
namespace { double* g_gn_ring = nullptr; unsigned g_gn_next = 0; constexpr int GN_SLOTS = 16; constexpr int GN_SLOT_DOUBLES = 4096; }

extern "C" int vps_groupnorm(const vps_tensor* x, const vps_tensor* y, const float* gamma, const float* beta,
                             int groups, float eps, int relu, void* stream) {
  VPS_CHECK_ARG(x->c % groups == 0 && x->c == y->c && x->h == y->h && x->w == y->w, "groupnorm: shape");
  VPS_CHECK_ARG(groups >= 1 && groups <= 64 && x->c <= GN_MAX_C, "groupnorm: groups %d not in [1, 64] or c %d > %d", groups, x->c, GN_MAX_C);
  const int64_t total = (int64_t)x->n * x->h * x->w * x->c;
  if (!total) return VPS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t need = (int64_t)x->n * groups * 2;
  VPS_CHECK_ARG(need <= GN_SLOT_DOUBLES, "groupnorm: n * groups too large (%lld)", (long long)need);
  if (!g_gn_ring && cudaMalloc(&g_gn_ring, sizeof(double) * GN_SLOTS * GN_SLOT_DOUBLES) != cudaSuccess) {
    vps::set_error("groupnorm: malloc");
    return VPS_E_CUDA;
  }
  double* g_gn_stats = g_gn_ring + (size_t)(g_gn_next++ % GN_SLOTS) * GN_SLOT_DOUBLES;
  cudaMemsetAsync(g_gn_stats, 0, need * sizeof(double), st);
  const int64_t per_group = (int64_t)x->h * x->w * (x->c / groups);
  int chunks = (int)((per_group + 256 * 32 - 1) / (256 * 32));
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  const int V = x->dtype == VPS_F32 ? 4 : 8;
  const int cg = x->c / groups;
  const bool vstats = vps::vec_ok(*x, x->c) && (x->c / V) <= 256 && 256 % (x->c / V) == 0 && (cg >= V ? cg % V == 0 : V == 2 * cg);
  if (vstats) {
    const int rows = 256 / (x->c / V);
    int64_t slices = ((int64_t)x->h * x->w + rows * 32 - 1) / (rows * 32);
    if (slices > 148 * 4) slices = 148 * 4;
    if (slices < 1) slices = 1;
    dim3 grid((unsigned)slices, 1, x->n);
    if (x->dtype == VPS_F32) gn_stats_vec_kernel<float, 4><<<grid, 256, 0, st>>>(vps::tv<const float>(*x), groups, g_gn_stats);
    else gn_stats_vec_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>(vps::tv<const __nv_bfloat16>(*x), groups, g_gn_stats);
  } else {
    dim3 grid(chunks, groups, x->n);
    VPS_DISPATCH_T(x->dtype, TI, (gn_stats_kernel<TI><<<grid, 256, 0, st>>>(vps::tv<const TI>(*x), groups, g_gn_stats)));
  }
  VPS_CUDA_LAST("gn_stats");
  const bool vec = x->dtype == y->dtype && vps::vec_ok(*x, x->c) && vps::vec_ok(*y, x->c);
  if (vec && y->dtype == VPS_F32)
    gn_apply_kernel<float, float, 4><<<vps::pix_grid(y->w, y->c / 4, y->h, y->n), 256, 0, st>>>(
        vps::tv<const float>(*x), vps::tv<float>(*y), g_gn_stats, gamma, beta, groups, eps, relu);
  else if (vec)
    gn_apply_kernel<__nv_bfloat16, __nv_bfloat16, 8><<<vps::pix_grid(y->w, y->c / 8, y->h, y->n), 256, 0, st>>>(
        vps::tv<const __nv_bfloat16>(*x), vps::tv<__nv_bfloat16>(*y), g_gn_stats, gamma, beta, groups, eps, relu);
  else
    DISPATCH_IO(x->dtype, y->dtype, TI, TO,
                (gn_apply_kernel<TI, TO, 1><<<vps::pix_grid(y->w, y->c, y->h, y->n), 256, 0, st>>>(
                    vps::tv<const TI>(*x), vps::tv<TO>(*y), g_gn_stats, gamma, beta, groups, eps, relu)));
  VPS_CUDA_LAST("gn_apply");
  return VPS_OK;
}

extern "C" int vps_im2col(const vps_tensor* x, const vps_tensor* cols, int kh, int kw, int sh, int sw, int ph, int pw,
                          void* stream) {
  VPS_CHECK_ARG(cols->c >= kh * kw * x->c, "im2col: cols.c %d < %d", cols->c, kh * kw * x->c);
  const int64_t total = (int64_t)cols->n * cols->h * cols->w * cols->c;
  if (!total) return VPS_OK;
  if (x->dtype == VPS_BF16 && cols->dtype == VPS_BF16 && cols->c % 8 == 0 && cols->cs % 8 == 0 &&
      ((uintptr_t)cols->ptr & 15) == 0) {
    const int chunks = cols->c / 8;
    const int64_t tot8 = total / 8;
    int64_t blocks = (tot8 + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    im2col_bf16x8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(vps::tv<const __nv_bfloat16>(*x),
                                                                        vps::tv<__nv_bfloat16>(*cols), kh, kw, sh, sw, ph,
                                                                        pw, chunks, tot8);
    VPS_CUDA_LAST("im2col_bf16x8");
    return VPS_OK;
  }
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

extern "C" int vps_space_to_depth2(const vps_tensor* x, const vps_tensor* y, void* stream) {
  VPS_CHECK_ARG(y->c == 4 * x->c && y->h == (x->h + 1) / 2 && y->w == (x->w + 1) / 2 && y->n == x->n, "space_to_depth2: shapes");
  if (!((int64_t)y->n * y->h * y->w * y->c)) return VPS_OK;
  if (x->dtype == VPS_BF16 && y->dtype == VPS_BF16 && y->cs % 8 == 0 && ((uintptr_t)y->ptr & 15) == 0 && y->cs >= y->c) {
    space_to_depth2_bf16x8_kernel<<<vps::pix_grid(y->w, y->cs / 8, y->h, y->n), 256, 0, (cudaStream_t)stream>>>(
        vps::tv<const __nv_bfloat16>(*x), vps::tv<__nv_bfloat16>(*y));
    VPS_CUDA_LAST("space_to_depth2");
    return VPS_OK;
  }
  DISPATCH_IO(x->dtype, y->dtype, TI, TO,
              (space_to_depth2_kernel<TI, TO><<<vps::pix_grid(y->w, y->c, y->h, y->n), 256, 0, (cudaStream_t)stream>>>(
                  vps::tv<const TI>(*x), vps::tv<TO>(*y))));
  VPS_CUDA_LAST("space_to_depth2");
  return VPS_OK;
}

// ---------------------------------------------------------------- input stage: Normalize + Pad + to-tensor
// What the reference's test pipeline does on the host between LoadImageFromFile and the model for img and ref_img
// (mmdet/datasets/pipelines/transforms.py:295-318 Normalize -> mmcv.imnormalize: float32(img), BGR->RGB, (img - mean) / std;
// :238-270 Pad(size_divisor=32) -> zero pad bottom / right; formating.py:46-68 ImageToTensor: HWC -> CHW): one pass from the
// uint8 HWC BGR frame to the fp32 NCHW padded tensor simple_test takes.  (x - mean) / std with IEEE fp32 subtract and divide:
// bit-identical to numpy's float32 arithmetic.  Uploading uint8 frames cuts the host->device bytes of a pair from 50 MB to 12.6 MB.
namespace {
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ src, int h, int w, float m0, float m1, float m2, float s0, float s1,
                                     float s2, int to_rgb, float* __restrict__ out, int hp, int wp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= wp) return;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;                     // pad_val = 0 (after normalisation, as in the reference: Pad follows Normalize)
  if (y < h && x < w) {
    const uint8_t* p = src + ((int64_t)y * w + x) * 3;
    const float b = (float)p[0], g = (float)p[1], r = (float)p[2];
    const float c0 = to_rgb ? r : b, c2 = to_rgb ? b : r;
    v0 = __fdiv_rn(__fsub_rn(c0, m0), s0);
    v1 = __fdiv_rn(__fsub_rn(g, m1), s1);
    v2 = __fdiv_rn(__fsub_rn(c2, m2), s2);
  }
  const int64_t plane = (int64_t)hp * wp, o = (int64_t)y * wp + x;
  out[o] = v0; out[plane + o] = v1; out[2 * plane + o] = v2;
}
}  // namespace

extern "C" int vps_preprocess_u8(const uint8_t* bgr_hwc, int h, int w, const float* mean3, const float* std3, int to_rgb,
                                 float* out_nchw, int hp, int wp, void* stream) {
  VPS_CHECK_ARG(h > 0 && w > 0 && hp >= h && wp >= w, "preprocess_u8: shapes %dx%d -> %dx%d", h, w, hp, wp);
  dim3 grid((unsigned)((wp + 255) / 256), (unsigned)hp);
  preprocess_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bgr_hwc, h, w, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2],
                                                              to_rgb, out_nchw, hp, wp);
  VPS_CUDA_LAST("preprocess_u8");
  return VPS_OK;
}

// ---------------------------------------------------------------- 3x3 convolutions with <= 3 output channels (predict_flow)
// FlowNet2's predict_flow layers (submodules.py:27-28: Conv2d(cin, 2, 3, 1, 1), 19 launches per pair) have K = 9*cin up to
// 9234 but N = 2: as implicit GEMMs they cost one K step per (tap, 32 channels) at any N.  They run as a 1x1 tensor-core
// convolution with the taps moved to the output-channel axis, z[p][t*cout+co] = sum_c x[p][c] * w[co][c][t] (9x fewer K
// steps, the input is read once instead of once per tap), followed by this gather:
//     out[n,y,x,co] = act(bias[co] + sum_t z[n, y + t/3 - 1, x + t%3 - 1, t*cout + co]) * out_scale    (zero outside the map)
namespace {
template <int CO>
__global__ void tap_gather3x3_kernel(vps::TV<const float> z, vps::TV<float> out, const float* __restrict__ bias, int act, float slope,
                                     float out_scale) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (x >= out.w) return;
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    if (yy < 0 || yy >= z.h || xx < 0 || xx >= z.w) continue;
    const float* zp = z.p + z.off(n, yy, xx) + t * CO;
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] += __ldg(zp + c);
  }
  float* op = out.p + out.off(n, y, x);
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    float v = acc[c];
    if (bias) v += __ldg(bias + c);
    op[c] = vps::apply_act(v, act, slope) * out_scale;
  }
}
}  // namespace

extern "C" int vps_tap_gather3x3(const vps_tensor* z, const vps_tensor* out, const float* bias, int act, float slope,
                                 float out_scale, void* stream) {
  VPS_CHECK_ARG(z->dtype == VPS_F32 && out->dtype == VPS_F32, "tap_gather3x3: fp32 tensors");
  VPS_CHECK_ARG(out->c >= 1 && out->c <= 3 && z->c == 9 * out->c, "tap_gather3x3: z.c %d != 9 * out.c %d", z->c, out->c);
  VPS_CHECK_ARG(z->n == out->n && z->h == out->h && z->w == out->w, "tap_gather3x3: shapes");
  if (!((int64_t)out->n * out->h * out->w)) return VPS_OK;
  dim3 grid((unsigned)vps::cdiv(out->w, 128), (unsigned)out->h, (unsigned)out->n);
  cudaStream_t st = (cudaStream_t)stream;
  if (out->c == 1) tap_gather3x3_kernel<1><<<grid, 128, 0, st>>>(vps::tv<const float>(*z), vps::tv<float>(*out), bias, act, slope, out_scale);
  else if (out->c == 2) tap_gather3x3_kernel<2><<<grid, 128, 0, st>>>(vps::tv<const float>(*z), vps::tv<float>(*out), bias, act, slope, out_scale);
  else tap_gather3x3_kernel<3><<<grid, 128, 0, st>>>(vps::tv<const float>(*z), vps::tv<float>(*out), bias, act, slope, out_scale);
  VPS_CUDA_LAST("tap_gather3x3");
  return VPS_OK;
}
