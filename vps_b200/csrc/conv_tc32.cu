// fp32-parity convolution on tcgen05 tensor cores (sm_100a): the "tc32" precision.
//
// The reference computes every convolution in fp32 (cuDNN, e.g. resnet.py:506-517, flownet2.py:133-198) and
// north_star asks for label maps / track ids bit-exact and logits within 1e-3 of it -- which a single bf16 pass
// (8 significant bits per operand) cannot give.  This kernel keeps activations and results fp32 in HBM and feeds the
// tensor cores three products per K slab that together carry ~21 significant bits of every operand:
//
//     a = A + a_lo,  A = tf32(a) (round to nearest, 11 significant bits),  b = B + b_lo likewise
//     a*b ~= A*B  [kind::tf32]  +  bf16(a_lo)*bf16(b)  [kind::f16]  +  bf16(a)*bf16(b_lo)  [kind::f16]
//
// (dropped: a_lo*b_lo ~ 2^-24, and the bf16 rounding of the two correction products ~ 2^-21 each), all accumulated
// into the same fp32 TMEM accumulator.  Cost: 2 (tf32 runs at half the bf16 rate) + 1 + 1 = 4 bf16-equivalent MMA
// passes instead of the 6 a three-way bf16 split (bf16x6/"bf16x9"-style) needs for the same accuracy class; the
// CPU emulation of this arithmetic through the whole FuseTrack path (tools/emulate_split.py) reproduces the fp32
// oracle's label maps / ids / proposals bit-exactly, which a two-way bf16 split (bf16x3) does not.
//
// Pipeline per CTA (persistent, one 128-pixel x block_n tile at a time, K consumed 32 channels per step):
//   warp 0     : TMA producer.  Activation boxes arrive as raw fp32 {32 ch, pixels} (SWIZZLE_128B rows of 128 B); weights
//                come pre-split from vps_pack_weights_tc32 (tf32 plane + two bf16 planes) through their own ring.
//   warps 6-9  : converters.  Rewrite the fp32 box IN PLACE as tf32(a) and emit the two bf16 planes bf16(a - tf32(a)),
//                bf16(a) as SWIZZLE_64B operand tiles next to it (generic-proxy writes -> fence.proxy.async -> mbarrier).
//                In halo mode (stride 1, > 1 tap) one converted (th+kh-1) x (tw+kw-1) box feeds all kh*kw taps.
//   warp 1     : MMA issuer: per (tap, 32-channel chunk) 4 x tcgen05.mma.kind::tf32 (K = 8) + 2 x 2 x kind::f16 (K = 16).
//   warps 8-15 : promotion + epilogue.  tcgen05.mma adds into its fp32 accumulator with TRUNCATION (measured here: the
//                error of a K-long chain grows like (#MMAs) * 2^-24, biased towards zero -- 1.3e-4 after 2300 MMAs), so a
//                chain is cut into groups of `group` K steps: the MMA warp starts every group on a fresh TMEM buffer
//                (accumulate = 0, buffers ping-pong), these warps drain finished groups with tcgen05.ld and keep the running
//                sum in registers with round-to-nearest fp32 adds (the classic fix for emulated-fp32 tensor-core GEMMs),
//                then apply bias / activation / residual and store.  setmaxnreg moves registers from the producer /
//                converter warpgroups to these two (128 running sums per thread for a 256-wide tile).
#include "conv_tc_common.cuh"

namespace {

constexpr int T32_EPI_WARPS = 8;            // warps 8..15: two per TMEM lane quarter, alternating 32-column chunks
constexpr int T32_CONV_WARPS = 6;           // warps 2..7
constexpr int T32_THREADS = 64 + 32 * (T32_EPI_WARPS + T32_CONV_WARPS);     // 512 = 4 warpgroups
constexpr int T32_MAX_BUF = 4;              // TMEM accumulator buffers (512 columns / block_n, at most 4)
constexpr int T32_REGS_LOW = 56, T32_REGS_HIGH = 200;    // setmaxnreg: 256 * 56 + 256 * 200 = 65536
constexpr int T32_KC = 32;                 // channels per K step: 128-byte tf32 rows, 64-byte bf16 rows

struct Tc32Extra {
  int rows;                  // activation rows (pixels) per A item: halo_h * halo_w, or 128
  int a_l_off, a_h_off;      // byte offsets of the two bf16 planes inside an A ring slot
  int b_half_bytes;          // block_n * 128: one B ring slot holds either the tf32 weight tile or the two bf16 tiles
  int nk8_last, nk16_last;   // K8 / K16 slabs of the last channel chunk that hold real channels
  int group;                 // K steps accumulated inside the tensor core before the sum is promoted to registers
  int alt_neg;               // 1: odd groups accumulate -A*B (instruction-descriptor negate bit) and are subtracted
  int nbuf, buf_cols;        // TMEM accumulator buffers and their column pitch
};

struct Ring32 {
  uint32_t a_base, a_stage_bytes, b_base, b_stage_bytes, bar_base;
  __device__ __forceinline__ uint32_t afull(int s) const { return bar_base + 8u * s; }
  __device__ __forceinline__ uint32_t aconv(int s) const { return bar_base + 8u * (MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t aempty(int s) const { return bar_base + 8u * (2 * MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t bfull(int s) const { return bar_base + 8u * (3 * MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t bempty(int s) const { return bar_base + 8u * (4 * MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t gfull(int a) const { return bar_base + 8u * (5 * MAX_STAGES + a); }
  __device__ __forceinline__ uint32_t gempty(int a) const { return bar_base + 8u * (5 * MAX_STAGES + T32_MAX_BUF + a); }
  __device__ __forceinline__ uint32_t tmem_slot() const { return bar_base + 8u * (5 * MAX_STAGES + 2 * T32_MAX_BUF); }
};
constexpr int T32_NBAR = 5 * MAX_STAGES + 2 * T32_MAX_BUF;
constexpr int T32_BAR_BYTES = 8 * (T32_NBAR + 2);

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile descriptor: layout 2 = SWIZZLE_128B (128-byte rows), 4 = SWIZZLE_64B (64-byte rows)
__device__ __forceinline__ uint64_t desc_hi(uint32_t layout, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint32_t cvt_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return u;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {     // lo -> bits [0,16)
  __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&b);
}

// ---------------------------------------------------------------- tile walk shared by all roles
struct TileCoord {
  int prob, n_idx, img, ty, tx;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvTcParams& p, int tile) {
  TileCoord t;
  t.prob = tile / p.tiles_per_prob;
  const int t_in = tile - t.prob * p.tiles_per_prob;
  t.n_idx = t_in % p.n_tiles_n;
  const int m_idx = t_in / p.n_tiles_n;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  t.img = m_idx / tiles_per_img;
  const int rem = m_idx - t.img * tiles_per_img;
  t.ty = rem / p.tiles_x;
  t.tx = rem - t.ty * p.tiles_x;
  return t;
}

// ---------------------------------------------------------------- warp 0: TMA producer
__device__ __forceinline__ void producer32(const ConvTcParams& p, const Tc32Extra& e, const Ring32& rg, const CUtensorMap* tmA,
                                           const CUtensorMap* tmBt, const CUtensorMap* tmBhl) {
  const int ntaps = p.kh * p.kw, kw = p.kw;
  const bool halo = p.halo != 0;
  const uint32_t a_box_bytes = (uint32_t)p.a_box_bytes, bhalf = (uint32_t)e.b_half_bytes;
  const int bn = p.block_n;
  int as = 0, bs = 0;
  uint32_t aphase = 0, bphase = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const TileCoord t = tile_coord(p, tile);
    const int x_base = t.tx * p.tw * p.sw - p.pw_[t.prob];
    const int y_base = t.ty * p.th * p.sh - p.ph_[t.prob];
    const int n0 = t.n_idx * bn;
    for (int cc = 0; cc < p.cin_chunks; ++cc) {
      int r = 0, s = 0;
      for (int tap = 0; tap < ntaps; ++tap) {
        if (!halo || tap == 0) {
          mbar_wait(rg.aempty(as), aphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(rg.afull(as), a_box_bytes);
            tma_load_4d(rg.a_base + as * rg.a_stage_bytes, tmA, rg.afull(as), cc * T32_KC, halo ? x_base : x_base + s,
                        halo ? y_base : y_base + r, t.img);
          }
          if (++as == p.a_stages) { as = 0; aphase ^= 1; }
        }
        // weight tiles of this (tap, chunk): slot 0 = tf32 plane, slot 1 = the two bf16 planes
        mbar_wait(rg.bempty(bs), bphase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(rg.bfull(bs), bhalf);
          tma_load_4d(rg.b_base + bs * rg.b_stage_bytes, tmBt, rg.bfull(bs), cc * T32_KC, n0, tap, t.prob);
        }
        if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
        mbar_wait(rg.bempty(bs), bphase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(rg.bfull(bs), bhalf);
          tma_load_5d(rg.b_base + bs * rg.b_stage_bytes, tmBhl, rg.bfull(bs), cc * T32_KC, n0, tap, t.prob, 0);
        }
        if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
        if (++s == kw) { s = 0; ++r; }
      }
    }
  }
}

// ---------------------------------------------------------------- warps 6..9: fp32 box -> tf32 (in place) + two bf16 planes
__device__ __forceinline__ void converter32(const ConvTcParams& p, const Tc32Extra& e, const Ring32& rg, int ctid) {
  const int ntaps = p.kh * p.kw;
  const int items_per_tile = p.cin_chunks * (p.halo ? 1 : ntaps);
  const int tasks = e.rows * 4;                       // 8 channels (two 16-byte fp32 chunks) per task
  int as = 0;
  uint32_t aphase = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    for (int it = 0; it < items_per_tile; ++it) {
      mbar_wait(rg.afull(as), aphase);
      const uint32_t slot = rg.a_base + as * rg.a_stage_bytes;
      for (int task = ctid; task < tasks; task += 32 * T32_CONV_WARPS) {
        const int r = task >> 2, j = task & 3;
        const uint32_t row_t = slot + (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        const uint32_t p0 = row_t + (((uint32_t)(2 * j) ^ sw) << 4), p1 = row_t + (((uint32_t)(2 * j + 1) ^ sw) << 4);
        float v[8];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(p0));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(p1));
        uint32_t t[8];
        float lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          t[q] = cvt_tf32(v[q]);
          lo[q] = v[q] - __uint_as_float(t[q]);       // exact: both share the exponent range, <= 13 significant bits
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p0), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p1), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]) : "memory");
        // bf16 planes: 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row r sits at chunk j ^ ((r >> 1) & 3)
        const uint32_t off64 = (uint32_t)r * 64u + ((((uint32_t)j) ^ ((uint32_t)(r >> 1) & 3u)) << 4);
        const uint32_t pl = slot + (uint32_t)e.a_l_off + off64, ph = slot + (uint32_t)e.a_h_off + off64;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pl), "r"(pack_bf16(lo[0], lo[1])), "r"(pack_bf16(lo[2], lo[3])),
                     "r"(pack_bf16(lo[4], lo[5])), "r"(pack_bf16(lo[6], lo[7])) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ph), "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3])),
                     "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor-core reads
      mbar_arrive(rg.aconv(as));
      if (++as == p.a_stages) { as = 0; aphase ^= 1; }
    }
  }
}

// ---------------------------------------------------------------- warp 1: MMA issuer
__device__ __forceinline__ void mma32(const ConvTcParams& p, const Tc32Extra& e, const Ring32& rg, uint32_t tmem_base) {
  const uint32_t nfield = ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | nfield;     // D = f32, A = B = tf32, K-major
  const uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | nfield;     // D = f32, A = B = bf16
  const int ntaps = p.kh * p.kw, kw = p.kw;
  const bool halo = p.halo != 0;
  const uint32_t hw = (uint32_t)p.halo_w;
  const uint64_t a_hi_t = desc_hi(2, halo ? hw * 128u : 1024u), b_hi_t = desc_hi(2, 1024u);
  const uint64_t a_hi_h = desc_hi(4, halo ? hw * 64u : 512u), b_hi_h = desc_hi(4, 512u);
  const uint32_t b_l_off = (uint32_t)p.block_n * 64u;       // second bf16 weight plane inside its B slot
  int as = 0, bs = 0, gb = 0;
  uint32_t aphase = 0, bphase = 0, gphase = 0;       // gphase: one parity bit per TMEM buffer
  const int G = e.group, total_steps = p.cin_chunks * ntaps;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    int step = 0, in_group = 0, gidx = 0;
    uint32_t d_tmem = 0, first = 0, neg = 0;
    for (int cc = 0; cc < p.cin_chunks; ++cc) {
      const bool last = cc == p.cin_chunks - 1;
      const int nk8 = last ? e.nk8_last : 4, nk16 = last ? e.nk16_last : 2;
      int r = 0, s = 0, a_cur = 0;
      uint32_t a_slot = 0;
      for (int tap = 0; tap < ntaps; ++tap) {
        if (in_group == 0) {           // a new accumulation group starts on a drained TMEM buffer with accumulate = 0
          mbar_wait(rg.gempty(gb), ((gphase >> gb) & 1u) ^ 1u);
          tc_fence_after();
          d_tmem = tmem_base + (uint32_t)(gb * e.buf_cols);
          first = 0;
          neg = (e.alt_neg && (gidx & 1)) ? (1u << 13) : 0u;
          ++gidx;
        }
        if (!halo || tap == 0) {
          mbar_wait(rg.aconv(as), aphase);
          a_cur = as;
          a_slot = rg.a_base + as * rg.a_stage_bytes;
          if (++as == p.a_stages) { as = 0; aphase ^= 1; }
        }
        const uint32_t shift = halo ? (uint32_t)(r * (int)hw + s) : 0u;
        const uint32_t a_t = a_slot + shift * 128u;
        const uint32_t a_l = a_slot + (uint32_t)e.a_l_off + shift * 64u, a_h = a_slot + (uint32_t)e.a_h_off + shift * 64u;
        // ---- tf32 x tf32
        mbar_wait(rg.bfull(bs), bphase);
        tc_fence_after();
        {
          const uint32_t b_t = rg.b_base + bs * rg.b_stage_bytes;
          if (elect_one()) {
            const uint64_t ad = a_hi_t | (uint64_t)((a_t & 0x3FFFF) >> 4), bd = b_hi_t | (uint64_t)((b_t & 0x3FFFF) >> 4);
            umma_tf32(d_tmem, ad, bd, idesc_tf32 | neg, first);
            if (nk8 > 1) umma_tf32(d_tmem, ad + 2, bd + 2, idesc_tf32 | neg, 1u);
            if (nk8 > 2) umma_tf32(d_tmem, ad + 4, bd + 4, idesc_tf32 | neg, 1u);
            if (nk8 > 3) umma_tf32(d_tmem, ad + 6, bd + 6, idesc_tf32 | neg, 1u);
            umma_commit(rg.bempty(bs));
          }
        }
        first = 1;
        if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
        // ---- bf16(a_lo) x bf16(b)  +  bf16(a) x bf16(b_lo)
        mbar_wait(rg.bfull(bs), bphase);
        tc_fence_after();
        ++step;
        const bool close = (++in_group == G) || step == total_steps;
        {
          const uint32_t b_h = rg.b_base + bs * rg.b_stage_bytes, b_l = b_h + b_l_off;
          if (elect_one()) {
            const uint64_t ald = a_hi_h | (uint64_t)((a_l & 0x3FFFF) >> 4), ahd = a_hi_h | (uint64_t)((a_h & 0x3FFFF) >> 4);
            const uint64_t bhd = b_hi_h | (uint64_t)((b_h & 0x3FFFF) >> 4), bld = b_hi_h | (uint64_t)((b_l & 0x3FFFF) >> 4);
            umma_bf16(d_tmem, ald, bhd, idesc_bf16 | neg, 1u);
            if (nk16 > 1) umma_bf16(d_tmem, ald + 2, bhd + 2, idesc_bf16 | neg, 1u);
            umma_bf16(d_tmem, ahd, bld, idesc_bf16 | neg, 1u);
            if (nk16 > 1) umma_bf16(d_tmem, ahd + 2, bld + 2, idesc_bf16 | neg, 1u);
            umma_commit(rg.bempty(bs));
            if (!halo || tap == ntaps - 1) umma_commit(rg.aempty(a_cur));
            if (close) umma_commit(rg.gfull(gb));
          }
        }
        if (close) {
          gphase ^= 1u << gb;
          if (++gb == e.nbuf) gb = 0;
          in_group = 0;
        }
        if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
        if (++s == kw) { s = 0; ++r; }
      }
    }
  }
}

// ---------------------------------------------------------------- warps 8..15: promotion (TMEM groups -> register sums) + epilogue
// warp -> TMEM lane quarter q = warp % 4 (hardware restriction); the two warps of a quarter take alternate 32-column
// chunks, so a thread owns one output pixel and up to 4 x 32 channels of running sums.
template <int ACT>
__device__ __forceinline__ void promote_epilogue(const ConvTcParams& p, const Tc32Extra& e, const Ring32& rg, uint32_t tmem_base,
                                                 int warp, int lane) {
  const int q = warp & 3, half = (warp - 8) >> 2;
  const int row = q * 32 + lane;
  const int ty_in = row / p.tw, tx_in = row - ty_in * p.tw;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const int total_steps = p.cin_chunks * p.kh * p.kw;
  const int ngroups = (total_steps + e.group - 1) / e.group;
  const int bn = p.block_n;
  int gb = 0;
  uint32_t gphase = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    float sum[4][32];
    for (int g = 0; g < ngroups; ++g) {
      mbar_wait(rg.gfull(gb), (gphase >> gb) & 1u);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(gb * e.buf_cols);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c0 = (half + 2 * k) * 32;
        if (c0 < bn) {
          uint32_t r[32];
          tmem_ld32(t_row + (uint32_t)c0, r);
          tmem_ld_wait();
          if (g == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[k][j] = __uint_as_float(r[j]);
          } else if (e.alt_neg && (g & 1)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[k][j] = __fsub_rn(sum[k][j], __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[k][j] = __fadd_rn(sum[k][j], __uint_as_float(r[j]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(rg.gempty(gb));
      gphase ^= 1u << gb;
      if (++gb == e.nbuf) gb = 0;
    }
    // ---- bias / activation / residual / store of this tile (same arithmetic as conv_tc.cu's epilogue)
    const int prob = tile / p.tiles_per_prob;
    const int t_in = tile - prob * p.tiles_per_prob;
    const int n_idx = t_in % p.n_tiles_n;
    const int m_idx = t_in / p.n_tiles_n;
    const int img = m_idx / tiles_per_img;
    const int rem = m_idx - img * tiles_per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int oy = ty * p.th + ty_in, ox = tx * p.tw + tx_in;
    const bool valid = (oy < p.oh) && (ox < p.ow);
    const int64_t pix = ((int64_t)img * p.y_h + (oy * p.oy_mul + p.oy_off_[prob])) * p.y_w + (ox * p.ox_mul + p.ox_off_[prob]);
    const int nbase = n_idx * bn;
    const int nlim = min(p.cout, nbase + bn);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c0 = (half + 2 * k) * 32;
      if (c0 < bn && valid && nbase + c0 < nlim) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(sum[k][j]);
        epi_chunk<ACT>(p, r, pix, nbase + c0, nlim);
      }
    }
  }
}

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(T32_THREADS, 1)
conv_igemm_tc32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBt,
                       const __grid_constant__ CUtensorMap tmBhl, const ConvTcParams p, const Tc32Extra e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  Ring32 rg;
  rg.a_base = smem_base; rg.a_stage_bytes = (uint32_t)p.a_stage_bytes;
  rg.b_base = smem_base + (uint32_t)p.a_stages * rg.a_stage_bytes;
  rg.b_stage_bytes = (uint32_t)e.b_half_bytes;
  rg.bar_base = rg.b_base + (uint32_t)p.b_stages * rg.b_stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 2) {
    for (int i = lane; i < T32_NBAR; i += 32) {
      uint32_t count = 1;
      if (i >= MAX_STAGES && i < 2 * MAX_STAGES) count = 32 * T32_CONV_WARPS;          // aconv: every converter thread
      if (i >= 5 * MAX_STAGES + T32_MAX_BUF) count = 32 * T32_EPI_WARPS;              // gempty: every promotion thread
      mbar_init(rg.bar_base + 8u * i, count);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBt) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBhl) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rg.tmem_slot()), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(rg.tmem_slot()) : "memory");
  // programmatic dependent launch: the prologue above overlaps the previous kernel's tail (see conv_tc.cu)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < 8) {
    // producer / MMA / converter warpgroups give registers away, the two promotion warpgroups take them
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T32_REGS_LOW));
    if (warp == 0) producer32(p, e, rg, &tmA, &tmBt, &tmBhl);
    else if (warp == 1) mma32(p, e, rg, tmem_base);
    else converter32(p, e, rg, (int)threadIdx.x - 64);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T32_REGS_HIGH));
    switch (p.act) {
      case VPS_ACT_RELU: promote_epilogue<VPS_ACT_RELU>(p, e, rg, tmem_base, warp, lane); break;
      case VPS_ACT_LRELU: promote_epilogue<VPS_ACT_LRELU>(p, e, rg, tmem_base, warp, lane); break;
      case VPS_ACT_SIGMOID: promote_epilogue<VPS_ACT_SIGMOID>(p, e, rg, tmem_base, warp, lane); break;
      default: promote_epilogue<VPS_ACT_NONE>(p, e, rg, tmem_base, warp, lane); break;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- weight packing
// planes [Bt fp32 (tf32-rounded) | Bh bf16 = bf16(w) | Bl bf16 = bf16(w - tf32(w))], each [nprob][cout_pad][tap][cin_pad]
__global__ void pack_weights_tc32_kernel(const float* __restrict__ src, const float* __restrict__ scale, float* __restrict__ bt,
                                         __nv_bfloat16* __restrict__ bh, __nv_bfloat16* __restrict__ bl, int cout, int cin, int kh,
                                         int kw, int cout_pad, int cin_pad, int transposed) {
  const int64_t total = (int64_t)cout_pad * kh * kw * cin_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    int64_t t = i / cin_pad;
    const int s = (int)(t % kw); t /= kw;
    const int r = (int)(t % kh); t /= kh;
    const int co = (int)t;
    float v = 0.f;
    if (co < cout && ci < cin) {
      const int64_t si = transposed ? ((((int64_t)ci * cout + co) * kh + r) * kw + s) : ((((int64_t)co * cin + ci) * kh + r) * kw + s);
      v = src[si];
      if (scale) v *= scale[co];
    }
    const float tv = __uint_as_float(cvt_tf32(v));
    bt[i] = tv;
    bh[i] = __float2bfloat16_rn(v);
    bl[i] = __float2bfloat16_rn(v - tv);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode32() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}
int g_num_sms32 = 0;

inline int64_t plane_elems(int cout, int cin, int kh, int kw) {
  const int64_t cout_pad = (cout + 15) / 16 * 16, cin_pad = (cin + T32_KC - 1) / T32_KC * T32_KC;
  return cout_pad * kh * kw * cin_pad;
}

}  // namespace

extern "C" int64_t vps_packed_tc32_bytes(int cout, int cin, int kh, int kw, int nprob) {
  return plane_elems(cout, cin, kh, kw) * 8 * nprob;
}

// problem `prob` of `nprob` (the stride phases of a transposed convolution share one packed buffer; nprob = 1 otherwise)
extern "C" int vps_pack_weights_tc32(const float* w, const float* scale, void* dst, int cout, int cin, int kh, int kw,
                                     int transposed, int prob, int nprob, void* stream) {
  VPS_CHECK_ARG(nprob >= 1 && nprob <= MAX_PROB && prob >= 0 && prob < nprob, "pack_weights_tc32: prob %d of %d", prob, nprob);
  const int cout_pad = (cout + 15) / 16 * 16, cin_pad = (cin + T32_KC - 1) / T32_KC * T32_KC;
  const int64_t n = plane_elems(cout, cin, kh, kw);
  float* bt = (float*)dst + (int64_t)prob * n;
  __nv_bfloat16* bh = (__nv_bfloat16*)((char*)dst + 4 * n * nprob) + (int64_t)prob * n;
  __nv_bfloat16* bl = (__nv_bfloat16*)((char*)dst + 6 * n * nprob) + (int64_t)prob * n;
  const int blocks = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
  pack_weights_tc32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, scale, bt, bh, bl, cout, cin, kh, kw, cout_pad, cin_pad,
                                                                    transposed);
  VPS_CUDA_LAST("pack_weights_tc32");
  return VPS_OK;
}

// fp32 activations, fp32 (or bf16) output; args[i].w = the shared buffer of vps_pack_weights_tc32(.., prob i, nprob)
extern "C" int vps_conv2d_tc32_multi(const vps_conv_args* args, int nprob, void* stream) {
  VPS_CHECK_ARG(nprob >= 1 && nprob <= MAX_PROB, "conv2d_tc32: nprob %d", nprob);
  const vps_conv_args* a = &args[0];
  VPS_CHECK_ARG(a->x.dtype == VPS_F32, "conv2d_tc32: x must be fp32");
  VPS_CHECK_ARG(a->x.cs % 4 == 0 && ((uintptr_t)a->x.ptr & 15) == 0, "conv2d_tc32: x not 16B aligned (cs=%d)", a->x.cs);
  VPS_CHECK_ARG(a->sh >= 1 && a->sh <= 2 && a->sw >= 1 && a->sw <= 2, "conv2d_tc32: stride must be 1 or 2");
  VPS_CHECK_ARG(a->cin == a->x.c, "conv2d_tc32: cin %d != x.c %d", a->cin, a->x.c);
  VPS_CHECK_ARG(((uintptr_t)a->w & 127) == 0, "conv2d_tc32: weights not aligned");
  for (int i = 0; i < nprob; ++i) {
    VPS_CHECK_ARG(args[i].w == a->w && args[i].x.ptr == a->x.ptr && args[i].y.ptr == a->y.ptr && args[i].kh == a->kh &&
                      args[i].kw == a->kw && args[i].oh == a->oh && args[i].ow == a->ow && args[i].cout == a->cout &&
                      args[i].bias == a->bias && args[i].act == a->act && args[i].oy_mul == a->oy_mul && args[i].ox_mul == a->ox_mul,
                  "conv2d_tc32_multi: problems must share geometry and the packed weight buffer");
  }
  auto encode = get_encode32();
  if (!encode) { vps::set_error("cuTensorMapEncodeTiled unavailable"); return VPS_E_CUDA; }
  if (!g_num_sms32) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms32, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms32 <= 0) { vps::set_error("no device"); return VPS_E_NODEV; }
  }
  ConvTcParams p = {};
  Tc32Extra e = {};
  p.bk = T32_KC;
  const int cin_pad = (a->cin + T32_KC - 1) / T32_KC * T32_KC;
  const int cout_pad = (a->cout + 15) / 16 * 16;
  p.n_img = a->x.n; p.oh = a->oh; p.ow = a->ow;
  const bool halo = a->sh == 1 && a->sw == 1 && a->kh * a->kw > 1 && a->kh <= 8 && a->kw <= 8;
  p.halo = halo ? 1 : 0;
  if (halo) {
    p.tw = 8; p.th = 16;
  } else {
    int best_tw = 16; int64_t best_area = -1;
    const int cands[5] = {16, 8, 32, 64, 128};
    for (int i = 0; i < 5; ++i) {
      const int tw = cands[i], th = 128 / tw;
      if (tw * a->sw > 256 || th * a->sh > 256) continue;
      const int64_t area = (int64_t)vps::cdiv(a->ow, tw) * tw * vps::cdiv(a->oh, th) * th;
      if (best_area < 0 || area < best_area) { best_area = area; best_tw = tw; }
    }
    p.tw = best_tw; p.th = 128 / best_tw;
  }
  p.halo_w = p.tw + a->kw - 1;
  const int halo_h = p.th + a->kh - 1;
  e.rows = halo ? halo_h * p.halo_w : BLOCK_M;
  p.a_box_bytes = e.rows * 128;
  e.a_l_off = (e.rows * 128 + 1023) / 1024 * 1024;
  e.a_h_off = e.a_l_off + (e.rows * 64 + 1023) / 1024 * 1024;
  p.a_stage_bytes = e.a_h_off + (e.rows * 64 + 1023) / 1024 * 1024;
  p.tiles_x = vps::cdiv(a->ow, p.tw); p.tiles_y = vps::cdiv(a->oh, p.th);
  p.kh = a->kh; p.kw = a->kw; p.sh = a->sh; p.sw = a->sw;
  p.cin_chunks = cin_pad / T32_KC;
  const int rem = a->cin - (p.cin_chunks - 1) * T32_KC;
  e.nk8_last = (rem + 7) / 8; e.nk16_last = (rem + 15) / 16;
  const int ntaps = a->kh * a->kw;
  p.a_stages = halo ? 2 : 3;
  const int smem_budget = 227 * 1024 - 1024 - T32_BAR_BYTES - 64;
  // N tile: divisor of cout_pad (multiple of 16, <= 256) minimising waves * (steps * step clocks + epilogue); a step is
  // 8 MMAs = 4*bn clocks at the MMA floor, ~350 clocks of issue / barrier latency, or its weight bytes at the L2 rate
  int block_n = 16;
  {
    const int64_t m_tiles = (int64_t)a->x.n * p.tiles_y * p.tiles_x * nprob;
    double best = -1.0;
    for (int bn = 16; bn <= 256 && bn <= cout_pad; bn += 16) {
      if (cout_pad % bn) continue;
      if (p.a_stages * p.a_stage_bytes + 2 * bn * 128 > smem_budget) continue;
      const int64_t tiles = m_tiles * (cout_pad / bn);
      const double waves = (double)((tiles + g_num_sms32 - 1) / g_num_sms32);
      const double step = fmax(fmax(350.0, 4.0 * bn), (double)(bn * 256) / 56.0);
      const double t = waves * ((double)(p.cin_chunks * ntaps) * step + 40.0 * bn + 1500.0);
      if (best < 0 || t < best * 0.999) { best = t; block_n = bn; }
    }
  }
  p.block_n = block_n; p.n_tiles_n = cout_pad / block_n;
  e.b_half_bytes = block_n * 128;
  {
    int bst = (smem_budget - p.a_stages * p.a_stage_bytes) / e.b_half_bytes;
    p.b_stages = bst > MAX_STAGES ? MAX_STAGES : bst;
    VPS_CHECK_ARG(p.b_stages >= 2, "conv2d_tc32: ring does not fit (%d x %d px halo, bn %d)", halo_h, p.halo_w, block_n);
  }
  static int group_env = -1;
  if (group_env < 0) { const char* ev = getenv("VPS_TC32_GROUP"); group_env = ev ? atoi(ev) : 4; }
  e.group = group_env < 1 ? 1 : group_env;
  static int neg_env = -1;
  if (neg_env < 0) { const char* ev = getenv("VPS_TC32_NEG"); neg_env = ev ? atoi(ev) : 0; }
  e.alt_neg = neg_env;
  e.nbuf = block_n > 128 ? 2 : 4;
  e.buf_cols = TMEM_COLS / e.nbuf;
  p.nprob = nprob;
  p.tiles_per_prob = p.n_img * p.tiles_y * p.tiles_x * p.n_tiles_n;
  p.total_tiles = p.tiles_per_prob * nprob;
  p.y = a->y.ptr; p.y_h = a->y.h; p.y_w = a->y.w; p.y_cs = a->y.cs; p.y_dtype = a->y.dtype;
  const int esz = a->y.dtype == VPS_BF16 ? 2 : 4;
  p.y_vec = (((uintptr_t)a->y.ptr & 15) == 0) && ((a->y.cs * esz) % 16 == 0);
  if (p.y_vec && (((uintptr_t)a->y.ptr & 31) == 0) && ((a->y.cs * esz) % 32 == 0)) p.y_vec = 2;
  p.oy_mul = a->oy_mul; p.ox_mul = a->ox_mul;
  for (int i = 0; i < MAX_PROB; ++i) {
    const vps_conv_args* q = &args[i < nprob ? i : 0];
    p.ph_[i] = q->ph; p.pw_[i] = q->pw; p.oy_off_[i] = q->oy_off; p.ox_off_[i] = q->ox_off;
    VPS_CHECK_ARG((a->oh - 1) * a->oy_mul + q->oy_off < a->y.h && (a->ow - 1) * a->ox_mul + q->ox_off < a->y.w,
                  "conv2d_tc32: output mapping out of range");
  }
  p.res = a->res.ptr; p.res_cs = a->res.cs; p.res_dtype = a->res.dtype; p.res_after_act = a->res_after_act;
  p.res_vec = a->res.ptr && (((uintptr_t)a->res.ptr & 15) == 0) && (a->res.cs % 8 == 0);
  if (p.res_vec && (((uintptr_t)a->res.ptr & 31) == 0) && (a->res.cs % 16 == 0)) p.res_vec = 2;
  VPS_CHECK_ARG(!a->bias || ((uintptr_t)a->bias & 15) == 0, "conv2d_tc32: bias must be 16-byte aligned");
  p.bias = a->bias; p.cout = a->cout; p.act = a->act; p.slope = a->slope; p.out_scale = a->out_scale;
  if (a->res.ptr) VPS_CHECK_ARG(a->res.h == a->y.h && a->res.w == a->y.w, "conv2d_tc32: residual geometry");
  p.stats = nullptr;
  if (p.total_tiles == 0) return VPS_OK;

  CUtensorMap tmA, tmBt, tmBhl;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a->x.c, (cuuint64_t)a->x.w, (cuuint64_t)a->x.h, (cuuint64_t)a->x.n};
    cuuint64_t strides[3] = {(cuuint64_t)a->x.cs * 4, (cuuint64_t)a->x.w * a->x.cs * 4, (cuuint64_t)a->x.h * a->x.w * a->x.cs * 4};
    cuuint32_t box[4] = {(cuuint32_t)T32_KC, (cuuint32_t)(halo ? p.halo_w : p.tw * a->sw), (cuuint32_t)(halo ? halo_h : p.th * a->sh), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)a->sw, (cuuint32_t)a->sh, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->x.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vps::set_error("conv2d_tc32: encode A failed (%d) dims %d,%d,%d,%d cs %d", (int)r, a->x.c, a->x.w, a->x.h, a->x.n, a->x.cs);
      return VPS_E_CUDA;
    }
  }
  const int64_t n_plane = (int64_t)cout_pad * ntaps * cin_pad;
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)ntaps, (cuuint64_t)nprob};
    cuuint64_t strides[3] = {(cuuint64_t)ntaps * cin_pad * 4, (cuuint64_t)cin_pad * 4, (cuuint64_t)n_plane * 4};
    cuuint32_t box[4] = {(cuuint32_t)T32_KC, (cuuint32_t)block_n, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmBt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a->w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vps::set_error("conv2d_tc32: encode Bt failed (%d)", (int)r); return VPS_E_CUDA; }
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)ntaps, (cuuint64_t)nprob, 2};
    cuuint64_t strides[4] = {(cuuint64_t)ntaps * cin_pad * 2, (cuuint64_t)cin_pad * 2, (cuuint64_t)n_plane * 2,
                             (cuuint64_t)n_plane * nprob * 2};
    cuuint32_t box[5] = {(cuuint32_t)T32_KC, (cuuint32_t)block_n, 1, 1, 2};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmBhl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (char*)a->w + 4 * n_plane * nprob, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vps::set_error("conv2d_tc32: encode Bhl failed (%d)", (int)r); return VPS_E_CUDA; }
  }
  const int smem = p.a_stages * p.a_stage_bytes + p.b_stages * e.b_half_bytes + 1024 + T32_BAR_BYTES;
  static bool smem_set = false;
  if (!smem_set) {
    if (cudaFuncSetAttribute(conv_igemm_tc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      vps::set_error("conv2d_tc32: cannot raise dynamic smem: %s", cudaGetErrorString(cudaGetLastError()));
      return VPS_E_CUDA;
    }
    smem_set = true;
  }
  const int grid = p.total_tiles < g_num_sms32 ? p.total_tiles : g_num_sms32;
  static int pdl_env = -1;
  if (pdl_env < 0) { const char* ev = getenv("VPS_PDL"); pdl_env = ev ? atoi(ev) : 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(T32_THREADS); cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_env ? 1 : 0;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, conv_igemm_tc32_kernel, tmA, tmBt, tmBhl, p, e);
  if (le != cudaSuccess) { vps::set_error("conv2d_tc32: launch failed: %s", cudaGetErrorString(le)); return VPS_E_CUDA; }
  VPS_CUDA_LAST("conv_igemm_tc32_kernel");
  return VPS_OK;
}

extern "C" int vps_conv2d_tc32(const vps_conv_args* a, void* stream) { return vps_conv2d_tc32_multi(a, 1, stream); }
