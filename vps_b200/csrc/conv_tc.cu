// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   M = 128 output pixels (a th x tw patch of one image), N = block_n output channels (<= 256),
//   K = kh*kw*cin_pad consumed 64 channels (=128 B, one SWIZZLE_128B row) per pipeline stage.
//
//   warp 0  : TMA producer.  The A tile of one (r,s,channel-chunk) K-step is ONE 4-D TMA box
//             {64 ch, tw, th, 1} of the NHWC activation tensor, shifted by the filter tap; the
//             tensor map's element strides implement the conv stride and TMA's out-of-bounds zero
//             fill implements the padding -- im2col is never materialised.  B tile = 2-D box of the
//             packed weights [cout_pad][K].
//   warp 1  : allocates TMEM (512 columns = two fp32 accumulators of up to 256 columns) and issues
//             tcgen05.mma (M=128, N=block_n, K=16) x4 per stage from one elected lane; tcgen05.commit
//             releases smem stages and publishes finished accumulators.
//   warps 2-5: epilogue.  tcgen05.ld the accumulator (thread = output pixel, 32 channels per load),
//             bias + activation + residual + scale, convert, vectorised NHWC store (also into a
//             channel slice of a concat buffer / interleaved pixels for transposed convs).
//   Persistent: grid = #SMs, static round-robin over tiles, double-buffered accumulators so the
//   epilogue of tile i overlaps the main loop of tile i+1.
//
// Replaces the cuDNN/cuBLAS calls behind every nn.Conv2d/ConvTranspose2d/Linear of the reference path.
#include "conv_tc_common.cuh"

namespace {

constexpr int NUM_EPI_WARPS = 8;                      // 2 per TMEM lane quarter, alternating 32-column chunks
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;   // warp 0 = TMA, warp 1 = MMA, warps 2.. = epilogue


// ---------------------------------------------------------------- single-issuer roles (warps 0 and 1)
// Both run their loops warp-uniformly and pick the issuing lane with elect.sync: code under `if (lane == 0)` is
// divergent to the compiler, which then wraps every TMA / tcgen05 instruction in a uniformity loop.  A lone warp
// retires a dependent instruction every ~5 clk, so the per-step instruction count of these loops IS the pipeline
// rate for small N (measured: 110 SASS instructions = 500 clk per K step): the loops are specialised on the mode
// and, in halo mode, one barrier round covers a whole filter row (kw taps, 4*kw MMAs).
struct Ring {
  uint32_t a_base, a_stage_bytes, b_base, b_stage_bytes, bar_base;
  __device__ __forceinline__ uint32_t afull(int s) const { return bar_base + 8u * s; }
  __device__ __forceinline__ uint32_t aempty(int s) const { return bar_base + 8u * (MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t bfull(int s) const { return bar_base + 8u * (2 * MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t bempty(int s) const { return bar_base + 8u * (3 * MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t tfull(int a) const { return bar_base + 8u * (4 * MAX_STAGES + a); }
  __device__ __forceinline__ uint32_t tempty(int a) const { return bar_base + 8u * (4 * MAX_STAGES + 2 + a); }
};

// ---- halo mode: one activation box per channel chunk (A ring), weights per tap or per filter row (B ring)
template <bool ROWG, bool STATS>
__device__ __forceinline__ void producer_halo(const ConvTcParams& p, const Ring& rg, const CUtensorMap* tmA,
                                              const CUtensorMap* tmB0, const CUtensorMap* tmB1, const CUtensorMap* tmB2,
                                              const CUtensorMap* tmB3, int lane) {
  const int bk = p.bk, kw = p.kw, cin_chunks = p.cin_chunks, a_stages = p.a_stages, b_stages = p.b_stages;
  const int ngrp = ROWG ? p.kh : p.kh * p.kw;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const uint32_t a_box_bytes = (uint32_t)p.a_box_bytes;
  int as = 0, bs = 0;
  uint32_t aphase = 0, bphase = 0;
  long long st_a = 0, st_b = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int prob = tile / p.tiles_per_prob;
    const int t_in = tile - prob * p.tiles_per_prob;
    const int n_idx = t_in % p.n_tiles_n;
    const int m_idx = t_in / p.n_tiles_n;
    const int img = m_idx / tiles_per_img;
    const int rem = m_idx - img * tiles_per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int x_base = tx * p.tw - p.pw_[prob];
    const int y_base = ty * p.th - p.ph_[prob];
    const int n0 = n_idx * p.block_n;
    const CUtensorMap* tmB = prob == 0 ? tmB0 : (prob == 1 ? tmB1 : (prob == 2 ? tmB2 : tmB3));
    for (int cc = 0; cc < cin_chunks; ++cc) {
      const long long t0 = STATS ? clock64() : 0;
      mbar_wait(rg.aempty(as), aphase ^ 1);
      if (STATS) st_a += clock64() - t0;
      if (elect_one()) {
        mbar_expect_tx(rg.afull(as), a_box_bytes);
        tma_load_4d(rg.a_base + as * rg.a_stage_bytes, tmA, rg.afull(as), cc * bk, x_base, y_base, img);
      }
      if (++as == a_stages) { as = 0; aphase ^= 1; }
      for (int g = 0; g < ngrp; ++g) {
        const long long t1 = STATS ? clock64() : 0;
        mbar_wait(rg.bempty(bs), bphase ^ 1);
        if (STATS) st_b += clock64() - t1;
        if (elect_one()) {
          mbar_expect_tx(rg.bfull(bs), rg.b_stage_bytes);
          tma_load_3d(rg.b_base + bs * rg.b_stage_bytes, tmB, rg.bfull(bs), cc * bk, n0, ROWG ? g * kw : g);
        }
        if (++bs == b_stages) { bs = 0; bphase ^= 1; }
      }
    }
  }
  if (STATS && lane == 0) { p.stats[blockIdx.x * 8 + 0] = st_a; p.stats[blockIdx.x * 8 + 1] = st_b; }
}

// ---- flat mode (strided / 1x1 convolutions): K steps in (chunk, tap) order, gsub steps share one ring slot and one
// barrier round (both operands arrive on the slot's `afull` barrier; the B barriers are unused)
template <bool STATS>
__device__ __forceinline__ void producer_flat(const ConvTcParams& p, const Ring& rg, const CUtensorMap* tmA,
                                              const CUtensorMap* tmB0, const CUtensorMap* tmB1, const CUtensorMap* tmB2,
                                              const CUtensorMap* tmB3, int lane) {
  const int bk = p.bk, kw = p.kw, kh = p.kh, stages = p.a_stages, G = p.gsub;
  const int T = p.cin_chunks * kh * kw;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const uint32_t a_box_bytes = (uint32_t)p.a_box_bytes;
  const uint32_t b_tile_bytes = (uint32_t)p.block_n * (uint32_t)bk * 2u;
  int st = 0;
  uint32_t phase = 0;
  long long st_a = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int prob = tile / p.tiles_per_prob;
    const int t_in = tile - prob * p.tiles_per_prob;
    const int n_idx = t_in % p.n_tiles_n;
    const int m_idx = t_in / p.n_tiles_n;
    const int img = m_idx / tiles_per_img;
    const int rem = m_idx - img * tiles_per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int x_base = tx * p.tw * p.sw - p.pw_[prob];
    const int y_base = ty * p.th * p.sh - p.ph_[prob];
    const int n0 = n_idx * p.block_n;
    const CUtensorMap* tmB = prob == 0 ? tmB0 : (prob == 1 ? tmB1 : (prob == 2 ? tmB2 : tmB3));
    int cc = 0, r = 0, sx = 0;
    for (int q0 = 0; q0 < T; q0 += G) {
      const int cnt = min(G, T - q0);
      const long long t0 = STATS ? clock64() : 0;
      mbar_wait(rg.aempty(st), phase ^ 1);
      if (STATS) st_a += clock64() - t0;
      const uint32_t a_slot = rg.a_base + st * rg.a_stage_bytes, b_slot = rg.b_base + st * rg.b_stage_bytes;
      if (elect_one()) mbar_expect_tx(rg.afull(st), (uint32_t)cnt * (a_box_bytes + b_tile_bytes));
      __syncwarp();
      for (int j = 0; j < cnt; ++j) {
        if (elect_one()) {
          tma_load_4d(a_slot + j * a_box_bytes, tmA, rg.afull(st), cc * bk, x_base + sx, y_base + r, img);
          tma_load_3d(b_slot + j * b_tile_bytes, tmB, rg.afull(st), cc * bk, n0, r * kw + sx);
        }
        if (++sx == kw) { sx = 0; if (++r == kh) { r = 0; ++cc; } }
      }
      if (++st == stages) { st = 0; phase ^= 1; }
    }
  }
  if (STATS && lane == 0) { p.stats[blockIdx.x * 8 + 0] = st_a; p.stats[blockIdx.x * 8 + 1] = 0; }
}

// the (up to) four K16 MMAs of one tap: 64 channels = one SWIZZLE_128B row; +2 in the (addr >> 4) field = 32 bytes
template <bool BK64>
__device__ __forceinline__ void issue_tap(uint32_t d_tmem, uint64_t a_hi, uint64_t b_hi, uint32_t a_addr, uint32_t b_addr,
                                          uint32_t idesc, int nk, uint32_t accumulate) {
  const uint64_t adesc = a_hi | (uint64_t)((a_addr & 0x3FFFF) >> 4);
  const uint64_t bdesc = b_hi | (uint64_t)((b_addr & 0x3FFFF) >> 4);
  umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
  if (BK64) {
    if (nk > 1) umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
    if (nk > 2) umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
    if (nk > 3) umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
  }
}

template <bool ROWG, bool BK64, bool STATS>
__device__ __forceinline__ void mma_halo(const ConvTcParams& p, const Ring& rg, uint32_t tmem_base, int lane) {
  constexpr uint32_t row_bytes = BK64 ? 128u : 32u;
  // instruction descriptor: D=f32, A=B=bf16, both K-major, N=block_n, M=128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                         ((uint32_t)(BLOCK_M >> 4) << 24);
  const int kw = p.kw, cin_chunks = p.cin_chunks, a_stages = p.a_stages, b_stages = p.b_stages;
  const int ngrp = ROWG ? p.kh : p.kh * p.kw;
  const int nk_last = p.nk_last;
  const uint32_t halo_pitch = (uint32_t)p.halo_w * row_bytes;
  const uint32_t tap_b_bytes = (uint32_t)p.block_n * row_bytes;
  // descriptor high words are loop constants; the low word is (smem address >> 4)
  const uint64_t a_hi = make_smem_desc(0, BK64 ? 64 : 16, halo_pitch);
  const uint64_t b_hi = make_smem_desc(0, BK64 ? 64 : 16, 8u * row_bytes);
  int as = 0, bs = 0, acc = 0;
  uint32_t aphase = 0, bphase = 0, acc_phase = 0;
  long long st_a = 0, st_b = 0, st_t = 0;
  const long long t_begin = STATS ? clock64() : 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const long long t2 = STATS ? clock64() : 0;
    mbar_wait(rg.tempty(acc), acc_phase ^ 1);
    if (STATS) st_t += clock64() - t2;
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
    uint32_t first = 0;
    for (int cc = 0; cc < cin_chunks; ++cc) {
      const int nk = cc == cin_chunks - 1 ? nk_last : 4;
      const long long t0 = STATS ? clock64() : 0;
      mbar_wait(rg.afull(as), aphase);
      if (STATS) st_a += clock64() - t0;
      uint32_t a_row = rg.a_base + as * rg.a_stage_bytes, a_tap = a_row;
      const int a_cur = as;
      int sx = 0;
      if (++as == a_stages) { as = 0; aphase ^= 1; }
      for (int g = 0; g < ngrp; ++g) {
        const long long t1 = STATS ? clock64() : 0;
        mbar_wait(rg.bfull(bs), bphase);
        if (STATS) st_b += clock64() - t1;
        tc_fence_after();
        const uint32_t b_addr = rg.b_base + bs * rg.b_stage_bytes;
        if (elect_one()) {
          if (ROWG) {      // the kw taps of filter row g: A start moves one pixel (row_bytes) per tap
            uint32_t at = a_row, bt = b_addr;
            issue_tap<BK64>(d_tmem, a_hi, b_hi, at, bt, idesc, nk, first);
            for (int j = 1; j < kw; ++j) {
              at += row_bytes; bt += tap_b_bytes;
              issue_tap<BK64>(d_tmem, a_hi, b_hi, at, bt, idesc, nk, 1u);
            }
          } else {
            issue_tap<BK64>(d_tmem, a_hi, b_hi, a_tap, b_addr, idesc, nk, first);
          }
          umma_commit(rg.bempty(bs));
          if (g == ngrp - 1) umma_commit(rg.aempty(a_cur));
        }
        first = 1;
        if (++bs == b_stages) { bs = 0; bphase ^= 1; }
        if (ROWG) {
          a_row += halo_pitch;
        } else {          // next tap: one pixel to the right, or the start of the next halo row
          a_tap += row_bytes;
          if (++sx == kw) { sx = 0; a_row += halo_pitch; a_tap = a_row; }
        }
      }
    }
    if (elect_one()) umma_commit(rg.tfull(acc));
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
  if (STATS && lane == 0) {
    p.stats[blockIdx.x * 8 + 2] = st_a; p.stats[blockIdx.x * 8 + 3] = st_b; p.stats[blockIdx.x * 8 + 4] = st_t;
    p.stats[blockIdx.x * 8 + 5] = clock64() - t_begin;
  }
}

template <bool BK64, bool STATS>
__device__ __forceinline__ void mma_flat(const ConvTcParams& p, const Ring& rg, uint32_t tmem_base, int lane) {
  constexpr uint32_t row_bytes = BK64 ? 128u : 32u;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                         ((uint32_t)(BLOCK_M >> 4) << 24);
  const int stages = p.a_stages, G = p.gsub, ntaps = p.kh * p.kw;
  const int T = p.cin_chunks * ntaps;
  const int q_last = T - ntaps;                     // steps >= q_last belong to the last channel chunk
  const int nk_last = p.nk_last;
  const uint32_t a_box_bytes = (uint32_t)p.a_box_bytes, b_tile_bytes = (uint32_t)p.block_n * row_bytes;
  const uint64_t a_hi = make_smem_desc(0, BK64 ? 64 : 16, 8u * row_bytes);
  const uint64_t b_hi = a_hi;
  int st = 0, acc = 0;
  uint32_t phase = 0, acc_phase = 0;
  long long st_a = 0, st_t = 0;
  const long long t_begin = STATS ? clock64() : 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const long long t2 = STATS ? clock64() : 0;
    mbar_wait(rg.tempty(acc), acc_phase ^ 1);
    if (STATS) st_t += clock64() - t2;
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
    for (int q0 = 0; q0 < T; q0 += G) {
      const int cnt = min(G, T - q0);
      const long long t0 = STATS ? clock64() : 0;
      mbar_wait(rg.afull(st), phase);
      if (STATS) st_a += clock64() - t0;
      tc_fence_after();
      const uint32_t a_slot = rg.a_base + st * rg.a_stage_bytes, b_slot = rg.b_base + st * rg.b_stage_bytes;
      if (elect_one()) {
        for (int j = 0; j < cnt; ++j)
          issue_tap<BK64>(d_tmem, a_hi, b_hi, a_slot + j * a_box_bytes, b_slot + j * b_tile_bytes, idesc,
                          q0 + j >= q_last ? nk_last : 4, (uint32_t)((q0 + j) != 0));
        umma_commit(rg.aempty(st));
      }
      if (++st == stages) { st = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(rg.tfull(acc));
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
  if (STATS && lane == 0) {
    p.stats[blockIdx.x * 8 + 2] = st_a; p.stats[blockIdx.x * 8 + 3] = 0; p.stats[blockIdx.x * 8 + 4] = st_t;
    p.stats[blockIdx.x * 8 + 5] = clock64() - t_begin;
  }
}

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_igemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                     const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                     const __grid_constant__ CUtensorMap tmB3, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned operand ring (SWIZZLE_128B requirement)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t row_bytes = (uint32_t)p.bk * 2u;
  const uint32_t a_stage_bytes = (uint32_t)p.a_stage_bytes;
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * row_bytes * (p.halo ? (p.rowg ? (uint32_t)p.kw : 1u) : (uint32_t)p.gsub);
  const uint32_t b_base = smem_base + (uint32_t)p.a_stages * a_stage_bytes;
  const uint32_t bar_base = b_base + (uint32_t)p.b_stages * b_stage_bytes;
  // barrier slots (8 B each): afull, aempty, bfull, bempty [MAX_STAGES each], tmem_full[2], tmem_empty[2], tmem ptr
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (3 * MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (4 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (4 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (4 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // prologue on the critical path of every launch: one barrier per thread of warp 2 (36 inits), descriptors prefetched by
  // warp 0, TMEM allocated by warp 1 -- all concurrently
  if (warp == 2) {
    if (lane < 4 * MAX_STAGES) mbar_init(bar_base + 8u * lane, 1);                    // afull / aempty / bfull / bempty
    if (lane >= 28) {                                                               // lanes 28..31: tfull[0,1], tempty[0,1]
      const int a = lane & 1;
      if (lane < 30) mbar_init(tfull_bar(a), 1); else mbar_init(tempty_bar(a), 32 * NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the
  // tail of the previous kernel in the stream; no global memory is touched before the wait.  Our own dependents are
  // released immediately -- they block at their own wait until this grid has completed and flushed.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  Ring rg;
  rg.a_base = smem_base; rg.a_stage_bytes = a_stage_bytes; rg.b_base = b_base; rg.b_stage_bytes = b_stage_bytes;
  rg.bar_base = bar_base;
  const bool st = p.stats != nullptr;
  if (warp == 0) {
    // ===================== TMA producer =====================
#define VPS_ROLE(FN, ...) \
    do { if (st) FN<__VA_ARGS__, true>(p, rg, &tmA, &tmB0, &tmB1, &tmB2, &tmB3, lane); \
         else FN<__VA_ARGS__, false>(p, rg, &tmA, &tmB0, &tmB1, &tmB2, &tmB3, lane); } while (0)
    if (p.halo) { if (p.rowg) VPS_ROLE(producer_halo, true); else VPS_ROLE(producer_halo, false); }
    else { if (st) producer_flat<true>(p, rg, &tmA, &tmB0, &tmB1, &tmB2, &tmB3, lane);
           else producer_flat<false>(p, rg, &tmA, &tmB0, &tmB1, &tmB2, &tmB3, lane); }
#undef VPS_ROLE
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
#define VPS_MMA(FN, ...) \
    do { if (st) FN<__VA_ARGS__, true>(p, rg, tmem_base, lane); else FN<__VA_ARGS__, false>(p, rg, tmem_base, lane); } while (0)
    if (p.bk == 64) {
      if (p.halo) { if (p.rowg) VPS_MMA(mma_halo, true, true); else VPS_MMA(mma_halo, false, true); }
      else VPS_MMA(mma_flat, true);
    } else {
      if (p.halo) { if (p.rowg) VPS_MMA(mma_halo, true, false); else VPS_MMA(mma_halo, false, false); }
      else VPS_MMA(mma_flat, false);
    }
#undef VPS_MMA
  } else {
    // ===================== epilogue (warps 2..9) =====================
    switch (p.act) {
      case VPS_ACT_RELU: epilogue_loop<VPS_ACT_RELU, NUM_EPI_WARPS / 4>(p, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane); break;
      case VPS_ACT_LRELU: epilogue_loop<VPS_ACT_LRELU, NUM_EPI_WARPS / 4>(p, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane); break;
      case VPS_ACT_SIGMOID: epilogue_loop<VPS_ACT_SIGMOID, NUM_EPI_WARPS / 4>(p, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane); break;
      default: epilogue_loop<VPS_ACT_NONE, NUM_EPI_WARPS / 4>(p, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane); break;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------- fused deformable convolution (DCNv1, 3x3, pad 1)
// Same implicit GEMM, but the A tile of a K step is not a TMA box: eight producer warps bilinearly sample the input
// at the learned offsets (deform_conv_cuda_kernel.cu: deformable_im2col) straight into the SWIZZLE_128B operand slot,
// so the 9x column matrix is never written to HBM.  K steps run chunk-major / tap-minor: for one 64-channel chunk the
// nine taps of a tile touch the same ~(th+2) x (tw+2) x 128 B of input, which stays in L1.  Per tile the sampling
// set-up of every (pixel, tap) -- 4 corner element offsets + 4 weights -- is computed once into shared memory.
constexpr int DCN_GATHER_WARPS = 16;     // sampling is ALU-issue bound: 4 producer warps per SM sub-partition
constexpr int DCN_EPI_WARPS = 4;
constexpr int DCN_THREADS = 64 + 32 * DCN_EPI_WARPS + 32 * DCN_GATHER_WARPS;
constexpr int DCN_SETUP_BYTES = 9 * BLOCK_M * 32;

struct DcnParams {
  const __nv_bfloat16* x;
  const float* off;
  int x_cs, off_cs, H, W;
};

__device__ __forceinline__ void dcn_gather_loop(const ConvTcParams& p, const DcnParams& d, const Ring& rg, uint32_t setup_base,
                                                int gtid) {
  const int stages = p.a_stages, cin_chunks = p.cin_chunks;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const int H = d.H, W = d.W;
  int st = 0;
  uint32_t phase = 0;
  const int j = gtid & 7;                    // 16-byte channel chunk of the 128-byte row
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int img = tile / tiles_per_img;
    const int rem = tile - img * tiles_per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    // ---- sampling set-up of all (tap, pixel) pairs of this tile
    for (int item = gtid; item < 9 * BLOCK_M; item += 32 * DCN_GATHER_WARPS) {
      const int k = item >> 7, r = item & (BLOCK_M - 1);
      const int ty_in = r / p.tw, tx_in = r - ty_in * p.tw;
      const int yo = ty * p.th + ty_in, xo = tx * p.tw + tx_in;
      float wts[4] = {0.f, 0.f, 0.f, 0.f};
      int offs[4] = {0, 0, 0, 0};
      if (yo < H && xo < W) {
        const float* op = d.off + ((int64_t)(img * H + yo) * W + xo) * d.off_cs;
        const float oh = __ldg(op + 2 * k), ow = __ldg(op + 2 * k + 1);
        const float h = (float)(yo - 1 + k / 3) + oh;
        const float w = (float)(xo - 1 + k % 3) + ow;
        if (h > -1.f && w > -1.f && h < (float)H && w < (float)W) {
          const int hl = (int)floorf(h), wl = (int)floorf(w);
          const int hh_ = hl + 1, wh_ = wl + 1;
          const float lh = h - (float)hl, lw = w - (float)wl;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const int base = img * H;
          if (hl >= 0 && wl >= 0) { wts[0] = hh * hw; offs[0] = ((base + hl) * W + wl) * d.x_cs; }
          if (hl >= 0 && wh_ <= W - 1) { wts[1] = hh * lw; offs[1] = ((base + hl) * W + wh_) * d.x_cs; }
          if (hh_ <= H - 1 && wl >= 0) { wts[2] = lh * hw; offs[2] = ((base + hh_) * W + wl) * d.x_cs; }
          if (hh_ <= H - 1 && wh_ <= W - 1) { wts[3] = lh * lw; offs[3] = ((base + hh_) * W + wh_) * d.x_cs; }
        }
      }
      const uint32_t sa = setup_base + (uint32_t)item * 32u;
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "f"(wts[0]), "f"(wts[1]), "f"(wts[2]), "f"(wts[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + 16u), "r"(offs[0]), "r"(offs[1]), "r"(offs[2]), "r"(offs[3]) : "memory");
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * DCN_GATHER_WARPS) : "memory");
    // ---- K steps: chunk-major, tap-minor
    for (int cc = 0; cc < cin_chunks; ++cc) {
      const __nv_bfloat16* xc = d.x + cc * 64 + j * 8;
      for (int k = 0; k < 9; ++k) {
        mbar_wait(rg.aempty(st), phase ^ 1);
        const uint32_t a_slot = rg.a_base + st * rg.a_stage_bytes;
        constexpr int ROWS_PER_PASS = 32 * DCN_GATHER_WARPS / 8;       // 8 lanes (16-byte chunks) per pixel row
#pragma unroll
        for (int i = 0; i < BLOCK_M / ROWS_PER_PASS; ++i) {
          const int r = (gtid >> 3) + ROWS_PER_PASS * i;
          const uint32_t sa = setup_base + (uint32_t)(k * BLOCK_M + r) * 32u;
          float w0, w1, w2, w3;
          int o0, o1, o2, o3;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w0), "=f"(w1), "=f"(w2), "=f"(w3) : "r"(sa));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(sa + 16u));
          const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(xc + o0));
          const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(xc + o1));
          const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(xc + o2));
          const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(xc + o3));
          const float wq[4] = {w0, w1, w2, w3};
          const uint4 qs[4] = {q0, q1, q2, q3};
          float acc[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) acc[t] = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {           // corners outside the image carry weight 0 (branch-free: finite inputs)
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&qs[q]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __bfloat1622float2(b2[t]);
              acc[2 * t] += wq[q] * f.x;
              acc[2 * t + 1] += wq[q] * f.y;
            }
          }
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __nv_bfloat162 b = __floats2bfloat162_rn(acc[2 * t], acc[2 * t + 1]);
            pk[t] = *reinterpret_cast<uint32_t*>(&b);
          }
          const uint32_t da = a_slot + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(da), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor-core reads
        mbar_arrive(rg.afull(st));
        if (++st == stages) { st = 0; phase ^= 1; }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * DCN_GATHER_WARPS) : "memory");   // set-up cache is rewritten next tile
  }
}

__global__ void __launch_bounds__(DCN_THREADS, 1)
dcn_igemm_tc_kernel(const __grid_constant__ CUtensorMap tmB, const ConvTcParams p, const DcnParams d) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  Ring rg;
  rg.a_base = smem_base; rg.a_stage_bytes = (uint32_t)p.a_stage_bytes;
  rg.b_base = smem_base + (uint32_t)p.a_stages * rg.a_stage_bytes;
  rg.b_stage_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t setup_base = rg.b_base + (uint32_t)p.a_stages * rg.b_stage_bytes;
  rg.bar_base = setup_base + DCN_SETUP_BYTES;
  const uint32_t tmem_slot = rg.bar_base + 8u * (4 * MAX_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(rg.afull(s), 1 + 32 * DCN_GATHER_WARPS);     // weight TMA (expect_tx arrival) + every gather thread
      mbar_init(rg.aempty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(rg.tfull(a), 1);
      mbar_init(rg.tempty(a), 32 * DCN_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // weight producer: one {64 ch, block_n, 1 tap} box per K step, completing on the step's `afull` barrier
    const int stages = p.a_stages;
    int st = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int cc = 0; cc < p.cin_chunks; ++cc) {
        for (int k = 0; k < 9; ++k) {
          mbar_wait(rg.aempty(st), phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(rg.afull(st), rg.b_stage_bytes);
            tma_load_3d(rg.b_base + st * rg.b_stage_bytes, &tmB, rg.afull(st), cc * 64, 0, k);
          }
          if (++st == stages) { st = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    mma_flat<true, false>(p, rg, tmem_base, lane);
  } else if (warp < 2 + DCN_EPI_WARPS) {
    epilogue_loop<VPS_ACT_NONE, DCN_EPI_WARPS / 4>(p, tmem_base, rg.tfull(0), rg.tempty(0), warp, lane);
  } else {
    dcn_gather_loop(p, d, rg, setup_base, (int)threadIdx.x - 32 * (2 + DCN_EPI_WARPS));
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- weight packing
// dst[co][ (r*kw+s)*cin_pad + ci ] (bf16), zero padded; src OIHW (or IOHW when transposed)
__global__ void pack_weights_tc_kernel(const float* __restrict__ src, const float* __restrict__ scale,
                                       __nv_bfloat16* __restrict__ dst, int cout, int cin, int kh, int kw,
                                       int cout_pad, int cin_pad, int transposed) {
  const int64_t total = (int64_t)cout_pad * kh * kw * cin_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    int64_t t = i / cin_pad;
    const int s = (int)(t % kw); t /= kw;
    const int r = (int)(t % kh); t /= kh;
    const int co = (int)t;
    float v = 0.f;
    if (co < cout && ci < cin) {
      const int64_t si = transposed ? ((((int64_t)ci * cout + co) * kh + r) * kw + s)
                                    : ((((int64_t)co * cin + ci) * kh + r) * kw + s);
      v = src[si];
      if (scale) v *= scale[co];
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------- host
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

int g_num_sms = 0;

}  // namespace

static inline int cin_pad_for(int cin, int gran) { return (cin + gran - 1) / gran * gran; }

extern "C" int64_t vps_packed_tc_bytes(int cout, int cin, int kh, int kw, int cin_gran) {
  const int64_t cout_pad = (cout + 15) / 16 * 16, cin_pad = cin_pad_for(cin, cin_gran == 16 ? 16 : 64);
  return cout_pad * kh * kw * cin_pad * 2;
}

extern "C" int vps_pack_weights_tc(const float* w, const float* scale, void* dst, int cout, int cin, int kh, int kw,
                                   int transposed, int cin_gran, void* stream) {
  const int cout_pad = (cout + 15) / 16 * 16, cin_pad = cin_pad_for(cin, cin_gran == 16 ? 16 : 64);
  const int64_t total = (int64_t)cout_pad * kh * kw * cin_pad;
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  pack_weights_tc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, scale, (__nv_bfloat16*)dst, cout, cin, kh,
                                                                  kw, cout_pad, cin_pad, transposed);
  VPS_CUDA_LAST("pack_weights_tc");
  return VPS_OK;
}

// nprob problems (<= 4) that share x / y / geometry / epilogue and differ in weights, padding and output pixel
// offset: the four stride phases of a transposed convolution run as ONE persistent launch.
extern "C" int vps_conv2d_tc_multi(const vps_conv_args* args, int nprob, void* stream) {
  VPS_CHECK_ARG(nprob >= 1 && nprob <= MAX_PROB, "conv2d_tc: nprob %d", nprob);
  const vps_conv_args* a = &args[0];
  VPS_CHECK_ARG(a->x.dtype == VPS_BF16, "conv2d_tc: x must be bf16");
  VPS_CHECK_ARG(a->x.cs % 8 == 0 && ((uintptr_t)a->x.ptr & 15) == 0, "conv2d_tc: x not 16B aligned (cs=%d)",
                a->x.cs);
  VPS_CHECK_ARG(a->sh >= 1 && a->sh <= 2 && a->sw >= 1 && a->sw <= 2, "conv2d_tc: stride must be 1 or 2");
  VPS_CHECK_ARG(a->cin == a->x.c, "conv2d_tc: cin %d != x.c %d", a->cin, a->x.c);
  for (int i = 0; i < nprob; ++i) {
    VPS_CHECK_ARG(((uintptr_t)args[i].w & 15) == 0, "conv2d_tc: weights not aligned");
    VPS_CHECK_ARG(args[i].x.ptr == a->x.ptr && args[i].y.ptr == a->y.ptr && args[i].kh == a->kh && args[i].kw == a->kw &&
                      args[i].oh == a->oh && args[i].ow == a->ow && args[i].cout == a->cout && args[i].bias == a->bias &&
                      args[i].act == a->act && args[i].oy_mul == a->oy_mul && args[i].ox_mul == a->ox_mul &&
                      args[i].cin_gran == a->cin_gran,
                  "conv2d_tc_multi: problems must share geometry");
  }
  auto encode = get_encode();
  if (!encode) { vps::set_error("cuTensorMapEncodeTiled unavailable"); return VPS_E_CUDA; }
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) { vps::set_error("no device"); return VPS_E_NODEV; }
  }

  ConvTcParams p;
  const int bk = a->cin_gran == 16 ? 16 : 64;
  p.bk = bk;
  const int cin_pad = cin_pad_for(a->cin, bk);
  const int cout_pad = (a->cout + 15) / 16 * 16;
  p.n_img = a->x.n; p.oh = a->oh; p.ow = a->ow;
  // halo mode (stride 1, more than one tap): the 8 rows of an MMA row group are 8 consecutive pixels of one halo row,
  // so the tile is 16 x 8 pixels and every tap reads the same (16+kh-1) x (8+kw-1) box at a shifted start address.
  static int halo_env = -1;
  if (halo_env < 0) {
    const char* e = getenv("VPS_CONV_HALO");
    halo_env = e ? atoi(e) : 2;                         // 0 = off, 1 = on wherever legal, 2 = heuristic
  }
  const bool halo_ok = a->sh == 1 && a->sw == 1 && a->kh * a->kw > 1 && a->kh <= 8 && a->kw <= 8;
  const bool halo = halo_ok && halo_env != 0;
  p.halo = halo ? 1 : 0;
  if (halo) {
    p.tw = 8; p.th = 16;
  } else {
    // pixel patch: minimise padded area; th*tw == 128, box extent tw*sw <= 256
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
  p.a_box_bytes = halo ? halo_h * p.halo_w * bk * 2 : BLOCK_M * bk * 2;
  p.a_stage_bytes = (p.a_box_bytes + 1023) / 1024 * 1024;
  p.tiles_x = vps::cdiv(a->ow, p.tw); p.tiles_y = vps::cdiv(a->oh, p.th);
  // N tile: pick the divisor of cout_pad (multiple of 16, <= 256) that minimises a simple time model
  //   waves(bn) * k_steps * max(fixed per-step latency, MMA time 2*bn clk, stage bytes / per-SM L2 bandwidth)
  // -- large tiles when there is enough parallelism, smaller N tiles to fill the persistent grid otherwise.
  int block_n = 16;
  {
    const int64_t m_tiles = (int64_t)a->x.n * p.tiles_y * p.tiles_x * nprob;
    double best = -1.0;
    for (int bn = 16; bn <= 256 && bn <= cout_pad; bn += 16) {
      if (cout_pad % bn) continue;
      const int64_t tiles = m_tiles * (cout_pad / bn);
      const double waves = (double)((tiles + g_num_sms - 1) / g_num_sms);
      const double epi = 40.0 * bn;     // epilogue clocks per tile (not hidden when a CTA runs a single tile)
      double t;
      if (halo) {   // per tap: MMA time / operand reads from smem / weight box; per chunk: one halo box
        const double step = fmax(fmax(215.0, 2.0 * bn), (double)(bn * bk * 2) / 40.0);
        t = waves * ((double)(cin_pad / bk) * ((double)(a->kh * a->kw) * step + (double)p.a_box_bytes / 20.0) + epi);
      } else {
        const double step = fmax(fmax(350.0, 2.0 * bn), (double)((BLOCK_M + bn) * bk * 2) / 80.0);
        t = waves * ((double)(a->kh * a->kw * (cin_pad / bk)) * step + epi);
      }
      if (best < 0 || t < best * 0.999) { best = t; block_n = bn; }
    }
  }
  p.block_n = block_n; p.n_tiles_n = cout_pad / block_n;
  p.kh = a->kh; p.kw = a->kw; p.sh = a->sh; p.sw = a->sw;
  p.cin_chunks = cin_pad / bk;
  // halo mode: one B ring slot = the kw taps of a filter row when that fits (<= 48 KB) -- one barrier round per row
  p.rowg = (halo && a->kw > 1 && a->kw * block_n * bk * 2 <= 48 * 1024) ? 1 : 0;
  p.nk_last = bk == 64 ? (a->cin - (p.cin_chunks - 1) * 64 + 15) / 16 : 1;
  // flat mode: gsub consecutive K steps share a ring slot (<= 48 KB of operands per barrier round, at most 4 steps)
  p.gsub = 1;
  if (!halo) {
    const int step_bytes = p.a_box_bytes + block_n * bk * 2;
    int g = (48 * 1024) / step_bytes;
    const int T = p.cin_chunks * a->kh * a->kw;
    if (g > 4) g = 4;
    if (g > T) g = T;
    if (g < 1) g = 1;
    p.gsub = g;
    p.a_stage_bytes = g * p.a_box_bytes;
  }
  const int b_stage_bytes = block_n * bk * 2 * (halo ? (p.rowg ? a->kw : 1) : p.gsub);
  if (halo) {
    p.a_stages = p.cin_chunks >= 3 ? 3 : 2;
    int bst = (200 * 1024 - p.a_stages * p.a_stage_bytes) / b_stage_bytes;
    p.b_stages = bst > MAX_STAGES ? MAX_STAGES : bst;
    VPS_CHECK_ARG(p.b_stages >= 2, "conv2d_tc: halo ring does not fit");
  } else {
    int stages = (200 * 1024) / (p.a_stage_bytes + b_stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    p.a_stages = p.b_stages = stages;
  }
  p.nprob = nprob;
  p.tiles_per_prob = p.n_img * p.tiles_y * p.tiles_x * p.n_tiles_n;
  p.total_tiles = p.tiles_per_prob * nprob;
  p.y = a->y.ptr; p.y_h = a->y.h; p.y_w = a->y.w; p.y_cs = a->y.cs; p.y_dtype = a->y.dtype;
  const int esz = a->y.dtype == VPS_BF16 ? 2 : 4;
  p.y_vec = (((uintptr_t)a->y.ptr & 15) == 0) && ((a->y.cs * esz) % 16 == 0);
  if (p.y_vec && (((uintptr_t)a->y.ptr & 31) == 0) && ((a->y.cs * esz) % 32 == 0)) p.y_vec = 2;     // 256-bit stores
  p.oy_mul = a->oy_mul; p.ox_mul = a->ox_mul;
  for (int i = 0; i < MAX_PROB; ++i) {
    const vps_conv_args* q = &args[i < nprob ? i : 0];
    p.ph_[i] = q->ph; p.pw_[i] = q->pw; p.oy_off_[i] = q->oy_off; p.ox_off_[i] = q->ox_off;
    VPS_CHECK_ARG((a->oh - 1) * a->oy_mul + q->oy_off < a->y.h && (a->ow - 1) * a->ox_mul + q->ox_off < a->y.w,
                  "conv2d_tc: output mapping out of range");
  }
  p.res = a->res.ptr; p.res_cs = a->res.cs; p.res_dtype = a->res.dtype; p.res_after_act = a->res_after_act;
  p.res_vec = a->res.ptr && (((uintptr_t)a->res.ptr & 15) == 0) && (a->res.cs % 8 == 0);
  if (p.res_vec && (((uintptr_t)a->res.ptr & 31) == 0) && (a->res.cs % 16 == 0)) p.res_vec = 2;       // 256-bit loads
  VPS_CHECK_ARG(!a->bias || ((uintptr_t)a->bias & 15) == 0, "conv2d_tc: bias must be 16-byte aligned");
  p.bias = a->bias; p.cout = a->cout; p.act = a->act; p.slope = a->slope; p.out_scale = a->out_scale;
  if (a->res.ptr) VPS_CHECK_ARG(a->res.h == a->y.h && a->res.w == a->y.w, "conv2d_tc: residual geometry");
  if (p.total_tiles == 0) return VPS_OK;

  CUtensorMap tmA, tmB[MAX_PROB];
  const CUtensorMapSwizzle swz = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a->x.c, (cuuint64_t)a->x.w, (cuuint64_t)a->x.h, (cuuint64_t)a->x.n};
    cuuint64_t strides[3] = {(cuuint64_t)a->x.cs * 2, (cuuint64_t)a->x.w * a->x.cs * 2,
                             (cuuint64_t)a->x.h * a->x.w * a->x.cs * 2};
    cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(halo ? p.halo_w : p.tw * a->sw),
                         (cuuint32_t)(halo ? halo_h : p.th * a->sh), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)a->sw, (cuuint32_t)a->sh, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->x.ptr, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vps::set_error("conv2d_tc: encode A failed (%d) dims %d,%d,%d,%d cs %d box %d,%d,%d", (int)r, a->x.c,
                     a->x.w, a->x.h, a->x.n, a->x.cs, bk, p.tw * a->sw, p.th * a->sh);
      return VPS_E_CUDA;
    }
  }
  for (int i = 0; i < MAX_PROB; ++i) {
    const vps_conv_args* q = &args[i < nprob ? i : 0];
    // packed weights [cout_pad][tap][cin_pad] viewed as {cin_pad, cout_pad, taps}: a box is {bk, block_n, taps-per-slot},
    // i.e. consecutive K-major [block_n][bk] tiles, one per tap
    const cuuint64_t K = (cuuint64_t)a->kh * a->kw * cin_pad;
    cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)(a->kh * a->kw)};
    cuuint64_t strides[2] = {K * 2, (cuuint64_t)cin_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)block_n, (cuuint32_t)(p.rowg ? a->kw : 1)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)q->w, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vps::set_error("conv2d_tc: encode B failed (%d)", (int)r); return VPS_E_CUDA; }
  }
  const int smem = p.a_stages * p.a_stage_bytes + p.b_stages * b_stage_bytes + 1024 + 8 * (4 * MAX_STAGES + 8);
  static bool smem_set = false;
  if (!smem_set) {
    if (cudaFuncSetAttribute(conv_igemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess) {
      vps::set_error("conv2d_tc: cannot raise dynamic smem: %s", cudaGetErrorString(cudaGetLastError()));
      return VPS_E_CUDA;
    }
    smem_set = true;
  }
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  static int stats_env = -1;
  static long long* stats_buf = nullptr;
  if (stats_env < 0) { const char* e = getenv("VPS_CONV_STATS"); stats_env = e ? atoi(e) : 0; }
  p.stats = nullptr;
  if (stats_env) {   // debugging aid: per-role barrier-wait clocks, printed after a device sync (never on in production)
    if (!stats_buf) cudaMalloc(&stats_buf, sizeof(long long) * 8 * 1024);
    cudaMemsetAsync(stats_buf, 0, sizeof(long long) * 8 * grid, (cudaStream_t)stream);
    p.stats = stats_buf;
  }
  static int pdl_env = -1;
  if (pdl_env < 0) { const char* e = getenv("VPS_PDL"); pdl_env = e ? atoi(e) : 1; }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_env ? 1 : 0;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, conv_igemm_tc_kernel, tmA, tmB[0], tmB[1], tmB[2], tmB[3], p);
    if (le != cudaSuccess) { vps::set_error("conv2d_tc: launch failed: %s", cudaGetErrorString(le)); return VPS_E_CUDA; }
  }
  VPS_CUDA_LAST("conv_igemm_tc_kernel");
  if (stats_env) {
    static long long h[8 * 1024];
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaMemcpy(h, stats_buf, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost);
    double m[8] = {0};
    for (int i = 0; i < grid; ++i) for (int j = 0; j < 8; ++j) m[j] += (double)h[i * 8 + j] / grid;
    const int tiles_cta = (p.total_tiles + grid - 1) / grid;
    fprintf(stderr, "conv_tc stats %dx%d %d->%d @%dx%d halo=%d/%d g%d bn=%d bk=%d stages a%d b%d tiles/cta %d steps/tile %d | clk/CTA: total %.0f  "
            "prod wait Aempty %.0f Bempty %.0f | mma wait Afull %.0f Bfull %.0f tmem-empty %.0f | epi wait tfull %.0f work %.0f\n",
            a->kh, a->kw, a->cin, a->cout, a->oh, a->ow, p.halo, p.rowg, p.gsub, block_n, bk, p.a_stages, p.b_stages, tiles_cta,
            a->kh * a->kw * p.cin_chunks, m[5], m[0], m[1], m[2], m[3], m[4], m[6], m[7]);
  }
  return VPS_OK;
}

extern "C" int vps_conv2d_tc(const vps_conv_args* a, void* stream) { return vps_conv2d_tc_multi(a, 1, stream); }


// Fused DCNv1 3x3 / stride 1 / pad 1 / dilation 1 / 1 deformable group (deform_conv.py:15-87 forward):
// x bf16 NHWC, offset f32 NHWC [.., 18] = (dy, dx) per tap, w = vps_pack_weights_tc layout of the [cout, cin, 3, 3]
// kernel (cin % 64 == 0, cout <= 256), y bf16 / f32 NHWC.  No bias (the reference's DeformConv has none).
extern "C" int vps_deform_conv_tc(const vps_tensor* x, const vps_tensor* offset, const void* w, int cout, const vps_tensor* y,
                                  void* stream) {
  VPS_CHECK_ARG(x->dtype == VPS_BF16 && offset->dtype == VPS_F32 && offset->c >= 18, "deform_conv_tc: dtypes");
  VPS_CHECK_ARG(x->c % 64 == 0 && x->cs % 8 == 0 && ((uintptr_t)x->ptr & 15) == 0, "deform_conv_tc: x must have cin %% 64 == 0");
  VPS_CHECK_ARG(offset->n == x->n && offset->h == x->h && offset->w == x->w && y->n == x->n && y->h == x->h && y->w == x->w &&
                    y->c == cout, "deform_conv_tc: shapes");
  VPS_CHECK_ARG((int64_t)x->n * x->h * x->w * x->cs < (1ll << 31), "deform_conv_tc: tensor too large for 32-bit offsets");
  const int cout_pad = (cout + 15) / 16 * 16;
  VPS_CHECK_ARG(cout_pad <= 256 && ((uintptr_t)w & 15) == 0, "deform_conv_tc: cout %d > 256", cout);
  auto encode = get_encode();
  if (!encode) { vps::set_error("cuTensorMapEncodeTiled unavailable"); return VPS_E_CUDA; }
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) { vps::set_error("no device"); return VPS_E_NODEV; }
  }
  ConvTcParams p = {};
  p.bk = 64; p.nprob = 1;
  p.n_img = x->n; p.oh = x->h; p.ow = x->w;
  int best_tw = 16; int64_t best_area = -1;
  const int cands[5] = {16, 8, 32, 64, 128};
  for (int i = 0; i < 5; ++i) {
    const int tw = cands[i], th = 128 / tw;
    const int64_t area = (int64_t)vps::cdiv(x->w, tw) * tw * vps::cdiv(x->h, th) * th;
    if (best_area < 0 || area < best_area) { best_area = area; best_tw = tw; }
  }
  p.tw = best_tw; p.th = 128 / best_tw;
  p.tiles_x = vps::cdiv(x->w, p.tw); p.tiles_y = vps::cdiv(x->h, p.th);
  p.n_tiles_n = 1; p.block_n = cout_pad;
  p.kh = p.kw = 3; p.sh = p.sw = 1; p.ph = p.pw = 1;
  p.cin_chunks = x->c / 64;
  p.gsub = 1; p.nk_last = 4; p.halo = 0; p.rowg = 0;
  p.a_box_bytes = BLOCK_M * 128; p.a_stage_bytes = p.a_box_bytes;
  const int stage_bytes = p.a_stage_bytes + cout_pad * 128;
  int stages = (200 * 1024 - DCN_SETUP_BYTES) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  VPS_CHECK_ARG(stages >= 2, "deform_conv_tc: ring does not fit");
  p.a_stages = p.b_stages = stages;
  p.tiles_per_prob = p.n_img * p.tiles_y * p.tiles_x;
  p.total_tiles = p.tiles_per_prob;
  p.y = y->ptr; p.y_h = y->h; p.y_w = y->w; p.y_cs = y->cs; p.y_dtype = y->dtype;
  const int esz = y->dtype == VPS_BF16 ? 2 : 4;
  p.y_vec = (((uintptr_t)y->ptr & 15) == 0) && ((y->cs * esz) % 16 == 0);
  if (p.y_vec && (((uintptr_t)y->ptr & 31) == 0) && ((y->cs * esz) % 32 == 0)) p.y_vec = 2;
  p.oy_mul = p.ox_mul = 1;
  p.res = nullptr; p.bias = nullptr; p.cout = cout; p.act = VPS_ACT_NONE; p.slope = 0.f; p.out_scale = 1.f;
  p.stats = nullptr;
  if (p.total_tiles == 0) return VPS_OK;
  DcnParams d;
  d.x = (const __nv_bfloat16*)x->ptr; d.off = (const float*)offset->ptr; d.x_cs = x->cs; d.off_cs = offset->cs;
  d.H = x->h; d.W = x->w;
  CUtensorMap tmB;
  {
    const cuuint64_t K = (cuuint64_t)9 * x->c;
    cuuint64_t dims[3] = {(cuuint64_t)x->c, (cuuint64_t)cout_pad, 9};
    cuuint64_t strides[2] = {K * 2, (cuuint64_t)x->c * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)cout_pad, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vps::set_error("deform_conv_tc: encode B failed (%d)", (int)r); return VPS_E_CUDA; }
  }
  const int smem = stages * stage_bytes + DCN_SETUP_BYTES + 1024 + 8 * (4 * MAX_STAGES + 8);
  static bool smem_set = false;
  if (!smem_set) {
    if (cudaFuncSetAttribute(dcn_igemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      vps::set_error("deform_conv_tc: cannot raise dynamic smem: %s", cudaGetErrorString(cudaGetLastError()));
      return VPS_E_CUDA;
    }
    smem_set = true;
  }
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(DCN_THREADS); cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, dcn_igemm_tc_kernel, tmB, p, d);
  if (le != cudaSuccess) { vps::set_error("deform_conv_tc: launch failed: %s", cudaGetErrorString(le)); return VPS_E_CUDA; }
  VPS_CUDA_LAST("dcn_igemm_tc_kernel");
  return VPS_OK;
}
