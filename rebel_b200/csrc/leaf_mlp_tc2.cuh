// Two-tiles-in-flight variant of the tcgen05 value-net kernel (leaf_mlp_tc.cuh): same math, same K-step order, bit-identical
// outputs — but the tensor pipe and the SIMT epilogue no longer wait for each other.
//
// In the one-tile kernel the epilogue warps idle while the last quarter of a layer's MMAs drains (clock64 traces: 840 + 500 of
// 11 000 cycles per tile, plus 500 for the output) and the tensor pipe idles through both epilogues.  TMEM (512 columns) cannot
// hold two tiles' accumulators (2 x 256) AND their A operands, so here the activations go back to the tensor cores through
// SHARED memory instead: a two-slot ring of 16 KB sub-chunks (128 rows x 64 K-values in UMMA core-matrix order) that the
// epilogue fills with 16-byte stores and the issuer drains with SS MMAs.  TMEM then holds exactly two accumulator regions,
// tile j uses region j & 1, and the epilogue warps run the pair schedule
//     E1(a) E1(b) E2(a) E2(b) out(a) out(b)
// so every wait of a tile is covered by an epilogue of the other tile.  The issuer thread follows the same order:
//     L1(b) | L2(a) L2(b) L3(a) L3(b) | L1(a')
// (L1 of a tile needs its region's previous output read: `outdone`).  Layer-3 accumulators (16 columns) reuse the tile's own
// region once the layer-2 accumulators are in registers.  Needs 32 KB more shared memory than the one-tile kernel, so it only
// fits when the padded query width is 32 (1x4f, 1x5f, 1x6f).
//
// MEASURED RESULT (B200, 1x6f, 540 672 rows): bit-identical to the one-tile kernel for every wave size, but 188 us per launch
// against 166 us.  The waits do shrink (ncu: 12 % of the samples in mbarrier / named-barrier waits instead of 17 %), but every
// epilogue gets ~25 % slower: SS MMAs fetch A (4 KB) and B (8 KB) per K-step from shared memory, 96 of the SM's 128 B/clk, and
// now run concurrently with the epilogues all the time, whose LayerNorm-parameter loads, A-ring stores and mbarrier traffic
// queue behind them.  The kernel is therefore OPT-IN (CFRB_TC2=1) and kept as the starting point for the cta_group::2 version,
// where each SM of a pair fetches only half of B.
#pragma once
#include "leaf_mlp_tc.cuh"

namespace cfrb {
namespace tc {

constexpr int kRingSlotBytes = kTileM * 64 * 2;      // one sub-chunk: 128 rows x (4 quarters x 16 features) fp16
struct Tc2Layout {
  BlobLayout L;
  int off_ring, smem_bytes;
  __host__ __device__ explicit Tc2Layout(int kp) : L(kp) {
    off_ring = (L.smem_bytes + 1023) / 1024 * 1024;
    smem_bytes = off_ring + 2 * kRingSlotBytes;
  }
};

// LayerNorm + GELU of this thread's quarter row, read from TMEM region `dcol`, handed to the issuer through the smem ring.
// `chunk` counts the sub-chunks produced so far by this CTA (uniform over the epilogue threads).
template <bool kGeluX2>
__device__ __forceinline__ void epilogue2_ln_gelu(uint32_t tmem_row, uint32_t dcol, int part_id, int row, const float2* __restrict__ ln,
                                                  float2* part, uint8_t* ring, uint32_t bar_afull, uint32_t bar_ringfree, uint32_t& chunk,
                                                  uint32_t bar_x_pending) {
  uint32_t xr[kColsPerThread];
  {
    uint32_t* lo = xr; uint32_t* hi = xr + 32;
    CFRB_TMEM_LD32(tmem_row + dcol + part_id * kColsPerThread, lo);
    CFRB_TMEM_LD32(tmem_row + dcol + part_id * kColsPerThread + 32, hi);
    tmem_wait_ld();
  }
  float sum, sumsq;
  {
    f32x2 s2[2], q2[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) { s2[u] = pack2(0.f, 0.f); q2[u] = pack2(0.f, 0.f); }
#pragma unroll
    for (int i = 0; i < kColsPerThread; i += 2) {
      const f32x2 x2 = pack2(__uint_as_float(xr[i]), __uint_as_float(xr[i + 1]));
      s2[(i >> 1) & 1] = add2(s2[(i >> 1) & 1], x2);
      q2[(i >> 1) & 1] = fma2(x2, x2, q2[(i >> 1) & 1]);
    }
    const f32x2 st = add2(s2[0], s2[1]), qt = add2(q2[0], q2[1]);
    float a0, a1, b0, b1;
    unpack2(st, a0, a1); unpack2(qt, b0, b1);
    sum = a0 + a1; sumsq = b0 + b1;
  }
  part[part_id * kTileM + row] = make_float2(sum, sumsq);
  tc_fence_before();                           // this thread's TMEM reads are complete before anyone overwrites the region
  named_bar_sync(1, kEpiThreads);
  if (bar_x_pending) {                         // the next query tile (cp.async issued before the TMEM load) has landed by now
    asm volatile("cp.async.wait_all;" ::: "memory");
    fence_proxy_async_smem();
    mbar_arrive(bar_x_pending);
  }
  sum = 0.f; sumsq = 0.f;
#pragma unroll
  for (int p = 0; p < kParts; ++p) {
    const float2 o = part[p * kTileM + row];
    sum += o.x; sumsq += o.y;
  }
  const float mean = sum * (1.f / kHid);
  const float var = fmaxf(sumsq * (1.f / kHid) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float shift = -mean * rstd;
  const f32x2 rstd2 = pack2(rstd, rstd), shift2 = pack2(shift, shift);
  // core-matrix position of this thread's 16 features inside a ring slot: K-step = its column quarter, two 8-wide K columns
  const int row_off = (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
  for (int c = 0; c < kSubChunks; ++c, ++chunk) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = c * 16 + 2 * i;
      const float4 pp = *reinterpret_cast<const float4*>(ln + part_id * kColsPerThread + j);
      float y0, y1;
      unpack2(fma2(fma2(pack2(__uint_as_float(xr[j]), __uint_as_float(xr[j + 1])), rstd2, shift2), pack2(pp.x, pp.y), pack2(pp.z, pp.w)), y0, y1);
      if (kGeluX2) {
        pk[i] = gelu_tc_x2(y0, y1);
      } else {
        const __half2 h = __floats2half2_rn(gelu_tc(y0), gelu_tc(y1));
        pk[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    const uint32_t slot = chunk & 1u, use = chunk >> 1;
    if (use > 0) mbar_wait(bar_ringfree + 8 * slot, (use - 1) & 1u);     // the MMAs that read the slot's previous content are done
    uint8_t* dst = ring + slot * kRingSlotBytes + part_id * 4096 + row_off;
    *reinterpret_cast<int4*>(dst) = make_int4((int)pk[0], (int)pk[1], (int)pk[2], (int)pk[3]);
    *reinterpret_cast<int4*>(dst + 2048) = make_int4((int)pk[4], (int)pk[5], (int)pk[6], (int)pk[7]);
    fence_proxy_async_smem();
    mbar_arrive(bar_afull + 8 * slot);
  }
}

template <bool kGeluX2>
__global__ void __launch_bounds__(kThreads, 1) leaf_mlp_tc2_kernel(TcArgs a) {
  constexpr int kMmaWarp = kEpiThreads / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  const Tc2Layout T(a.Kp);
  const BlobLayout& L = T.L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = *a.rows_ptr;
  const int ntiles = (rows + kTileM - 1) / kTileM;
  if ((int)blockIdx.x >= ntiles) return;
  const int J = (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;      // tiles of this CTA: blockIdx.x + j * gridDim.x

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  // 0 x | 1,2 d1[region] | 3,4 d2 | 5,6 d3 | 7,8 outdone | 9,10 afull[slot] | 11,12 ringfree[slot]
  const uint32_t bar_x = smem_u32(bars + 0), bar_d1 = smem_u32(bars + 1), bar_d2 = smem_u32(bars + 3), bar_d3 = smem_u32(bars + 5),
                 bar_out = smem_u32(bars + 7), bar_afull = smem_u32(bars + 9), bar_ringfree = smem_u32(bars + 11);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  {
    const int4* src = reinterpret_cast<const int4*>(a.blob);
    int4* dst = reinterpret_cast<int4*>(smem);
    for (int i = tid; i < L.blob_bytes / 16; i += kThreads) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    mbar_init(bar_x, kEpiThreads);
    for (int r = 0; r < 2; ++r) {
      mbar_init(bar_d1 + 8 * r, 1); mbar_init(bar_d2 + 8 * r, 1); mbar_init(bar_d3 + 8 * r, 1);
      mbar_init(bar_out + 8 * r, kTileM);
      mbar_init(bar_afull + 8 * r, kEpiThreads); mbar_init(bar_ringfree + 8 * r, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sx = smem_u32(smem + L.off_x), sw1 = smem_u32(smem + L.off_w1), sw2 = smem_u32(smem + L.off_w2),
                 sw3 = smem_u32(smem + L.off_w3), sones = smem_u32(smem + L.off_ones), sbias2 = smem_u32(smem + L.off_bias2),
                 sring = smem_u32(smem + T.off_ring);
  const int x_tile_int4 = kTileM * a.Kp * 2 / 16;
  const int npairs = (J + 1) / 2;

  if (warp == kMmaWarp) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc256 = make_idesc(kTileM, kHid), idesc16 = make_idesc(kTileM, kNout);
      const uint32_t lbo_x = (kTileM / 8) * 128, lbo_w = (kHid / 8) * 128, lbo_w3 = (kNout / 8) * 128;
      uint32_t chunk = 0;          // sub-chunks consumed, same order as the epilogue produces them
      uint32_t xphase = 0;         // query tiles consumed
      auto layer1 = [&](int region) {
        mbar_wait(bar_x, xphase & 1u); ++xphase;
        tc_fence_after();
        for (int k = 0; k < a.Kp / 16; ++k)
          mma_ss(tmem_base + region * 256, make_desc(sx + k * 2 * lbo_x, lbo_x, 128), make_desc(sw1 + k * 2 * lbo_w, lbo_w, 128), idesc256, k > 0);
        tc_commit(bar_d1 + 8 * region);
      };
      auto layer23 = [&](int region, bool third) {
        const uint32_t d = tmem_base + region * 256;          // layer 3 accumulates into the first 16 columns of the same region
        for (int c = 0; c < kSubChunks; ++c, ++chunk) {
          const uint32_t slot = chunk & 1u, use = chunk >> 1;
          mbar_wait(bar_afull + 8 * slot, use & 1u);
          tc_fence_after();
          if (!third && c == 0) mma_ss(d, make_desc(sones, lbo_x, 128), make_desc(sbias2, lbo_w, 128), idesc256, 0);
          for (int q = 0; q < kParts; ++q) {
            const int k = q * kSubChunks + c;
            const uint64_t adesc = make_desc(sring + slot * kRingSlotBytes + q * 4096, lbo_x, 128);
            if (third) mma_ss(d, adesc, make_desc(sw3 + k * 2 * lbo_w3, lbo_w3, 128), idesc16, (c | q) != 0);
            else mma_ss(d, adesc, make_desc(sw2 + k * 2 * lbo_w, lbo_w, 128), idesc256, 1);
          }
          tc_commit(bar_ringfree + 8 * slot);
        }
        tc_commit((third ? bar_d3 : bar_d2) + 8 * region);
      };
      layer1(0);
      for (int m = 0; m < npairs; ++m) {
        const bool has_b = 2 * m + 1 < J, has_next = 2 * m + 2 < J;
        const uint32_t pp = m & 1u;
        if (has_b) {
          if (m > 0) { mbar_wait(bar_out + 8, (m - 1) & 1u); tc_fence_after(); }     // region 1's previous output has been read
          layer1(1);
        }
        layer23(0, false);
        if (has_b) layer23(1, false);
        layer23(0, true);
        if (has_b) layer23(1, true);
        if (has_next) {
          mbar_wait(bar_out, pp);
          tc_fence_after();
          layer1(0);
        }
      }
    }
  } else {
    // ===================== epilogue warps: thread == (row, column quarter) =====================
    const int quad = warp & 3, part_id = warp >> 2;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float2* ln1 = reinterpret_cast<const float2*>(smem + L.off_ln1);
    const float2* ln2 = reinterpret_cast<const float2*>(smem + L.off_ln2);
    const float* b3 = reinterpret_cast<const float*>(smem + L.off_b3);
    float2* part1 = reinterpret_cast<float2*>(smem + L.off_part);
    float2* part2 = part1 + kParts * kTileM;
    uint8_t* ring = smem + T.off_ring;
    int4* xdst = reinterpret_cast<int4*>(smem + L.off_x);
    const int x_items = x_tile_int4;                               // <= 2 int4 per thread
    const int4* xbase = reinterpret_cast<const int4*>(a.Xh);
    auto tile_of = [&](int j) { return (int)blockIdx.x + j * (int)gridDim.x; };
    {
      const int4* xsrc = xbase + (size_t)tile_of(0) * x_tile_int4;
      for (int i = tid; i < x_items; i += kEpiThreads) xdst[i] = __ldg(xsrc + i);
      fence_proxy_async_smem();
      mbar_arrive(bar_x);
    }
    uint32_t chunk = 0;
    uint32_t nepi = 0;             // epilogues run so far: consecutive ones alternate between the two LayerNorm exchange buffers, so
                                   // a buffer is rewritten only after every thread passed the barrier of the epilogue in between
    auto output = [&](int j, int region, uint32_t pp) {
      if (part_id != 0) return;
      mbar_wait(bar_d3 + 8 * region, pp);
      tc_fence_after();
      uint32_t v[16];
      CFRB_TMEM_LD16(tmem_row + region * 256, v);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_out + 8 * region);                           // the region may take the next layer-1 accumulators
      const int row = tile_of(j) * kTileM + row_in_tile;
      if (row < rows) {
        float* o = a.out + (size_t)row * a.Hout;
#pragma unroll
        for (int h = 0; h < kNout; ++h) if (h < a.H) o[h] = __uint_as_float(v[h]) + b3[h];
      }
    };
    for (int m = 0; m < npairs; ++m) {
      const int ja = 2 * m, jb = 2 * m + 1;
      const bool has_b = jb < J;
      const uint32_t pp = m & 1u;
      // E1(a) E1(b) E2(a) E2(b): one (not unrolled) loop, so the epilogue body exists once in the kernel
#pragma unroll 1
      for (int ph = 0; ph < 4; ++ph) {
        const int region = ph & 1, layer = ph >> 1;
        if (region && !has_b) continue;
        const int j = ja + region;
        mbar_wait((layer ? bar_d2 : bar_d1) + 8 * region, pp);
        tc_fence_after();
        uint32_t x_pending = 0;
        if (layer == 0 && j + 1 < J) {   // layer 1 of tile j has read the query tile: copy the next one in asynchronously
          const int4* xn = xbase + (size_t)tile_of(j + 1) * x_tile_int4;
          for (int i = tid; i < x_items; i += kEpiThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(xdst + i)), "l"(xn + i) : "memory");
          x_pending = bar_x;
        }
        epilogue2_ln_gelu<kGeluX2>(tmem_row, region * 256, part_id, row_in_tile, layer ? ln2 : ln1, (nepi++ & 1u) ? part2 : part1, ring,
                                   bar_afull, bar_ringfree, chunk, x_pending);
      }
      output(ja, 0, pp);
      if (has_b) output(jb, 1, pp);
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace tc
}  // namespace cfrb
