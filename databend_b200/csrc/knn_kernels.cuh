// knn_kernels.cuh — vector distance kernels for sm_100a.
//
//  (1) exact row-wise cosine / L2 in f32 with the reference's evaluation order
//      (src/common/vector/src/distance.rs:19-35,65-80; ndarray 0.15.6 unrolled_fold for the
//      cosine sums) — the ScalarFunction::eval replacement and the re-rank of kNN candidates;
//  (2) the batched query x corpus similarity GEMM on the 5th-generation tensor cores:
//      TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma (bf16 in, f32
//      accumulators in TMEM) -> tcgen05.ld epilogue that turns dot products into similarities and
//      keeps only entries that beat the per-query boundary (the k'-th best so far), i.e. the
//      score matrix is never written to HBM.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace dbx {

// ---------------------------------------------------------------- exact f32 distances
// ndarray's unrolled_fold: eight interleaved accumulators over full chunks of 8, combined as
// ((((0 + (p0+p4)) + (p1+p5)) + (p2+p6)) + (p3+p7)), then the < 8 tail sequentially.  Products
// are rounded to f32 before the add (`&a * &b` materialises an f32 array): no FMA contraction.
struct CosAcc {
  float aa[8], bb[8], ab[8];
};
__device__ __forceinline__ void cos_acc_init(CosAcc& c) {
#pragma unroll
  for (int j = 0; j < 8; ++j) c.aa[j] = c.bb[j] = c.ab[j] = 0.0f;
}
__device__ __forceinline__ float fold8(const float (&p)[8]) {
  float acc = 0.0f;
  acc = __fadd_rn(acc, __fadd_rn(p[0], p[4]));
  acc = __fadd_rn(acc, __fadd_rn(p[1], p[5]));
  acc = __fadd_rn(acc, __fadd_rn(p[2], p[6]));
  acc = __fadd_rn(acc, __fadd_rn(p[3], p[7]));
  return acc;
}
__device__ __forceinline__ float exact_cosine(const float* __restrict__ a, const float* __restrict__ b, int dim) {
  CosAcc c;
  cos_acc_init(c);
  int i = 0;
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  for (; i + 8 <= dim; i += 8) {
    float x[8], y[8];
    if (vec) {
      float4 x0 = *reinterpret_cast<const float4*>(a + i), x1 = *reinterpret_cast<const float4*>(a + i + 4);
      float4 y0 = *reinterpret_cast<const float4*>(b + i), y1 = *reinterpret_cast<const float4*>(b + i + 4);
      x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
      y[0] = y0.x; y[1] = y0.y; y[2] = y0.z; y[3] = y0.w; y[4] = y1.x; y[5] = y1.y; y[6] = y1.z; y[7] = y1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { x[j] = a[i + j]; y[j] = b[i + j]; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c.aa[j] = __fadd_rn(c.aa[j], __fmul_rn(x[j], x[j]));
      c.bb[j] = __fadd_rn(c.bb[j], __fmul_rn(y[j], y[j]));
      c.ab[j] = __fadd_rn(c.ab[j], __fmul_rn(x[j], y[j]));
    }
  }
  float aa = fold8(c.aa), bb = fold8(c.bb), ab = fold8(c.ab);
  for (; i < dim; ++i) {
    aa = __fadd_rn(aa, __fmul_rn(a[i], a[i]));
    bb = __fadd_rn(bb, __fmul_rn(b[i], b[i]));
    ab = __fadd_rn(ab, __fmul_rn(a[i], b[i]));
  }
  // 1 - ab / (sqrt(aa) * sqrt(bb))
  return __fsub_rn(1.0f, __fdiv_rn(ab, __fmul_rn(__fsqrt_rn(aa), __fsqrt_rn(bb))));
}
// l2_distance: strictly sequential f32 fold of (a-b)^2, then sqrt
__device__ __forceinline__ float exact_l2(const float* __restrict__ a, const float* __restrict__ b, int dim) {
  float acc = 0.0f;
  for (int i = 0; i < dim; ++i) {
    float d = __fsub_rn(a[i], b[i]);
    acc = __fadd_rn(acc, __fmul_rn(d, d));
  }
  return __fsqrt_rn(acc);
}
__device__ __forceinline__ float exact_distance(int kind, const float* a, const float* b, int dim) {
  return kind == DBX_DIST_COSINE ? exact_cosine(a, b, dim) : exact_l2(a, b, dim);
}

// Coalesced form of exact_cosine: the 8 lanes of a group own the 8 interleaved accumulators of
// ONE row (lane j accumulates elements 8c+j, c ascending — the same chains in the same order),
// so a group reads one 32-byte sector per step; the fold and the tail are then done redundantly
// by every lane of the group.  All 32 lanes of the warp must call this together.
__device__ __forceinline__ float exact_cosine_g8(const float* __restrict__ a, const float* __restrict__ b, int dim,
                                                 int lane) {
  const int sub = lane & 7, gbase = lane & 24;
  float aa = 0.0f, bb = 0.0f, ab = 0.0f;
  const int n_chunks = dim >> 3;
  int c = 0;
  for (; c + 8 <= n_chunks; c += 8) {
    float x[8], y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { x[u] = __ldg(a + (c + u) * 8 + sub); y[u] = __ldg(b + (c + u) * 8 + sub); }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      aa = __fadd_rn(aa, __fmul_rn(x[u], x[u]));
      bb = __fadd_rn(bb, __fmul_rn(y[u], y[u]));
      ab = __fadd_rn(ab, __fmul_rn(x[u], y[u]));
    }
  }
  for (; c < n_chunks; ++c) {
    const float x = __ldg(a + c * 8 + sub), y = __ldg(b + c * 8 + sub);
    aa = __fadd_rn(aa, __fmul_rn(x, x));
    bb = __fadd_rn(bb, __fmul_rn(y, y));
    ab = __fadd_rn(ab, __fmul_rn(x, y));
  }
  float paa[8], pbb[8], pab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    paa[j] = __shfl_sync(0xffffffffu, aa, gbase + j);
    pbb[j] = __shfl_sync(0xffffffffu, bb, gbase + j);
    pab[j] = __shfl_sync(0xffffffffu, ab, gbase + j);
  }
  float saa = fold8(paa), sbb = fold8(pbb), sab = fold8(pab);
  for (int i = n_chunks * 8; i < dim; ++i) {
    const float x = __ldg(a + i), y = __ldg(b + i);
    saa = __fadd_rn(saa, __fmul_rn(x, x));
    sbb = __fadd_rn(sbb, __fmul_rn(y, y));
    sab = __fadd_rn(sab, __fmul_rn(x, y));
  }
  return __fsub_rn(1.0f, __fdiv_rn(sab, __fmul_rn(__fsqrt_rn(saa), __fsqrt_rn(sbb))));
}

// calculate_distance (scalars/vector.rs:497-556), either side may be const.
// cosine: 8 lanes per row (coalesced sectors); L2 (one strictly sequential chain): one thread per row.
__global__ void distance_rows_kernel(int kind, const float* lhs, int lhs_const, const float* rhs, int rhs_const,
                                     int64_t rows, int dim, const uint8_t* lv, int64_t lv_off, const uint8_t* rv,
                                     int64_t rv_off, float* out, uint8_t* out_valid_bytes) {
  if (kind == DBX_DIST_COSINE) {
    const int lane = threadIdx.x & 31, g = lane >> 3;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * 4; base < rows; base += n_warps * 4) {
      const int64_t r = base + g;
      const bool in = r < rows;
      const int64_t rr = in ? r : rows - 1;
      const float d = exact_cosine_g8(lhs + (lhs_const ? 0 : rr * dim), rhs + (rhs_const ? 0 : rr * dim), dim, lane);
      if (in && (lane & 7) == 0) {
        const bool ok = (!lv || bit_test(lv, lv_off + r)) && (!rv || bit_test(rv, rv_off + r));
        out[r] = ok ? d : 0.0f;
        if (out_valid_bytes) out_valid_bytes[r] = ok ? 1 : 0;
      }
    }
    return;
  }
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    bool ok = (!lv || bit_test(lv, lv_off + r)) && (!rv || bit_test(rv, rv_off + r));
    const float* a = lhs + (lhs_const ? 0 : r * dim);
    const float* b = rhs + (rhs_const ? 0 : r * dim);
    out[r] = ok ? exact_l2(a, b, dim) : 0.0f;
    if (out_valid_bytes) out_valid_bytes[r] = ok ? 1 : 0;
  }
}

// ---------------------------------------------------------------- corpus / query preparation
// f32 rows -> bf16 operand of the similarity GEMM (round to nearest even):
//   cosine: the row is normalised first (x / |x|), so the GEMM yields cosine similarities directly
//           and the epilogue is a bare compare; scale[r] = 1/|x| (kept for inspection);
//   L2:     the row is copied as is; scale[r] = |x|^2, combined in the epilogue.
// The sums here only steer candidate selection (returned distances are recomputed exactly).
// A zero vector becomes a NaN operand row: its similarities are NaN and never pass the filter.
// max_norm_bits (optional): bit pattern of the largest row norm (non-negative floats order like
// their bit patterns), the corpus-side constant of the L2 certificate.
__global__ void prep_rows_kernel(const float* src, int64_t rows, int dim, int dim_pad, __nv_bfloat16* dst, float* scale,
                                 int kind, unsigned int* max_norm_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float wmax = 0.0f;
  for (int64_t r = warp; r < rows; r += n_warps) {
    const float* a = src + r * dim;
    __nv_bfloat16* d = dst + r * dim_pad;
    float s = 0.0f;
    for (int i = lane; i < dim; i += 32) { const float x = a[i]; s += x * x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mul = kind == DBX_DIST_COSINE ? rsqrtf(s) : 1.0f;
    for (int i = lane * 2; i < dim_pad; i += 64) {  // dim_pad is a multiple of 64
      const float x0 = i < dim ? a[i] * mul : 0.0f, x1 = i + 1 < dim ? a[i + 1] * mul : 0.0f;
      *reinterpret_cast<__nv_bfloat162*>(d + i) = __floats2bfloat162_rn(x0, x1);
    }
    if (lane == 0) scale[r] = kind == DBX_DIST_COSINE ? mul : s;
    if (s == s) wmax = fmaxf(wmax, sqrtf(s));
  }
  if (max_norm_bits && lane == 0 && wmax > 0.0f) atomicMax(max_norm_bits, __float_as_uint(wmax));
}

// ---------------------------------------------------------------- tcgen05 GEMM with fused filter
constexpr int kGemmBM = 128;      // queries per tile      (UMMA M)
constexpr int kGemmBN = 256;      // corpus rows per tile  (UMMA N)
constexpr int kGemmBK = 64;       // bf16 per k-block = 128 bytes = one swizzle row
constexpr int kGemmStages = 4;
constexpr int kGemmThreads = 192; // warp 0: TMA, warp 1: MMA (+TMEM alloc), warps 2-5: epilogue
constexpr int kUmmaK = 16;
constexpr uint32_t kTmemCols = 512;  // two 256-column accumulators
constexpr uint32_t kStageBytesA = kGemmBM * kGemmBK * 2;
constexpr uint32_t kStageBytesB = kGemmBN * kGemmBK * 2;
constexpr int kCandStage = 512;   // staged survivors per epilogue warp

struct GemmSmem {
  alignas(1024) uint8_t a[kGemmStages][kStageBytesA];
  alignas(1024) uint8_t b[kGemmStages][kStageBytesB];
  alignas(8) uint64_t full_bar[kGemmStages];
  uint64_t empty_bar[kGemmStages];
  uint64_t tmem_full_bar[2];
  uint64_t tmem_empty_bar[2];
  uint32_t tmem_base;
  // per epilogue warp: survivors are staged here and written out in coalesced bursts, one
  // reservation (atomic on the global candidate counter) per burst instead of one per survivor
  alignas(16) uint64_t stage_key[4][kCandStage];
  uint32_t stage_row[4][kCandStage];
};

struct KnnGemmParams {
  int32_t kind;
  int32_t nq;            // valid queries
  int32_t nq_pad;        // multiple of kGemmBM * cluster size
  int32_t dim_pad;       // multiple of kGemmBK
  int64_t n0;            // first corpus row of this pass
  int64_t n_rows;        // corpus rows in this pass
  const float* q_scale;  // L2: |q|^2                (cosine: unused, operands are pre-normalised)
  const float* c_scale;  // L2: |c|^2 by global row  (cosine: unused)
  const float* bound;    // per query: only score >= bound can still reach the top k'
  uint64_t* cand_key;    // (query << 32) | ~ordered(score): ascending sort = best first
  uint32_t* cand_row;    // corpus row
  unsigned long long* cand_count;
  int64_t cand_cap;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint64_t* bar, void* dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Same box delivered to the same shared-memory offset of every CTA in `mask`; each destination
// CTA's mbarrier (same offset) receives the complete_tx for the bytes that landed in ITS smem.
__device__ __forceinline__ void tma_load_2d_multicast(const void* tmap, uint64_t* bar, void* dst, int32_t c0, int32_t c1,
                                                      uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%4, %5}], [%2], %3;" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp layout):
// start>>4 | LBO(=1, ignored for swizzled K-major)<<16 | SBO(1024 B between 8-row groups)>>4 <<32 |
// version 1 <<46 | layout SWIZZLE_128B(2) <<61
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;               // c_format = F32
  d |= 1u << 7;               // a_format = BF16
  d |= 1u << 10;              // b_format = BF16
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive (once the MMAs issued so far retire) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t f32_to_ordered32(float f) {
  if (f != f) return 0u;  // NaN: worst similarity
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// One reservation + coalesced copy of a warp's staged survivors into the global candidate list.
__device__ __forceinline__ void flush_stage(const uint64_t* skey, const uint32_t* srow, int cnt, int lane, const KnnGemmParams& p) {
  if (cnt == 0) return;
  __syncwarp();
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(p.cand_count, (unsigned long long)cnt);
  base = __shfl_sync(0xffffffffu, base, 0);
  if ((int64_t)(base + cnt) <= p.cand_cap) {
    for (int i = lane; i < cnt; i += 32) {
      p.cand_key[base + i] = skey[i];
      p.cand_row[base + i] = srow[i];
    }
  }
  __syncwarp();
}

// Persistent, warp-specialised kernel.  A cluster of C CTAs works on C consecutive query blocks
// against the SAME 256-row corpus tile: every CTA streams its own query tile (A) and 1/C of the
// corpus tile (B), which TMA multicasts into the shared memory of all C CTAs — so per CTA the
// L2 -> SM traffic per k-block drops from 16+32 KB to 16+32/C KB (the 1-CTA kernel is bound by
// exactly that traffic: 96 B/clk/SM at full tensor rate).  Consecutive cluster tiles walk the
// query blocks first, so a corpus tile is fetched from HBM once and re-read from L2.
template <int C>
__global__ void __launch_bounds__(kGemmThreads, 1)
knn_gemm_filter_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                       const __grid_constant__ KnnGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = C > 1 ? cluster_ctarank() : 0u;
  const int64_t cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
  const int n_mgrp = p.nq_pad / (kGemmBM * C);  // groups of C query blocks
  const int64_t n_nblk = (p.n_rows + kGemmBN - 1) / kGemmBN;
  const int64_t n_tiles = n_nblk * n_mgrp;      // cluster-level tiles
  const int n_kblk = p.dim_pad / kGemmBK;
  constexpr uint16_t kMask = (uint16_t)((1u << C) - 1u);
  constexpr int kSliceRows = kGemmBN / C;       // corpus rows this CTA fetches for the whole cluster

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGemmStages; ++s) { mbar_init(&sm.full_bar[s], 1); mbar_init(&sm.empty_bar[s], C); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full_bar[a], 1); mbar_init(&sm.tmem_empty_bar[a], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
  }
  if (warp == 1) {  // one warp allocates TMEM and later frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (C > 1) cluster_sync_all(); else __syncthreads();  // peers' barriers are initialised before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
        const int m_blk = (int)(t % n_mgrp) * C + (int)cta_rank;
        const int64_t n_blk = t / n_mgrp;
        const int32_t row_b = (int32_t)(p.n0 + n_blk * kGemmBN) + (int32_t)cta_rank * kSliceRows;
        for (int kb = 0; kb < n_kblk; ++kb) {
          mbar_wait(&sm.empty_bar[stage], phase ^ 1);  // all C CTAs have consumed this slot
          mbar_expect_tx(&sm.full_bar[stage], kStageBytesA + kStageBytesB);
          tma_load_2d(&tmap_q, &sm.full_bar[stage], sm.a[stage], kb * kGemmBK, m_blk * kGemmBM);
          if (C > 1)
            tma_load_2d_multicast(&tmap_c, &sm.full_bar[stage], sm.b[stage] + cta_rank * (kSliceRows * kGemmBK * 2), kb * kGemmBK, row_b, kMask);
          else
            tma_load_2d(&tmap_c, &sm.full_bar[stage], sm.b[stage], kb * kGemmBK, row_b);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
      }
      if (C > 1) {  // tail: do not leave while peers may still signal this CTA's barriers
        for (int i = 0; i < kGemmStages; ++i) {
          mbar_wait(&sm.empty_bar[stage], phase ^ 1);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kGemmBM, kGemmBN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
        mbar_wait(&sm.tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * kGemmBN;
        for (int kb = 0; kb < n_kblk; ++kb) {
          mbar_wait(&sm.full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(sm.a[stage]), b_addr = smem_u32(sm.b[stage]);
#pragma unroll
          for (int k = 0; k < kGemmBK / kUmmaK; ++k) {
            const uint64_t adesc = make_smem_desc(a_addr + k * kUmmaK * 2);
            const uint64_t bdesc = make_smem_desc(b_addr + k * kUmmaK * 2);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage once these MMAs retire — in every CTA of the cluster, because
          // each of them writes a slice of the next fill into this CTA's slot
          if (C > 1) umma_commit_multicast(&sm.empty_bar[stage], kMask); else umma_commit(&sm.empty_bar[stage]);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&sm.tmem_full_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> score -> boundary filter -> candidate list =====
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool is_l2 = p.kind != DBX_DIST_COSINE;
    int stage_cnt = 0;  // warp-uniform
    for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
      const int m_blk = (int)(t % n_mgrp) * C + (int)cta_rank;
      const int64_t n_blk = t / n_mgrp;
      const int q = m_blk * kGemmBM + quarter * 32 + lane;
      const bool q_ok = q < p.nq;
      const float qq = (q_ok && is_l2) ? p.q_scale[q] : 0.0f;
      const float bound = q_ok ? p.bound[q] : __int_as_float(0x7f800000);  // +inf: nothing passes
      mbar_wait(&sm.tmem_full_bar[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row0 = p.n0 + n_blk * kGemmBN;
      const int64_t row_end = p.n0 + p.n_rows;
#pragma unroll 1
      for (int c = 0; c < kGemmBN / 32; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kGemmBN + c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        const int64_t rbase = row0 + c * 32;
        const int64_t left = row_end - rbase;  // rows of this chunk that exist
        const uint32_t col_mask = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
        float cc_lane = 0.0f;
        if (is_l2) cc_lane = (lane < left) ? __ldg(p.c_scale + rbase + lane) : 0.0f;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t pass = 0;
        if (!is_l2) {
          // cosine: operands are pre-normalised, the accumulator IS the similarity
#pragma unroll
          for (int j = 0; j < 32; ++j) pass |= (__uint_as_float(v[j]) >= bound) ? (1u << j) : 0u;
        } else {
          // L2: score = -(|q|^2 + |c|^2 - 2 q.c)   (larger = closer)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float cc = __shfl_sync(0xffffffffu, cc_lane, j);
            const float sc = 2.0f * __uint_as_float(v[j]) - qq - cc;
            v[j] = __float_as_uint(sc);
            pass |= (sc >= bound) ? (1u << j) : 0u;
          }
        }
        pass &= col_mask;
        if (__any_sync(0xffffffffu, pass != 0)) {
          const int n = __popc(pass);
          int incl = n;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          const int excl = incl - n;
          if (total > kCandStage / 2) {
            // dense chunk (loose boundary in the first passes): reserve once per warp, write direct
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(p.cand_count, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((int64_t)(base + total) <= p.cand_cap) {
              unsigned long long pos = base + excl;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if ((pass >> j) & 1) {
                  p.cand_key[pos] = ((uint64_t)(uint32_t)q << 32) | (uint64_t)(~f32_to_ordered32(__uint_as_float(v[j])));
                  p.cand_row[pos] = (uint32_t)(rbase + j);
                  ++pos;
                }
              }
            }
          } else {
            if (stage_cnt + total > kCandStage) {
              flush_stage(sm.stage_key[quarter], sm.stage_row[quarter], stage_cnt, lane, p);
              stage_cnt = 0;
            }
            int pos = stage_cnt + excl;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if ((pass >> j) & 1) {
                sm.stage_key[quarter][pos] = ((uint64_t)(uint32_t)q << 32) | (uint64_t)(~f32_to_ordered32(__uint_as_float(v[j])));
                sm.stage_row[quarter][pos] = (uint32_t)(rbase + j);
                ++pos;
              }
            }
            stage_cnt += total;
            __syncwarp();
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&sm.tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    flush_stage(sm.stage_key[quarter], sm.stage_row[quarter], stage_cnt, lane, p);
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (C > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// Reference similarity pass on CUDA cores (same bf16 inputs, f32 accumulation, same filter):
// used by the tests to validate the tcgen05 path (env DBX_KNN_REF_GEMM=1), never by default.
__global__ void knn_ref_filter_kernel(const __nv_bfloat16* q, const __nv_bfloat16* c, const __grid_constant__ KnnGemmParams p) {
  const int64_t total = (int64_t)p.nq * p.n_rows;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int qi = (int)(i / p.n_rows);
    const int64_t r = p.n0 + i % p.n_rows;
    const __nv_bfloat16* a = q + (int64_t)qi * p.dim_pad;
    const __nv_bfloat16* b = c + r * p.dim_pad;
    float dot = 0.0f;
    for (int k = 0; k < p.dim_pad; ++k) dot += __bfloat162float(a[k]) * __bfloat162float(b[k]);
    const float s = p.kind == DBX_DIST_COSINE ? dot : 2.0f * dot - p.q_scale[qi] - p.c_scale[r];
    if (s >= p.bound[qi]) {
      unsigned long long pos = atomicAdd(p.cand_count, 1ULL);
      if ((int64_t)pos < p.cand_cap) {
        p.cand_key[pos] = ((uint64_t)(uint32_t)qi << 32) | (uint64_t)(~f32_to_ordered32(s));
        p.cand_row[pos] = (uint32_t)r;
      }
    }
  }
}

}  // namespace dbx
