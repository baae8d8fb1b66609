// radix_sort.cuh — hand-written onesweep-style LSD radix sort for (u64 key, u32 value) pairs.
// Used by DBX_OP_TOPK (full ORDER BY, final order of large top-k results) and by the kNN candidate
// cuts.  Included by the .cu files that sort (kernels live in an unnamed namespace).
#pragma once
#include <algorithm>
#include <atomic>

#include "runtime.h"

namespace dbx {
namespace {

// ================================================================ onesweep LSD radix sort
// keys: u64; values: u32 (row ids / permutation indices).  One pass per 8-bit digit:
//   tile = 4096 keys of one CTA, loaded warp-striped (warp w owns 512 consecutive keys, item i of
//   lane l at 32 i + l), ranked with __match_any_sync in (item, lane) order = memory order, so a
//   pass is STABLE; per-digit tile counts are chained to the previous tiles with decoupled
//   look-back (status word = 2-bit flag | 30-bit count, so the data travels with the flag); keys
//   and values are reordered through shared memory and written out in per-digit runs.
// Tiles take their index from an atomic ticket, so a tile only ever waits for tiles that already
// run; waits are bounded and set a fail flag instead of hanging.
namespace rs {
constexpr int kThreads = 256, kItems = 16, kTile = kThreads * kItems, kRadix = 256, kWarps = kThreads / 32;
constexpr uint32_t kFlagAgg = 1u << 30, kFlagPrefix = 2u << 30, kValMask = (1u << 30) - 1;
constexpr int64_t kMaxRows = (1LL << 30) - 1;

__global__ void hist_kernel(const uint64_t* __restrict__ keys, int64_t n_host, const unsigned long long* n_dev, int begin_bit, int n_passes,
                            unsigned long long* __restrict__ hist /* [8][256] */) {
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;  // the row count may live on the device (no host sync in front of the sort)
  __shared__ unsigned int s_h[8][kRadix];
  for (int i = threadIdx.x; i < 8 * kRadix; i += blockDim.x) (&s_h[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane); i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + lane;
    const bool in = i < n;
    const uint64_t k = in ? keys[i] : 0;
    const unsigned act = __ballot_sync(0xffffffffu, in);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if (p >= n_passes) break;
      const int d = (int)((k >> (begin_bit + 8 * p)) & 255);
      // Digits of real columns are often constant across a warp (high bytes of doubles in [0,1),
      // sign extension of small integers): one vote tells; only then is the add warp-aggregated.
      // Random digits take one shared-memory atomic per key (aggregating those costs more than it saves).
      const int d0 = __shfl_sync(0xffffffffu, d, __ffs(act) - 1);
      const unsigned same = __ballot_sync(0xffffffffu, in && d == d0);
      if (same == act) {
        if (in && lane == __ffs(act) - 1) atomicAdd(&s_h[p][d], (unsigned)__popc(act));
      } else if (in) {
        atomicAdd(&s_h[p][d], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_passes * kRadix; i += blockDim.x) {
    const unsigned v = (&s_h[0][0])[i];
    if (v) atomicAdd(&hist[i], (unsigned long long)v);
  }
}

// exclusive scan of each pass's 256 bins: one CTA of 256 threads per pass
__global__ void scan_hist_kernel(const unsigned long long* __restrict__ hist, unsigned long long* __restrict__ base) {
  __shared__ unsigned long long s[kRadix];
  const int p = blockIdx.x, d = threadIdx.x;
  const unsigned long long v = hist[p * kRadix + d];
  s[d] = v;
  __syncthreads();
  for (int o = 1; o < kRadix; o <<= 1) {
    const unsigned long long t = d >= o ? s[d - o] : 0;
    __syncthreads();
    s[d] += t;
    __syncthreads();
  }
  base[p * kRadix + d] = s[d] - v;
}

struct PassArgs {
  const uint64_t* kin;
  uint64_t* kout;
  const uint32_t* vin;  // may be nullptr: keys only
  uint32_t* vout;
  int64_t n;
  const unsigned long long* n_dev;  // when set: the real row count (<= n, the bound the grid was sized for)
  int shift;            // digit = (key >> shift) & 255; shift < 0: digit = bit 31 of the VALUE (xor flip)
  int flip;
  const unsigned long long* base;  // [256] exclusive digit offsets of this pass
  uint32_t* status;                // [tiles][256], zeroed
  unsigned int* ticket;            // zeroed
  unsigned int* fail;
};

__global__ void __launch_bounds__(kThreads, 3) onesweep_kernel(const __grid_constant__ PassArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem);                         // [kTile]
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem + (size_t)kTile * 8);      // [kTile]
  uint32_t(*s_cnt)[kRadix] = reinterpret_cast<uint32_t(*)[kRadix]>(smem + (size_t)kTile * 12);  // [kWarps][256]
  uint32_t* s_start = reinterpret_cast<uint32_t*>(smem + (size_t)kTile * 12 + sizeof(uint32_t) * kWarps * kRadix);  // [256]
  unsigned long long* s_goff = reinterpret_cast<unsigned long long*>(s_start + kRadix);                            // [256]
  uint32_t* s_count = reinterpret_cast<uint32_t*>(s_goff + kRadix);                                                  // [256] tile counts
  __shared__ unsigned int s_tile;
  __shared__ uint32_t s_wsum[kWarps];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
  for (int i = tid; i < kWarps * kRadix; i += kThreads) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t tile_base = tile * kTile;
  const int64_t n = a.n_dev ? (int64_t)*a.n_dev : a.n;
  if (tile_base >= n) return;  // grid sized for an upper bound: tiles past the end are never waited on
  const int64_t wbase = tile_base + (int64_t)warp * (32 * kItems);
  const bool by_val = a.shift < 0;

  uint64_t key[kItems];
  uint16_t pos[kItems];
  uint8_t dig[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t idx = wbase + i * 32 + lane;
    key[i] = idx < n ? a.kin[idx] : ~0ULL;
    if (by_val) {
      const uint32_t v = idx < n ? a.vin[idx] : 0;
      dig[i] = idx < n ? (uint8_t)(((v >> 31) ^ (uint32_t)a.flip) & 1u) : (uint8_t)255;
    } else {
      dig[i] = (uint8_t)((key[i] >> a.shift) & 255);
    }
  }
  // rank inside the warp, in (item, lane) order
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int d = dig[i];
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader) { old = s_cnt[warp][d]; s_cnt[warp][d] = old + __popc(peers); }
    old = __shfl_sync(0xffffffffu, old, leader);
    pos[i] = (uint16_t)(old + __popc(peers & lt_mask));
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: counts of digit d over the warps -> tile count, published for the following tiles
    const int d = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { const uint32_t c = s_cnt[w][d]; s_cnt[w][d] = run; run += c; }
    const uint32_t count = run;
    s_count[d] = count;
    volatile uint32_t* st = a.status;
    st[tile * kRadix + d] = (tile == 0 ? kFlagPrefix : kFlagAgg) | count;
    // exclusive scan of `count` over the 256 digits -> start of each digit's run inside the tile
    uint32_t incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase_sum = 0;
    for (int w = 0; w < warp; ++w) wbase_sum += s_wsum[w];
    s_start[d] = wbase_sum + incl - count;
  }
  // Decoupled look-back: thread d walks back over the preceding tiles' status words of digit d,
  // adding local counts until it meets a tile that already published an inclusive prefix.  (A
  // warp-parallel variant that reads 32 predecessors at once was measured SLOWER — 2.75 vs 1.66 ms
  // per pass over 1e8 keys: its retries re-read whole windows and flood L2 with polling.)
  {
    const int d = tid;
    volatile uint32_t* st = a.status;
    uint32_t excl = 0;
    if (tile > 0) {
      int64_t t = tile - 1;
      long long spins = 0;
      while (true) {
        const uint32_t sw = st[t * kRadix + d];
        const uint32_t f = sw >> 30;
        if (f == 0) {
          if (++spins > (1LL << 22)) { atomicExch(a.fail, 1u); break; }  // a few seconds: never a hang
          continue;
        }
        excl += sw & kValMask;
        if (f == 2) break;
        --t;
      }
      st[tile * kRadix + d] = (2u << 30) | (excl + s_count[d]);
    }
    s_goff[d] = a.base[d] + excl;
  }
  __syncthreads();
  // reorder through shared memory
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int d = dig[i];
    const uint32_t p = s_start[d] + s_cnt[warp][d] + pos[i];
    pos[i] = (uint16_t)p;
    s_keys[p] = key[i];
  }
  if (a.vin) {
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int64_t idx = wbase + i * 32 + lane;
      s_vals[pos[i]] = idx < n ? a.vin[idx] : 0;
    }
  }
  __syncthreads();
  const int64_t rem = n - tile_base;
  const int valid = (int)(rem < kTile ? rem : kTile);  // padding keys are the greatest and come last: skipped
  for (int j = tid; j < valid; j += kThreads) {
    const uint64_t k = s_keys[j];
    int d;
    if (by_val) d = (int)(((s_vals[j] >> 31) ^ (uint32_t)a.flip) & 1u);
    else d = (int)((k >> a.shift) & 255);
    const unsigned long long o = s_goff[d] + (unsigned long long)(j - (int)s_start[d]);
    a.kout[o] = k;
    if (a.vin) a.vout[o] = s_vals[j];
  }
}

constexpr size_t kSmemBytes = (size_t)kTile * 12 + sizeof(uint32_t) * kWarps * kRadix + sizeof(uint32_t) * kRadix + sizeof(unsigned long long) * kRadix + sizeof(uint32_t) * kRadix;
}  // namespace rs

// Host driver: sorts n (key, value) pairs by key bits [begin_bit, end_bit) with stable LSD passes,
// ping-ponging between (k0, v0) and (k1, v1); *result_buf tells which pair holds the result.
// Optionally one more stable pass keyed on bit 31 of the value (the NULL flag of a row id).
struct RadixSorter {
  DevBuf meta;    // [8][256] hist, [9][256] base, tickets[16], fail
  DevBuf status;  // [tiles][256] u32, re-zeroed per pass
  unsigned long long* hist() { return (unsigned long long*)meta.p; }
  unsigned long long* base() { return (unsigned long long*)meta.p + 8 * 256; }
  unsigned int* tickets() { return (unsigned int*)((unsigned long long*)meta.p + 17 * 256); }
  unsigned int* fail() { return tickets() + 16; }

  // n_dev != nullptr: the row count is read on the device (it must be <= n, which sizes the grids)
  int32_t sort(ErrorSink& err, cudaStream_t stream, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, int64_t n,
               int begin_bit, int end_bit, bool null_flag_pass, int nulls_first, int64_t n_nulls, int* result_buf,
               const unsigned long long* n_dev = nullptr) {
    *result_buf = 0;
    if (n <= 1 && !n_dev) return DBX_OK;
    if (n < 1) return DBX_OK;
    if (n > rs::kMaxRows) { err.set("sort: more than 2^30 - 1 rows in one sort are not supported"); return DBX_ERR_UNSUPPORTED; }
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    DBX_CUDA_TRY(err, cudaGetDevice(&dev));
    if (!attr_set[dev]) {
      DBX_CUDA_TRY(err, cudaFuncSetAttribute(rs::onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs::kSmemBytes));
      attr_set[dev] = true;
    }
    const int n_passes = (end_bit - begin_bit + 7) / 8;
    const int64_t tiles = (n + rs::kTile - 1) / rs::kTile;
    const size_t meta_bytes = (size_t)17 * 256 * 8 + 17 * 4 + 64;
    DBX_CUDA_TRY(err, meta.ensure(meta_bytes));
    DBX_CUDA_TRY(err, status.ensure((size_t)tiles * 256 * 4));
    DBX_CUDA_TRY(err, cudaMemsetAsync(meta.p, 0, meta_bytes, stream));
    rs::hist_kernel<<<(int)std::min<int64_t>(kNumSMs * 8, (n + 255) / 256), 256, 0, stream>>>(k0, n, n_dev, begin_bit, n_passes, hist());
    rs::scan_hist_kernel<<<n_passes, 256, 0, stream>>>(hist(), base());
    count_launch(2);
    if (null_flag_pass) {  // digit 0 / 1 of the extra pass: base offsets from the NULL count
      unsigned long long b2[2];
      const unsigned long long nn = (unsigned long long)n_nulls;
      // digit = is_null ^ nulls_first: with nulls_first the NULL rows get digit 0
      b2[0] = 0;
      b2[1] = nulls_first ? nn : (unsigned long long)n - nn;
      DBX_CUDA_TRY(err, cudaMemcpyAsync(base() + 8 * 256, b2, 16, cudaMemcpyHostToDevice, stream));
    }
    int cur = 0;
    for (int p = 0; p < n_passes + (null_flag_pass ? 1 : 0); ++p) {
      DBX_CUDA_TRY(err, cudaMemsetAsync(status.p, 0, (size_t)tiles * 256 * 4, stream));
      rs::PassArgs a;
      a.kin = cur ? k1 : k0; a.kout = cur ? k0 : k1;
      a.vin = cur ? v1 : v0; a.vout = cur ? v0 : v1;
      a.n = n;
      a.n_dev = n_dev;
      const bool extra = p == n_passes;
      a.shift = extra ? -1 : begin_bit + 8 * p;
      a.flip = nulls_first ? 1 : 0;
      a.base = base() + (extra ? 8 : p) * 256;
      a.status = (uint32_t*)status.p;
      a.ticket = tickets() + p;
      a.fail = fail();
      rs::onesweep_kernel<<<(unsigned)tiles, rs::kThreads, rs::kSmemBytes, stream>>>(a);
      count_launch();
      DBX_CUDA_TRY(err, cudaGetLastError());
      cur ^= 1;
    }
    *result_buf = cur;
    return DBX_OK;
  }
};


}  // namespace
}  // namespace dbx
