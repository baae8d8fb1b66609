// sort.cu — DBX_OP_TOPK: `ORDER BY key [ASC|DESC] [NULLS FIRST|LAST] [LIMIT k]` on the device.
//
// Reference pipeline replaced (paths relative to /root/reference):
//   TransformSortPartial (per block sort + limit)   src/query/pipeline/transforms/src/processors/transforms/sorts/sort_partial.rs:24-60
//     DataBlock::sort_with_type / SortCompare       src/query/expression/src/kernels/sort.rs:91-111, sort_compare.rs:197-296
//   limit-aware k-way merge                         sorts/sort_merge*.rs, sorts/core/{merger,loser_tree}.rs
//   fused TopN with a runtime boundary filter       src/query/service/src/pipelines/processors/transforms/top_n/transform_partial_top_n.rs:73-130
//
// Every key is mapped to an order-preserving u64 (OrderedFloat order: NaN greatest, -0 == +0,
// src/common/base/src/base/ordered_float.rs:147-201); ties are broken by ascending row id, as in
// the oracle.  Everything below is hand-written (no CUB):
//
//   LIMIT k (k <= 4 Mi)   streaming top-k: the column is read ONCE (8 B/row, 256-bit streaming
//                         loads); a row survives only if it beats the boundary (the k-th best key
//                         so far — the reference's TopN boundary filter), kept in DEVICE memory and
//                         tightened by a one-CTA radix-select ("cut") kernel that runs between
//                         scan launches.  No host synchronisation between chunks: the host only
//                         chooses chunk sizes that provably fit the candidate list, or launches
//                         one optimistic scan over the rest and checks an overflow flag once.
//   no LIMIT (limit = 0)  full sort: (ordered key, row id) pairs are radix-sorted with a
//                         onesweep-style LSD sort (one global histogram pass for all digits, then
//                         one read+write pass per 8-bit digit with decoupled look-back between
//                         tiles); LSD passes are stable, so equal keys stay in row order.
#include <algorithm>
#include <vector>

#include "radix_sort.cuh"
#include "runtime.h"

namespace dbx {

namespace {

// ================================================================ key images
__device__ __forceinline__ uint64_t key_to_ord(uint64_t bits, int cls, bool asc) {
  uint64_t o;
  if (cls == VC_FLT) {
    double d = __longlong_as_double((long long)bits);
    if (d == 0.0) d = 0.0;  // -0 == +0
    o = f64_to_ordered(d);
  } else if (cls == VC_INT) {
    o = bits ^ 0x8000000000000000ULL;
  } else {
    o = bits;
  }
  return asc ? o : ~o;
}

__device__ __forceinline__ uint64_t load_widened(const DevCol& c, int64_t row, uint64_t pol) {
  const char* base = (const char*)c.data;
  switch (c.dtype) {
    case DBX_I64: case DBX_U64: case DBX_F64: return ld_stream_u64(base + row * 8, pol);
    case DBX_I32: return (uint64_t)(int64_t)(int32_t)ld_stream_u32(base + row * 4, pol);
    case DBX_U32: return ld_stream_u32(base + row * 4, pol);
    case DBX_F32: return (uint64_t)__double_as_longlong((double)__uint_as_float(ld_stream_u32(base + row * 4, pol)));
    case DBX_I16: return (uint64_t)(int64_t)(int16_t)ld_stream_u16(base + row * 2, pol);
    case DBX_U16: return ld_stream_u16(base + row * 2, pol);
    case DBX_I8: return (uint64_t)(int64_t)(int8_t)ld_stream_u8(base + row, pol);
    default: return ld_stream_u8(base + row, pol);
  }
}

__device__ __forceinline__ void store_narrow_key(void* out, int64_t i, int dtype, uint64_t b) {
  switch (dtype) {
    case DBX_I8: case DBX_U8: ((uint8_t*)out)[i] = (uint8_t)b; break;
    case DBX_I16: case DBX_U16: ((uint16_t*)out)[i] = (uint16_t)b; break;
    case DBX_I32: case DBX_U32: ((uint32_t*)out)[i] = (uint32_t)b; break;
    case DBX_F32: ((float*)out)[i] = (float)__longlong_as_double((long long)b); break;
    default: ((uint64_t*)out)[i] = b; break;
  }
}

inline int grid_1d(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8)); }

// (the radix sort lives in radix_sort.cuh)
// ================================================================ streaming top-k
// Device state words of one candidate list
enum : int { ST_COUNT = 0, ST_BOUND = 1, ST_OVERFLOW = 2, ST_WORDS = 4 };

struct CandList {
  uint64_t* ord;    // order-preserving image (nullptr: list keyed by row id only — the NULL rows)
  uint64_t* rowid;  // global row ordinal
  uint64_t* bits;   // original value bits (nullptr for the NULL list)
  unsigned long long* state;  // [ST_WORDS]
  int64_t cap;
};

__device__ __forceinline__ void cand_append_warp(const CandList& l, bool keep, uint64_t o, uint64_t rid, uint64_t b, int lane) {
  const uint32_t bal = __ballot_sync(0xffffffffu, keep);
  if (!bal) return;
  unsigned long long base = 0;
  if (lane == __ffs(bal) - 1) base = atomicAdd(&l.state[ST_COUNT], (unsigned long long)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
  if (keep) {
    const unsigned long long pos = base + __popc(bal & ((1u << lane) - 1));
    if ((int64_t)pos < l.cap) {
      if (l.ord) l.ord[pos] = o;
      l.rowid[pos] = rid;
      if (l.bits) l.bits[pos] = b;
    } else {
      l.state[ST_OVERFLOW] = 1;  // never silently: the host replays the chunk in pieces that fit
    }
  }
}

// One pass over `n` rows of the key column.  Only ord <= boundary (read from device memory) can
// still be in the top k.  FAST: 8-byte column, 32 B aligned, no validity: a tile is 2048 rows, two
// 256-bit loads per thread, and the next tile's two loads are issued before this tile is examined
// (128 B in flight per thread); the grid is exactly the number of resident CTAs.
template <bool FAST>
__global__ void __launch_bounds__(256) topk_scan_kernel(const __grid_constant__ DevCol col, int64_t n, int64_t row_base, int cls,
                                                        int asc, const __grid_constant__ CandList cand,
                                                        const __grid_constant__ CandList nulls) {
  const uint64_t pol = make_policy_evict_first();
  const int lane = threadIdx.x & 31;
  const uint64_t boundary = cand.state[ST_BOUND];
  const uint64_t null_boundary = nulls.rowid ? nulls.state[ST_BOUND] : 0;
  if (FAST) {
    const int64_t n_tiles = n / 2048;  // whole tiles; the tail goes through the generic loop below
    const char* base = (const char*)col.data;
    u64x4 q0, q1;
    q0.x = q0.y = q0.z = q0.w = q1.x = q1.y = q1.z = q1.w = 0;
    int64_t tile = blockIdx.x;
    if (tile < n_tiles) {
      q0 = ld_stream_256(base + (tile * 2048 + 4 * (int64_t)threadIdx.x) * 8);
      q1 = ld_stream_256(base + (tile * 2048 + 1024 + 4 * (int64_t)threadIdx.x) * 8);
    }
    for (; tile < n_tiles; tile += gridDim.x) {
      __syncwarp();
      const u64x4 c0 = q0, c1 = q1;
      const int64_t nt = tile + gridDim.x;
      if (nt < n_tiles) {
        q0 = ld_stream_256(base + (nt * 2048 + 4 * (int64_t)threadIdx.x) * 8);
        q1 = ld_stream_256(base + (nt * 2048 + 1024 + 4 * (int64_t)threadIdx.x) * 8);
      }
      const uint64_t v[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      uint64_t o[8];
      bool any = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j] = key_to_ord(v[j], cls, asc != 0); any |= o[j] <= boundary; }
      if (!__any_sync(0xffffffffu, any)) continue;  // the common case once the boundary is tight
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t r = tile * 2048 + (j >> 2) * 1024 + 4 * (int64_t)threadIdx.x + (j & 3);
        cand_append_warp(cand, o[j] <= boundary, o[j], (uint64_t)(row_base + r), v[j], lane);
      }
    }
  }
  const int64_t first = FAST ? (n / 2048) * 2048 : 0;
  const int64_t n_tiles = (n - first + 1023) / 1024;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncwarp();
    const int64_t r0 = first + tile * 1024 + 4 * (int64_t)threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = r0 + j < n;
      const bool ok = in && (!col.validity || bit_test(col.validity, col.vbit_off + r0 + j));
      const uint64_t v = ok ? load_widened(col, r0 + j, pol) : 0;
      const uint64_t o = key_to_ord(v, cls, asc != 0);
      const uint64_t rid = (uint64_t)(row_base + r0 + j);
      cand_append_warp(cand, ok && o <= boundary, o, rid, v, lane);
      if (nulls.rowid) cand_append_warp(nulls, in && !ok && rid <= null_boundary, 0, rid, 0, lane);
    }
  }
}

// Radix select inside ONE CTA: keep the k smallest entries of a candidate list under the
// (ord, rowid) order, compact them to the front and publish the new boundary (the k-th entry's
// ord; for the NULL list its row id).  MSD passes over 8-bit digits of the 128-bit key only build
// a 256-bin histogram of the entries that still match the prefix found so far; the loop ends as
// soon as the bucket that contains the k-th entry is taken whole.  Does nothing when the list
// holds <= threshold entries.
struct CutArgs {
  CandList l;
  uint64_t* alt_ord;    // [k] scratch
  uint64_t* alt_rowid;
  uint64_t* alt_bits;
  int64_t k;
  int64_t threshold;
};
constexpr int kCutActive = 4096;  // entries of the bucket under examination kept in shared memory
// bucket that contains the k_rem-th entry of a 256-bin histogram: warp 0, eight bins per lane
__device__ __forceinline__ void cut_find_bucket(const unsigned int* hist, int64_t k_rem, unsigned int* out3, int tid) {
  if (tid >= 32) return;
  unsigned int c[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = hist[tid * 8 + j]; sum += c[j]; }
  unsigned int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (tid >= o) incl += up;
  }
  unsigned int cum = incl - sum;
  // the lane whose range [cum, cum + sum) contains the k_rem-th entry (k_rem >= 1)
  const bool mine = (int64_t)cum < k_rem && k_rem <= (int64_t)cum + sum;
  const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
  if (mine) {
    int b = 0;
    for (; b < 7; ++b) {
      if ((int64_t)cum + c[b] >= k_rem) break;
      cum += c[b];
    }
    out3[0] = (unsigned)(tid * 8 + b); out3[1] = cum; out3[2] = c[b];
  } else if (tid == 31 && (int64_t)total < k_rem) {  // cannot happen (k_rem <= matching entries); keep the state sane
    out3[0] = 255; out3[1] = total - c[7]; out3[2] = c[7];
  }
}
__global__ void __launch_bounds__(1024) topk_cut_kernel(const __grid_constant__ CutArgs a) {
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned int s_pick[3];  // bucket, entries before it, entries in it
  __shared__ unsigned int s_out, s_nact;
  extern __shared__ __align__(16) uint64_t s_cut_dyn[];  // active set (ord, rowid) once the bucket is small
  uint64_t* s_ao = s_cut_dyn;
  uint64_t* s_ar = s_cut_dyn + kCutActive;
  const CandList& l = a.l;
  const int64_t cnt = (int64_t)l.state[ST_COUNT];
  const int64_t n = cnt < l.cap ? cnt : l.cap;
  if (n <= a.threshold || n <= a.k) return;
  const int tid = threadIdx.x, lane = tid & 31;
  uint64_t th_hi = 0, th_lo = 0;  // prefix of the threshold key (ord, rowid)
  int64_t k_rem = a.k;
  bool closed = false;
  bool in_smem = false;   // the entries that still match the prefix sit in s_ao / s_ar
  int64_t n_act = n;      // entries matching the prefix found so far
  for (int p = l.ord ? 0 : 8; p < 16 && !closed; ++p) {
    if (tid < 256) s_hist[tid] = 0;
    if (tid == 0) s_nact = 0;
    __syncthreads();
    const int sh = 56 - 8 * (p & 7);
    // a bucket that fits is first copied to shared memory (same pass as its histogram)
    const bool gather = !in_smem && n_act <= kCutActive;
    const int64_t n_scan = in_smem ? n_act : n;
    for (int64_t i0 = tid - lane; i0 < n_scan; i0 += 1024) {
      const int64_t i = i0 + lane;
      bool m = i < n_scan;
      int d = 0;
      uint64_t o = 0, r = 0;
      if (m) {
        if (in_smem) { o = s_ao[i]; r = s_ar[i]; }
        else { o = l.ord ? l.ord[i] : 0; r = l.rowid[i]; }
        if (p < 8) {
          m = in_smem || p == 0 || (o >> (sh + 8)) == (th_hi >> (sh + 8));
          d = (int)((o >> sh) & 255);
        } else {
          m = in_smem || (o == th_hi && (p == 8 || (r >> (sh + 8)) == (th_lo >> (sh + 8))));
          d = (int)((r >> sh) & 255);
        }
      }
      const unsigned mm = __ballot_sync(0xffffffffu, m);
      if (!mm) continue;
      if (gather) {  // (all matching entries of this pass: exactly n_act of them)
        unsigned int base = 0;
        if (lane == __ffs(mm) - 1) base = atomicAdd(&s_nact, (unsigned)__popc(mm));
        base = __shfl_sync(0xffffffffu, base, __ffs(mm) - 1);
        if (m) {
          const unsigned int q = base + __popc(mm & ((1u << lane) - 1));
          if (q < (unsigned)kCutActive) { s_ao[q] = o; s_ar[q] = r; }
        }
      }
      const int d0 = __shfl_sync(0xffffffffu, d, __ffs(mm) - 1);
      const unsigned same = __ballot_sync(0xffffffffu, m && d == d0);
      if (same == mm) { if (lane == __ffs(mm) - 1) atomicAdd(&s_hist[d0], (unsigned)__popc(mm)); }
      else if (m) atomicAdd(&s_hist[d], 1u);
    }
    __syncthreads();
    cut_find_bucket(s_hist, k_rem, s_pick, tid);
    __syncthreads();
    const uint64_t bkt = s_pick[0];
    if (p < 8) th_hi |= bkt << sh; else th_lo |= bkt << sh;
    k_rem -= s_pick[1];
    if (gather) {
      // the gathered set matched the OLD prefix; keep only the chosen bucket for the next passes
      in_smem = true;
      __syncthreads();
      // in-place compaction by one warp-strided sweep through a second counter
      if (tid == 0) s_out = 0;
      __syncthreads();
      uint64_t ko[(kCutActive + 1023) / 1024], kr[(kCutActive + 1023) / 1024];
      bool kk[(kCutActive + 1023) / 1024];
#pragma unroll
      for (int t = 0; t < (kCutActive + 1023) / 1024; ++t) {
        const int64_t i = tid + 1024 * t;
        kk[t] = false;
        if (i < n_act) {
          ko[t] = s_ao[i]; kr[t] = s_ar[i];
          const int d = p < 8 ? (int)((ko[t] >> sh) & 255) : (int)((kr[t] >> sh) & 255);
          kk[t] = d == (int)bkt;
        }
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < (kCutActive + 1023) / 1024; ++t) {
        if (kk[t]) { const unsigned int q = atomicAdd(&s_out, 1u); s_ao[q] = ko[t]; s_ar[q] = kr[t]; }
      }
      __syncthreads();
    } else if (in_smem) {
      // narrow the shared-memory set to the chosen bucket
      if (tid == 0) s_out = 0;
      __syncthreads();
      uint64_t ko[(kCutActive + 1023) / 1024], kr[(kCutActive + 1023) / 1024];
      bool kk[(kCutActive + 1023) / 1024];
#pragma unroll
      for (int t = 0; t < (kCutActive + 1023) / 1024; ++t) {
        const int64_t i = tid + 1024 * t;
        kk[t] = false;
        if (i < n_act) {
          ko[t] = s_ao[i]; kr[t] = s_ar[i];
          const int d = p < 8 ? (int)((ko[t] >> sh) & 255) : (int)((kr[t] >> sh) & 255);
          kk[t] = d == (int)bkt;
        }
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < (kCutActive + 1023) / 1024; ++t) {
        if (kk[t]) { const unsigned int q = atomicAdd(&s_out, 1u); s_ao[q] = ko[t]; s_ar[q] = kr[t]; }
      }
      __syncthreads();
    }
    n_act = s_pick[2];
    if ((int64_t)s_pick[2] == k_rem) {  // the whole bucket is in: every key with this prefix passes
      const uint64_t low = sh ? ((1ULL << sh) - 1) : 0;
      if (p < 8) { th_hi |= low; th_lo = ~0ULL; } else { th_lo |= low; }
      closed = true;
    }
    __syncthreads();
  }
  // compaction of the entries <= threshold into the scratch arrays (at most k of them)
  if (tid == 0) s_out = 0;
  __syncthreads();
  for (int64_t i0 = tid - lane; i0 < n; i0 += 1024) {
    const int64_t i = i0 + lane;
    uint64_t o = 0, r = 0;
    bool keep = false;
    if (i < n) {
      o = l.ord ? l.ord[i] : 0;
      r = l.rowid[i];
      keep = o < th_hi || (o == th_hi && r <= th_lo);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (!bal) continue;
    unsigned int base = 0;
    if (lane == __ffs(bal) - 1) base = atomicAdd(&s_out, (unsigned)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (keep) {
      const unsigned int q = base + __popc(bal & ((1u << lane) - 1));
      if ((int64_t)q < a.k) {
        if (l.ord) a.alt_ord[q] = o;
        a.alt_rowid[q] = r;
        if (l.bits) a.alt_bits[q] = l.bits[i];
      }
    }
  }
  __syncthreads();
  const int64_t kept = (int64_t)s_out < a.k ? (int64_t)s_out : a.k;
  for (int64_t i = tid; i < kept; i += 1024) {
    if (l.ord) l.ord[i] = a.alt_ord[i];
    l.rowid[i] = a.alt_rowid[i];
    if (l.bits) l.bits[i] = a.alt_bits[i];
  }
  if (tid == 0) {
    l.state[ST_COUNT] = (unsigned long long)kept;
    l.state[ST_BOUND] = l.ord ? th_hi : th_lo;
  }
}

// Sort n <= 4096 candidates by (ord, rowid) inside one CTA by counting, for every entry, the
// entries that precede it (n^2 / 1024 comparisons per thread out of shared memory).
__global__ void __launch_bounds__(1024) small_rank_sort_kernel(const uint64_t* ord, const uint64_t* rowid, const uint64_t* bits, int n,
                                                               uint64_t* out_rowid, uint64_t* out_bits) {
  extern __shared__ __align__(16) uint64_t s_kr[];  // [n] ord, [n] rowid
  uint64_t* s_o = s_kr;
  uint64_t* s_r = s_kr + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { s_o[i] = ord ? ord[i] : 0; s_r[i] = rowid[i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t o = s_o[i], r = s_r[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (s_o[j] < o) || (s_o[j] == o && s_r[j] < r);
    out_rowid[rank] = r;
    if (bits) out_bits[rank] = bits[i];
  }
}

__global__ void topk_reset_kernel(unsigned long long* state) {
  if (threadIdx.x < 3 * ST_WORDS) state[threadIdx.x] = (threadIdx.x % ST_WORDS) == ST_BOUND && threadIdx.x < 2 * ST_WORDS ? ~0ULL : 0ULL;
}
__global__ void iota_u32_kernel(uint32_t* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}
__global__ void gather_u64_kernel(const uint64_t* src, const uint32_t* idx, uint64_t* dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

// Assemble the result block on the device: [key (original dtype, Nullable), row_id Int64].
struct EmitArgs {
  const uint64_t* bits;    // sorted valid rows: original value bits
  const uint64_t* rowid;   // sorted valid rows
  const uint64_t* null_rowid;  // sorted NULL rows
  int64_t take_valid, take_null;
  int32_t nulls_first, dtype;
  void* out_key;
  int64_t* out_row;
  uint8_t* out_valid_bytes;  // one byte per row
};
__global__ void topk_emit_kernel(const __grid_constant__ EmitArgs a) {
  const int64_t n_out = a.take_valid + a.take_null;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t first = a.nulls_first ? a.take_null : a.take_valid;
    const bool in_first = i < first;
    const bool is_null = a.nulls_first ? in_first : !in_first;
    const int64_t j = in_first ? i : i - first;
    if (is_null) {
      store_narrow_key(a.out_key, i, a.dtype, 0);
      a.out_row[i] = (int64_t)a.null_rowid[j];
      a.out_valid_bytes[i] = 0;
    } else {
      store_narrow_key(a.out_key, i, a.dtype, a.bits[j]);
      a.out_row[i] = (int64_t)a.rowid[j];
      a.out_valid_bytes[i] = 1;
    }
  }
}
__global__ void pack_bits_kernel(const uint8_t* bytes, int64_t n, uint8_t* bits) {
  const int64_t nb = (n + 7) / 8;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    for (int k = 0; k < 8; ++k) {
      const int64_t i = b * 8 + k;
      if (i < n && bytes[i]) v |= 1u << k;
    }
    bits[b] = (uint8_t)v;
  }
}

// ================================================================ full sort: ingest
// Appends one chunk of the key column to the (ord, row id | NULL flag, original bits) arrays.
__global__ void __launch_bounds__(256) sort_ingest_kernel(const __grid_constant__ DevCol col, int64_t n, int64_t row_base, int cls, int asc,
                                                          uint64_t* ord, uint32_t* rid, uint64_t* bits, unsigned long long* n_null,
                                                          unsigned long long* inexact) {
  const uint64_t pol = make_policy_evict_first();
  unsigned int nulls = 0;
  bool lossy = false;  // -0.0 and NaN payloads do not survive key -> ordered image -> key
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool ok = !col.validity || bit_test(col.validity, col.vbit_off + i);
    const uint64_t v = ok ? load_widened(col, i, pol) : 0;
    ord[row_base + i] = ok ? key_to_ord(v, cls, asc != 0) : 0;  // NULL rows: placed by the extra pass on the flag
    rid[row_base + i] = (uint32_t)(row_base + i) | (ok ? 0u : 0x80000000u);
    if (bits) bits[row_base + i] = v;
    nulls += !ok;
    if (ok && cls == VC_FLT) {
      const double d = __longlong_as_double((long long)v);
      lossy |= (d != d && v != 0x7FF8000000000000ULL) || (d == 0.0 && (v >> 63));
    }
  }
  if (inexact && __any_sync(0xffffffffu, lossy) && (threadIdx.x & 31) == 0) *inexact = 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nulls += __shfl_xor_sync(0xffffffffu, nulls, o);
  if ((threadIdx.x & 31) == 0 && nulls) atomicAdd(n_null, (unsigned long long)nulls);
}
// Multi-column ORDER BY: the keys of column c in the order the less significant columns have
// established so far (perm = sorted row id | flag of the previous step; nullptr = input order).
__global__ void __launch_bounds__(256) sort_gather_kernel(const uint64_t* ord_c, const uint32_t* rid_c, const uint32_t* perm, int64_t n,
                                                          uint64_t* o_ord, uint32_t* o_rid) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = perm ? (perm[i] & 0x7FFFFFFFu) : (uint32_t)i;
    o_ord[i] = ord_c[r];
    o_rid[i] = r | (rid_c[r] & 0x80000000u);
  }
}
struct SortEmitArgs {
  const uint32_t* rid;   // sorted: row id | NULL flag
  const uint64_t* bits;  // by original row id; nullptr: invert the sorted ordered images instead (no gather)
  const uint64_t* ord;   // sorted ordered images
  int32_t cls, asc;
  int64_t n;
  int32_t dtype;
  void* out_key;
  int64_t* out_row;
  uint8_t* out_valid_bytes;  // nullptr: key not nullable
};
__global__ void sort_emit_kernel(const __grid_constant__ SortEmitArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = a.rid[i];
    const uint32_t row = r & 0x7FFFFFFFu;
    const bool is_null = r >> 31;
    uint64_t b = 0;
    if (!is_null) {
      if (a.bits) b = a.bits[row];
      else {
        const uint64_t o = a.asc ? a.ord[i] : ~a.ord[i];
        b = a.cls == VC_FLT ? (uint64_t)__double_as_longlong(ordered_to_f64(o)) : (a.cls == VC_INT ? o ^ 0x8000000000000000ULL : o);
      }
    }
    store_narrow_key(a.out_key, i, a.dtype, b);
    a.out_row[i] = (int64_t)row;
    if (a.out_valid_bytes) a.out_valid_bytes[i] = is_null ? 0 : 1;
  }
}

}  // namespace

// ================================================================ operator
class TopkOp : public Op {
 public:
  dbx_topk_params prm;
  int n_cols = 0;
  int key_dtype = 0;
  bool key_nullable = false;
  int cls = 0;
  bool full_sort = false;
  Stager stager;
  RadixSorter sorter;
  // ---- top-k mode
  DevBuf ord, rowid, bits, null_rowid, state, alt_ord, alt_rowid, alt_bits, alt_null;
  DevBuf f_idx0, f_idx1, f_key0, f_key1, f_rowid, f_bits, f_null;  // finish(): sorted candidates
  PinnedBuf host;
  int64_t cap = 0;
  int64_t count_ub = 0, null_ub = 0;   // upper bounds on the list sizes known to the host without a sync
  int64_t rows_seen = 0;
  int64_t rows_at_last_cut = 0;
  int scan_ctas_per_sm = 0;
  DevBuf snap_ord, snap_rowid, snap_bits, snap_null;  // candidate lists as they were before the first optimistic scan of a push
  // ---- full-sort mode
  DevBuf s_ord[2], s_rid[2], s_bits;
  int64_t s_cap = 0;
  // further sort keys (ORDER BY a, b, ...): ordered images and row id | NULL flag per key, in input order
  int n_extra = 0;
  int x_dtype[DBX_MAX_SORT_KEYS - 1] = {}, x_cls[DBX_MAX_SORT_KEYS - 1] = {};
  bool x_nullable[DBX_MAX_SORT_KEYS - 1] = {};
  DevBuf x_ord[DBX_MAX_SORT_KEYS - 1], x_rid[DBX_MAX_SORT_KEYS - 1], x_cnt, w_ord[2], w_rid[2];
  std::unique_ptr<OwnedBlock> result;
  bool pulled = false;

  int32_t init(const dbx_topk_params* p, const int32_t* types, int32_t n, int dev) {
    DBX_TRY(base_init(dev));
    prm = *p;
    n_cols = n;
    if (p->key_col < 0 || p->key_col >= n) { err.set("top-k: key column outside the input schema"); return DBX_ERR_INVALID; }
    if (p->limit < 0) { err.set("top-k: negative limit"); return DBX_ERR_INVALID; }
    key_dtype = types[p->key_col] & 0xFF;
    key_nullable = (types[p->key_col] & DBX_NULLABLE) != 0;
    if (dtype_size(key_dtype) == 0) { err.set("top-k: key must be a numeric column"); return DBX_ERR_UNSUPPORTED; }
    cls = key_dtype == DBX_U64 ? VC_UINT : (dtype_class(key_dtype) == VC_FLT ? VC_FLT : VC_INT);
    full_sort = p->limit == 0 || p->limit > (1 << 22);  // no LIMIT (or one too large for the candidate list): sort everything
    n_extra = p->n_extra_keys;
    if (n_extra < 0 || n_extra > DBX_MAX_SORT_KEYS - 1) { err.set("sort: at most 4 sort keys"); return DBX_ERR_INVALID; }
    for (int j = 0; j < n_extra; ++j) {
      const int c = p->extra_key_cols[j];
      if (c < 0 || c >= n) { err.set("sort: key column outside the input schema"); return DBX_ERR_INVALID; }
      x_dtype[j] = types[c] & 0xFF;
      x_nullable[j] = (types[c] & DBX_NULLABLE) != 0;
      if (dtype_size(x_dtype[j]) == 0) { err.set("sort: keys must be numeric columns"); return DBX_ERR_UNSUPPORTED; }
      x_cls[j] = x_dtype[j] == DBX_U64 ? VC_UINT : (dtype_class(x_dtype[j]) == VC_FLT ? VC_FLT : VC_INT);
    }
    if (n_extra > 0) full_sort = true;  // several keys: sort everything, LIMIT cuts the sorted result
    DBX_CUDA_TRY(err, x_cnt.ensure(64));
    DBX_TRY(stager.init(dev, stream, &err));
    DBX_CUDA_TRY(err, host.ensure(256));
    DBX_CUDA_TRY(err, state.ensure(8 * ST_WORDS * 3 + 64));
    if (!full_sort) {
      const int64_t k = p->limit;
      // candidate list: large enough that a whole device-resident column usually fits behind the
      // boundary of its first few million rows; 3 x 8 B per entry
      static const int64_t cap_env = getenv("DBX_TOPK_CAP") ? atoll(getenv("DBX_TOPK_CAP")) : 0;
      cap = cap_env > 0 ? cap_env : std::max<int64_t>(1 << 22, 8 * k);
      cap = std::max<int64_t>(cap, 4 * k + 4096);
      DBX_CUDA_TRY(err, ord.ensure(cap * 8));
      DBX_CUDA_TRY(err, rowid.ensure(cap * 8));
      DBX_CUDA_TRY(err, bits.ensure(cap * 8));
      DBX_CUDA_TRY(err, alt_ord.ensure(k * 8));
      DBX_CUDA_TRY(err, alt_rowid.ensure(k * 8));
      DBX_CUDA_TRY(err, alt_bits.ensure(k * 8));
      if (key_nullable) {
        DBX_CUDA_TRY(err, null_rowid.ensure(cap * 8));
        DBX_CUDA_TRY(err, alt_null.ensure(k * 8));
      }
    }
    return reset();
  }

  int32_t reset() override {
    topk_reset_kernel<<<1, 32, 0, stream>>>((unsigned long long*)state.p);  // count 0, boundary = everything passes
    count_launch();
    DBX_CUDA_TRY(err, cudaMemsetAsync(x_cnt.p, 0, 64, stream));
    DBX_CUDA_TRY(err, cudaGetLastError());
    count_ub = null_ub = 0;
    rows_seen = 0;
    rows_at_last_cut = 0;
    result.reset();
    pulled = false;
    return DBX_OK;
  }

  CandList cand_list() const {
    CandList l;
    l.ord = (uint64_t*)ord.p; l.rowid = (uint64_t*)rowid.p; l.bits = (uint64_t*)bits.p;
    l.state = (unsigned long long*)state.p;
    l.cap = cap;
    return l;
  }
  CandList null_list() const {
    CandList l;
    l.ord = nullptr; l.rowid = (uint64_t*)null_rowid.p; l.bits = nullptr;  // rowid == nullptr: key not nullable
    l.state = (unsigned long long*)state.p + ST_WORDS;
    l.cap = cap;
    return l;
  }

  // cut both lists back to k when they hold more than `threshold` entries (device decides)
  int32_t launch_cuts(int64_t threshold, int64_t rows_so_far) {
    CutArgs a;
    a.l = cand_list();
    a.alt_ord = (uint64_t*)alt_ord.p; a.alt_rowid = (uint64_t*)alt_rowid.p; a.alt_bits = (uint64_t*)alt_bits.p;
    a.k = prm.limit; a.threshold = threshold;
    static std::atomic<bool> attr_set[64];
    if (!attr_set[device]) {
      DBX_CUDA_TRY(err, cudaFuncSetAttribute(topk_cut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCutActive * 16));
      attr_set[device] = true;
    }
    topk_cut_kernel<<<1, 1024, kCutActive * 16, stream>>>(a);
    count_launch();
    if (key_nullable) {
      CutArgs b;
      b.l = null_list();
      b.alt_ord = nullptr; b.alt_rowid = (uint64_t*)alt_null.p; b.alt_bits = nullptr;
      b.k = prm.limit; b.threshold = threshold;
      topk_cut_kernel<<<1, 1024, kCutActive * 16, stream>>>(b);
      count_launch();
    }
    DBX_CUDA_TRY(err, cudaGetLastError());
    count_ub = std::min(count_ub, std::max(threshold, prm.limit));
    null_ub = std::min(null_ub, std::max(threshold, prm.limit));
    rows_at_last_cut = rows_so_far;
    return DBX_OK;
  }

  int32_t launch_scan(const DevCol& col, int64_t off, int64_t m, int64_t row_base) {
    DevCol c = col;
    const int esz = dtype_size(key_dtype);
    c.data = (const char*)col.data + off * esz;
    if (c.validity) c.vbit_off += off;
    const bool fast = esz == 8 && !c.validity && ((reinterpret_cast<uintptr_t>(c.data) & 31) == 0);
    if (scan_ctas_per_sm == 0) {
      int a = 0, b2 = 0;
      DBX_CUDA_TRY(err, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, topk_scan_kernel<true>, 256, 0));
      DBX_CUDA_TRY(err, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b2, topk_scan_kernel<false>, 256, 0));
      scan_ctas_per_sm = std::max(1, std::min(a, b2));
    }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((m + 2047) / 2048, (int64_t)kNumSMs * scan_ctas_per_sm));
    CandList nl = null_list();
    if (!key_nullable) nl.rowid = nullptr;
    if (fast) topk_scan_kernel<true><<<grid, 256, 0, stream>>>(c, m, row_base, cls, prm.asc, cand_list(), nl);
    else topk_scan_kernel<false><<<grid, 256, 0, stream>>>(c, m, row_base, cls, prm.asc, cand_list(), nl);
    count_launch();
    DBX_CUDA_TRY(err, cudaGetLastError());
    return DBX_OK;
  }

  // rows [off, off + m) in pieces that provably fit the lists; a cut follows every piece
  int32_t scan_guaranteed(const DevCol& col, int64_t off, int64_t m) {
    const int64_t k = prm.limit;
    int64_t done = 0;
    while (done < m) {
      int64_t room = cap - std::max(count_ub, null_ub);
      if (room < cap / 2) { DBX_TRY(launch_cuts(2 * k, rows_seen + off + done)); room = cap - std::max(count_ub, null_ub); }
      const int64_t piece = std::min(m - done, room);
      DBX_TRY(launch_scan(col, off + done, piece, rows_seen + off + done));
      count_ub += piece;
      if (key_nullable) null_ub += piece;
      done += piece;
    }
    return DBX_OK;
  }

  // One pushed block.  Chunks grow geometrically (x8), each followed by a cut that tightens the
  // device-resident boundary, so a chunk adds about 8k survivors however long the column is.  A
  // chunk that provably fits the candidate list is "guaranteed"; larger ones are launched
  // optimistically (expected survivors << capacity) behind a snapshot of the lists, and the
  // overflow flags are checked ONCE at the end of the push: an overflow restores the snapshot and
  // replays the range in guaranteed pieces.  No host synchronisation otherwise.
  int32_t push_topk(const DevCol& col, int64_t n) {
    const int64_t k = prm.limit;
    int64_t done = 0;
    int64_t chunk = std::max<int64_t>(8 * k, 1 << 14);
    bool snap = false;
    int64_t snap_done = 0, snap_count_ub = 0, snap_null_ub = 0, snap_cut_pos = 0;
    unsigned long long* st = (unsigned long long*)state.p;
    while (done < n) {
      const int64_t seen = rows_seen + done;
      const int64_t rest = n - done;
      int64_t room = cap - std::max(count_ub, null_ub);
      const int64_t m = std::min(rest, chunk);
      if (room < m && std::max(count_ub, null_ub) > 2 * k) {  // the host's bounds are pessimistic: let the device cut
        DBX_TRY(launch_cuts(2 * k, seen));
        room = cap - std::max(count_ub, null_ub);
      }
      if (m > room && !snap) {  // first optimistic chunk of this push: snapshot the (cut) lists and their state
        DBX_TRY(launch_cuts(k, seen));
        DBX_CUDA_TRY(err, snap_ord.ensure(k * 8));
        DBX_CUDA_TRY(err, snap_rowid.ensure(k * 8));
        DBX_CUDA_TRY(err, snap_bits.ensure(k * 8));
        DBX_CUDA_TRY(err, cudaMemcpyAsync(snap_ord.p, ord.p, k * 8, cudaMemcpyDeviceToDevice, stream));
        DBX_CUDA_TRY(err, cudaMemcpyAsync(snap_rowid.p, rowid.p, k * 8, cudaMemcpyDeviceToDevice, stream));
        DBX_CUDA_TRY(err, cudaMemcpyAsync(snap_bits.p, bits.p, k * 8, cudaMemcpyDeviceToDevice, stream));
        if (key_nullable) {
          DBX_CUDA_TRY(err, snap_null.ensure(k * 8));
          DBX_CUDA_TRY(err, cudaMemcpyAsync(snap_null.p, null_rowid.p, k * 8, cudaMemcpyDeviceToDevice, stream));
        }
        DBX_CUDA_TRY(err, cudaMemcpyAsync(host.p, st, 8 * ST_WORDS * 2, cudaMemcpyDeviceToHost, stream));
        snap = true;
        snap_done = done; snap_count_ub = count_ub; snap_null_ub = null_ub; snap_cut_pos = rows_at_last_cut;
      }
      DBX_TRY(launch_scan(col, done, m, seen));
      count_ub += m;
      if (key_nullable) null_ub += m;
      done += m;
      // cut when the rows seen have doubled since the last cut (always inside a multi-chunk push)
      if (rows_seen + done - rows_at_last_cut >= rows_at_last_cut || done < n) DBX_TRY(launch_cuts(2 * k, rows_seen + done));
      chunk *= 8;
    }
    if (snap) {
      DBX_CUDA_TRY(err, cudaMemcpyAsync((char*)host.p + 64, st, 8 * ST_WORDS * 2, cudaMemcpyDeviceToHost, stream));
      DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
      const unsigned long long* before = (const unsigned long long*)host.p;
      const unsigned long long* after = (const unsigned long long*)((char*)host.p + 64);
      if (after[ST_OVERFLOW] || after[ST_WORDS + ST_OVERFLOW]) {
        // an optimistic chunk did not fit: back to the snapshot, then the same rows in pieces that do
        unsigned long long back[2 * ST_WORDS];
        memcpy(back, before, sizeof(back));
        back[ST_OVERFLOW] = back[ST_WORDS + ST_OVERFLOW] = 0;
        DBX_CUDA_TRY(err, cudaMemcpyAsync(ord.p, snap_ord.p, k * 8, cudaMemcpyDeviceToDevice, stream));
        DBX_CUDA_TRY(err, cudaMemcpyAsync(rowid.p, snap_rowid.p, k * 8, cudaMemcpyDeviceToDevice, stream));
        DBX_CUDA_TRY(err, cudaMemcpyAsync(bits.p, snap_bits.p, k * 8, cudaMemcpyDeviceToDevice, stream));
        if (key_nullable) DBX_CUDA_TRY(err, cudaMemcpyAsync(null_rowid.p, snap_null.p, k * 8, cudaMemcpyDeviceToDevice, stream));
        DBX_CUDA_TRY(err, cudaMemcpyAsync(st, back, sizeof(back), cudaMemcpyHostToDevice, stream));
        DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
        count_ub = std::min<int64_t>(snap_count_ub, (int64_t)before[ST_COUNT]);
        null_ub = std::min<int64_t>(snap_null_ub, (int64_t)before[ST_WORDS + ST_COUNT]);
        rows_at_last_cut = snap_cut_pos;
        DBX_TRY(scan_guaranteed(col, snap_done, n - snap_done));
      }
    }
    return DBX_OK;
  }

  // grow the full-sort arrays, keeping their contents
  int32_t sort_reserve(int64_t rows) {
    if (rows <= s_cap) return DBX_OK;
    int64_t ncap = std::max<int64_t>(rows, std::max<int64_t>(s_cap * 2, 1 << 20));
    auto grow = [&](DevBuf& b, size_t elt, bool keep) -> int32_t {
      DevBuf nb;
      DBX_CUDA_TRY(err, nb.ensure((size_t)ncap * elt));
      if (keep && b.p && rows_seen) DBX_CUDA_TRY(err, cudaMemcpyAsync(nb.p, b.p, (size_t)rows_seen * elt, cudaMemcpyDeviceToDevice, stream));
      DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
      b = std::move(nb);
      return DBX_OK;
    };
    DBX_TRY(grow(s_ord[0], 8, true));
    DBX_TRY(grow(s_rid[0], 4, true));
    DBX_TRY(grow(s_bits, 8, true));
    for (int j = 0; j < n_extra; ++j) { DBX_TRY(grow(x_ord[j], 8, true)); DBX_TRY(grow(x_rid[j], 4, true)); }
    s_cap = ncap;
    return DBX_OK;
  }

  int32_t push(const dbx_block* b) override {
    if (b->num_cols != n_cols) { err.set("push: block column count differs from the operator's input schema"); return DBX_ERR_INVALID; }
    const dbx_column& kc = b->cols[prm.key_col];
    if (kc.dtype != key_dtype || kc.len != b->num_rows) { err.set("push: key column does not match the input schema"); return DBX_ERR_INVALID; }
    const int64_t n = b->num_rows;
    if (n == 0) return DBX_OK;
    if (kc.is_const) { err.set("top-k over a constant key column is not supported"); return DBX_ERR_UNSUPPORTED; }
    if (kc.validity && !key_nullable) { err.set("push: validity bitmap on a key column declared non-nullable"); return DBX_ERR_INVALID; }
    DevCol col;
    DBX_TRY(stager.begin());
    DBX_TRY(stager.stage(kc, 0, &col));
    DBX_TRY(timing_begin());
    if (full_sort) {
      if (rows_seen + n > rs::kMaxRows) { err.set("sort: more than 2^30 - 1 rows are not supported"); return DBX_ERR_UNSUPPORTED; }
      DBX_TRY(sort_reserve(rows_seen + n));
      sort_ingest_kernel<<<grid_1d(n), 256, 0, stream>>>(col, n, rows_seen, cls, prm.asc, (uint64_t*)s_ord[0].p, (uint32_t*)s_rid[0].p,
                                                         (uint64_t*)s_bits.p, (unsigned long long*)state.p + ST_WORDS + ST_COUNT,
                                                         (unsigned long long*)state.p + ST_WORDS + ST_OVERFLOW);
      count_launch();
      DBX_CUDA_TRY(err, cudaGetLastError());
      for (int j = 0; j < n_extra; ++j) {
        const dbx_column& xc = b->cols[prm.extra_key_cols[j]];
        if (xc.dtype != x_dtype[j] || xc.len != n || xc.is_const) { err.set("push: sort key column does not match the input schema (constant keys unsupported)"); return DBX_ERR_INVALID; }
        if (xc.validity && !x_nullable[j]) { err.set("push: validity bitmap on a key column declared non-nullable"); return DBX_ERR_INVALID; }
        DevCol xcol;
        DBX_TRY(stager.stage(xc, 1 + j, &xcol));
        sort_ingest_kernel<<<grid_1d(n), 256, 0, stream>>>(xcol, n, rows_seen, x_cls[j], prm.extra_asc[j], (uint64_t*)x_ord[j].p, (uint32_t*)x_rid[j].p,
                                                           nullptr, (unsigned long long*)x_cnt.p + j, nullptr);
        count_launch();
        DBX_CUDA_TRY(err, cudaGetLastError());
      }
    } else {
      DBX_TRY(push_topk(col, n));
    }
    rows_seen += n;
    DBX_TRY(timing_end());
    DBX_TRY(stager.end());
    return DBX_OK;
  }

  int32_t dev_alloc(OwnedBlock* ob, size_t bytes, void** p) {
    DBX_CUDA_TRY(err, pool_alloc(device, stream, bytes ? bytes : 1, p));
    ob->dev_allocs.push_back(*p);
    return DBX_OK;
  }

  int32_t finish_full_sort() {
    const int64_t n = rows_seen;
    auto ob = std::make_unique<OwnedBlock>();
    ob->stream = stream;  // freed in order behind this operator's enqueued work
    ob->device = device;
    DBX_CUDA_TRY(err, cudaMemcpyAsync(host.p, (unsigned long long*)state.p + ST_WORDS, 8 * ST_WORDS, cudaMemcpyDeviceToHost, stream));
    DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
    const int64_t n_nulls = (int64_t)((unsigned long long*)host.p)[ST_COUNT];
    const bool needs_gather = ((unsigned long long*)host.p)[ST_OVERFLOW] != 0;  // a -0.0 or a NaN with a payload was ingested
    int buf = 0;
    const uint64_t* sorted_ord = nullptr;
    const uint32_t* sorted_rid = nullptr;
    if (n_extra > 0 && n > 1) {
      // least significant key first; every step is a STABLE sort of (key image, row id) in the order
      // the previous steps established, so earlier keys dominate and input order breaks the last ties
      unsigned long long xn[DBX_MAX_SORT_KEYS] = {};
      DBX_CUDA_TRY(err, cudaMemcpyAsync(xn, x_cnt.p, 8 * (DBX_MAX_SORT_KEYS - 1), cudaMemcpyDeviceToHost, stream));
      DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
      for (int i = 0; i < 2; ++i) { DBX_CUDA_TRY(err, w_ord[i].ensure((size_t)n * 8)); DBX_CUDA_TRY(err, w_rid[i].ensure((size_t)n * 4)); }
      int res = 1;  // which work pair holds the current order (none yet: the first gather goes to pair 0)
      const uint32_t* perm = nullptr;
      for (int c = n_extra - 1; c >= -1; --c) {
        const uint64_t* col_ord = c >= 0 ? (const uint64_t*)x_ord[c].p : (const uint64_t*)s_ord[0].p;
        const uint32_t* col_rid = c >= 0 ? (const uint32_t*)x_rid[c].p : (const uint32_t*)s_rid[0].p;
        const int64_t c_nulls = c >= 0 ? (int64_t)xn[c] : n_nulls;
        const int c_nulls_first = c >= 0 ? prm.extra_nulls_first[c] : prm.nulls_first;
        const int in = res ^ 1;
        sort_gather_kernel<<<grid_1d(n), 256, 0, stream>>>(col_ord, col_rid, perm, n, (uint64_t*)w_ord[in].p, (uint32_t*)w_rid[in].p);
        count_launch();
        DBX_CUDA_TRY(err, cudaGetLastError());
        int rb = 0;
        DBX_TRY(sorter.sort(err, stream, (uint64_t*)w_ord[in].p, (uint64_t*)w_ord[in ^ 1].p, (uint32_t*)w_rid[in].p, (uint32_t*)w_rid[in ^ 1].p, n, 0,
                            64, c_nulls > 0, c_nulls_first, c_nulls, &rb));
        res = rb ? (in ^ 1) : in;
        perm = (const uint32_t*)w_rid[res].p;
      }
      sorted_ord = (const uint64_t*)w_ord[res].p;
      sorted_rid = (const uint32_t*)w_rid[res].p;
    } else if (n > 1) {
      DBX_CUDA_TRY(err, s_ord[1].ensure((size_t)s_cap * 8));
      DBX_CUDA_TRY(err, s_rid[1].ensure((size_t)s_cap * 4));
      DBX_TRY(sorter.sort(err, stream, (uint64_t*)s_ord[0].p, (uint64_t*)s_ord[1].p, (uint32_t*)s_rid[0].p, (uint32_t*)s_rid[1].p, n, 0,
                          64, n_nulls > 0, prm.nulls_first, n_nulls, &buf));
    }
    if (!sorted_ord) { sorted_ord = (const uint64_t*)s_ord[buf].p; sorted_rid = (const uint32_t*)s_rid[buf].p; }
    const int64_t n_in = n;
    const int64_t n_out = (n_extra > 0 && prm.limit > 0) ? std::min<int64_t>(n_in, prm.limit) : n_in;
    void *okey = nullptr, *orow = nullptr, *ovb = nullptr, *obits = nullptr;
    const int esz = dtype_size(key_dtype);
    DBX_TRY(dev_alloc(ob.get(), (size_t)n_out * esz, &okey));
    DBX_TRY(dev_alloc(ob.get(), (size_t)n_out * 8, &orow));
    if (key_nullable) {
      DBX_TRY(dev_alloc(ob.get(), (size_t)n_out, &ovb));
      DBX_TRY(dev_alloc(ob.get(), (size_t)(n_out + 7) / 8 + 8, &obits));
    }
    if (n_out) {
      SortEmitArgs ea;
      ea.rid = sorted_rid; ea.bits = needs_gather ? (const uint64_t*)s_bits.p : nullptr; ea.n = n_out; ea.dtype = key_dtype;
      ea.ord = sorted_ord; ea.cls = cls; ea.asc = prm.asc;
      ea.out_key = okey; ea.out_row = (int64_t*)orow; ea.out_valid_bytes = (uint8_t*)ovb;
      sort_emit_kernel<<<grid_1d(n_out), 256, 0, stream>>>(ea);
      count_launch();
      if (key_nullable) { pack_bits_kernel<<<grid_1d((n_out + 7) / 8), 256, 0, stream>>>((const uint8_t*)ovb, n_out, (uint8_t*)obits); count_launch(); }
      DBX_CUDA_TRY(err, cudaGetLastError());
    }
    DBX_CUDA_TRY(err, cudaMemcpyAsync(host.p, sorter.meta.p ? (void*)sorter.fail() : state.p, 4, cudaMemcpyDeviceToHost, stream));
    DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
    if (sorter.meta.p && n > 1 && *(unsigned int*)host.p) { err.set("internal: radix sort look-back timed out"); return DBX_ERR_CUDA; }
    dbx_column kcol;
    memset(&kcol, 0, sizeof(kcol));
    kcol.dtype = key_dtype; kcol.mem = DBX_MEM_DEVICE; kcol.len = n_out; kcol.data = okey;
    if (key_nullable) { kcol.validity = (const uint8_t*)obits; kcol.null_count = n_out == n_in ? n_nulls : -1; }
    dbx_column rcol;
    memset(&rcol, 0, sizeof(rcol));
    rcol.dtype = DBX_I64; rcol.mem = DBX_MEM_DEVICE; rcol.len = n_out; rcol.data = orow;
    ob->cols.push_back(kcol);
    ob->cols.push_back(rcol);
    result = std::move(ob);
    return DBX_OK;
  }

  // sort `n` (ord?, rowid, bits?) entries by (ord, rowid) into out_rowid / out_bits
  int32_t sort_candidates(const uint64_t* c_ord, const uint64_t* c_rowid, const uint64_t* c_bits, int64_t n, uint64_t* out_rowid,
                          uint64_t* out_bits) {
    if (n == 0) return DBX_OK;
    if (n <= 4096) {
      static std::atomic<bool> attr_set[64];
      if (!attr_set[device]) {
        DBX_CUDA_TRY(err, cudaFuncSetAttribute(small_rank_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
        attr_set[device] = true;
      }
      small_rank_sort_kernel<<<1, 1024, (size_t)n * 16, stream>>>(c_ord, c_rowid, c_bits, (int)n, out_rowid, out_bits);
      count_launch();
      DBX_CUDA_TRY(err, cudaGetLastError());
      return DBX_OK;
    }
    // two stable radix sorts of a permutation: by row id, then by the ordered key
    DBX_CUDA_TRY(err, f_idx0.ensure(n * 4));
    DBX_CUDA_TRY(err, f_idx1.ensure(n * 4));
    DBX_CUDA_TRY(err, f_key0.ensure(n * 8));
    DBX_CUDA_TRY(err, f_key1.ensure(n * 8));
    iota_u32_kernel<<<grid_1d(n), 256, 0, stream>>>((uint32_t*)f_idx0.p, n);
    count_launch();
    DBX_CUDA_TRY(err, cudaMemcpyAsync(f_key0.p, c_rowid, n * 8, cudaMemcpyDeviceToDevice, stream));
    int buf = 0;
    DBX_TRY(sorter.sort(err, stream, (uint64_t*)f_key0.p, (uint64_t*)f_key1.p, (uint32_t*)f_idx0.p, (uint32_t*)f_idx1.p, n, 0, 48, false, 0, 0, &buf));
    uint32_t* idx = (uint32_t*)(buf ? f_idx1.p : f_idx0.p);
    uint32_t* idx_other = (uint32_t*)(buf ? f_idx0.p : f_idx1.p);
    if (c_ord) {
      uint64_t* kb = (uint64_t*)(buf ? f_key0.p : f_key1.p);  // the buffer the first sort left free
      uint64_t* kb_other = (uint64_t*)(buf ? f_key1.p : f_key0.p);
      gather_u64_kernel<<<grid_1d(n), 256, 0, stream>>>(c_ord, idx, kb, n);
      count_launch();
      int buf2 = 0;
      DBX_TRY(sorter.sort(err, stream, kb, kb_other, idx, idx_other, n, 0, 64, false, 0, 0, &buf2));
      if (buf2) idx = idx_other;
    }
    gather_u64_kernel<<<grid_1d(n), 256, 0, stream>>>(c_rowid, idx, out_rowid, n);
    if (c_bits) gather_u64_kernel<<<grid_1d(n), 256, 0, stream>>>(c_bits, idx, out_bits, n);
    count_launch(2);
    DBX_CUDA_TRY(err, cudaGetLastError());
    return DBX_OK;
  }

  int32_t finish() override {
    if (full_sort) return finish_full_sort();
    const int64_t k = prm.limit;
    DBX_TRY(launch_cuts(k, rows_seen));  // both lists down to <= k entries
    unsigned long long* st = (unsigned long long*)state.p;
    DBX_CUDA_TRY(err, cudaMemcpyAsync(host.p, st, 8 * ST_WORDS * 2, cudaMemcpyDeviceToHost, stream));
    DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
    const unsigned long long* h = (const unsigned long long*)host.p;
    if (h[ST_OVERFLOW] || h[ST_WORDS + ST_OVERFLOW]) { err.set("internal: top-k candidate list overflow"); return DBX_ERR_CUDA; }
    const int64_t n_valid = std::min<int64_t>((int64_t)h[ST_COUNT], k);
    const int64_t n_nulls = key_nullable ? std::min<int64_t>((int64_t)h[ST_WORDS + ST_COUNT], k) : 0;
    int64_t take_null, take_valid;
    if (prm.nulls_first) { take_null = n_nulls; take_valid = std::min<int64_t>(k - take_null, n_valid); }
    else { take_valid = n_valid; take_null = std::min<int64_t>(k - take_valid, n_nulls); }
    const int64_t n_out = take_null + take_valid;
    DBX_CUDA_TRY(err, f_rowid.ensure(std::max<int64_t>(n_valid, 1) * 8));
    DBX_CUDA_TRY(err, f_bits.ensure(std::max<int64_t>(n_valid, 1) * 8));
    DBX_TRY(sort_candidates((const uint64_t*)ord.p, (const uint64_t*)rowid.p, (const uint64_t*)bits.p, n_valid, (uint64_t*)f_rowid.p, (uint64_t*)f_bits.p));
    if (n_nulls) {
      DBX_CUDA_TRY(err, f_null.ensure(n_nulls * 8));
      DBX_TRY(sort_candidates(nullptr, (const uint64_t*)null_rowid.p, nullptr, n_nulls, (uint64_t*)f_null.p, nullptr));
    }
    // output block: [key (original dtype, nullable), row_id Int64], assembled on the device
    auto ob = std::make_unique<OwnedBlock>();
    ob->stream = stream;  // freed in order behind this operator's enqueued work
    ob->device = device;
    const int esz = dtype_size(key_dtype);
    void *okey = nullptr, *orow = nullptr, *ovb = nullptr, *obits = nullptr;
    DBX_TRY(dev_alloc(ob.get(), (size_t)n_out * esz, &okey));
    DBX_TRY(dev_alloc(ob.get(), (size_t)n_out * 8, &orow));
    DBX_TRY(dev_alloc(ob.get(), (size_t)n_out + 1, &ovb));
    DBX_TRY(dev_alloc(ob.get(), (size_t)(n_out + 7) / 8 + 8, &obits));
    if (n_out) {
      EmitArgs ea;
      ea.bits = (const uint64_t*)f_bits.p; ea.rowid = (const uint64_t*)f_rowid.p; ea.null_rowid = (const uint64_t*)f_null.p;
      ea.take_valid = take_valid; ea.take_null = take_null; ea.nulls_first = prm.nulls_first; ea.dtype = key_dtype;
      ea.out_key = okey; ea.out_row = (int64_t*)orow; ea.out_valid_bytes = (uint8_t*)ovb;
      topk_emit_kernel<<<grid_1d(n_out), 256, 0, stream>>>(ea);
      pack_bits_kernel<<<grid_1d((n_out + 7) / 8), 256, 0, stream>>>((const uint8_t*)ovb, n_out, (uint8_t*)obits);
      count_launch(2);
      DBX_CUDA_TRY(err, cudaGetLastError());
    }
    dbx_column kcol;
    memset(&kcol, 0, sizeof(kcol));
    kcol.dtype = key_dtype; kcol.mem = DBX_MEM_DEVICE; kcol.len = n_out; kcol.data = okey;
    if (key_nullable) { kcol.validity = (const uint8_t*)obits; kcol.null_count = take_null; }
    dbx_column rcol;
    memset(&rcol, 0, sizeof(rcol));
    rcol.dtype = DBX_I64; rcol.mem = DBX_MEM_DEVICE; rcol.len = n_out; rcol.data = orow;
    ob->cols.push_back(kcol);
    ob->cols.push_back(rcol);
    result = std::move(ob);
    return DBX_OK;
  }

  int32_t pull(int32_t out_mem, dbx_block* out, int32_t* has_block) override;
};

}  // namespace dbx

namespace dbx {

int32_t TopkOp::pull(int32_t out_mem, dbx_block* out, int32_t* has_block) {
  if (!finished) { err.set("pull before finish"); return DBX_ERR_STATE; }
  if (pulled || !result) { *has_block = 0; return DBX_OK; }
  pulled = true;
  *has_block = 1;
  return pull_owned_block(result, device, stream, err, out_mem, out);
}

Op* make_topk_op(const dbx_topk_params* p, const int32_t* types, int32_t n, int device, int32_t* st) {
  auto* op = new TopkOp();
  *st = op->init(p, types, n, device);
  if (*st != DBX_OK) { g_create_error.set(op->err.msg); delete op; return nullptr; }
  return op;
}

}  // namespace dbx
