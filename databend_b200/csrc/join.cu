// join.cu — DBX_OP_JOIN: inner hash join on one integer key column.
//
// Reference replaced (paths relative to /root/reference/src/query/service/src/pipelines/processors/transforms):
//   Join trait (add_block / final_build / probe_block -> JoinStream / final_probe)   new_hash_join/join.rs:26-53
//   TransformHashJoin stage machine Build -> BuildFinal -> Probe                    new_hash_join/transform_hash_join.rs:39-230
//   BasicHashJoin::{add_block, final_build}                                         new_hash_join/memory/basic.rs:77-160
//   HashJoinHashTable::{with_build_row_num, insert, probe}                          hash_join_table/hashjoin_hashtable.rs:95-190
//   InnerHashJoin::probe_block / InnerHashJoinStream::next                          new_hash_join/memory/inner_join.rs:122-262
//
// B200 design.  Build rows stay in HBM as columns.  The table is an open-addressed multimap of
// 32-byte entries {key, build_row + 1 | validity flags, payload0, payload1} = one sector = one
// 256-bit load: the probe of a row with up to two 8-byte build columns next to the key touches
// ONE random sector (the reference reads an 8-byte header, then chases the entry chain, then
// gathers the build row).  A 1e9-row probe into a 1e7-row build side is bound by HBM's
// random-sector rate, not by bytes: every avoided gather is worth as much as the probe itself.
// Build columns that do not fit the entry are gathered by build row as before.  The probe kernel streams the probe key column once, walks
// buckets until it sees an empty entry, and writes each joined row directly into the output
// columns at a position claimed with a warp-aggregated atomic: no (probe,build) index pairs are
// materialised and no second gather pass runs (the reference does DataBlock::take +
// take_column_vec).  Output row order is therefore unspecified — like the reference's when
// several threads build the chains — and results are compared as multisets.
// NULL keys never match (fixed_keys.rs: rows with a NULL key are skipped on both sides).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "runtime.h"

namespace dbx {

namespace {

constexpr int kJoinBlock = 256;
constexpr int kMaxJoinCols = 16;

struct JoinEntry {
  uint64_t key;
  uint64_t row1;  // (build row + 1) | validity of p0 << 62 | validity of p1 << 63; 0 = empty
  uint64_t p0, p1;  // raw bytes of up to two build columns (zero-extended to 8 bytes)
};
constexpr uint64_t kRowMask = (1ULL << 62) - 1;

// Radix layout: the table is cut into n_part regions of `region` entries (both powers of two);
// a key lives in region part_owner(key, n_part) (top hash bits) at slot hash & (region - 1)
// (low hash bits), probing wraps inside the region.  A probe block that was hash-partitioned the
// same way walks ONE region at a time, which then sits in L2 (and in the TLB) instead of
// scattering single-sector reads over a table many times larger than either.
struct JoinTableDev {
  JoinEntry* entries;  // cap entries (power of two), one per 32-byte sector
  int64_t cap;
  int64_t region;
  int32_t n_part;
  int32_t pad;
};
__device__ __forceinline__ int64_t join_home(const JoinTableDev& t, uint64_t k, int64_t* region_base) {
  const uint64_t h = agg_hash_u64(k);
  const int64_t part = t.n_part > 1 ? (int64_t)hash_to_part(h, t.n_part) : 0;
  *region_base = part * t.region;
  return (int64_t)(h & (uint64_t)(t.region - 1));
}

// build column carried inside the table entry
struct InlineColDev {
  const void* src;
  const uint8_t* valid_bytes;  // one byte per build row, or null
  int32_t size;
  int32_t on;
};

// One column copied into the output for every match.
struct JoinColDev {
  const void* src;
  void* dst;
  const uint8_t* src_validity;  // may be null
  int64_t src_vbit_off;
  uint8_t* dst_valid;           // one byte per output row, or null
  int32_t size;                 // bytes per value
  int32_t from;                 // build side: 0 gather by build row, 1 the entry's key, 2 entry.p0, 3 entry.p1
};

struct JoinProbeParams {
  DevCol key;
  JoinTableDev table;
  JoinColDev probe_cols[kMaxJoinCols];
  JoinColDev build_cols[kMaxJoinCols];
  int32_t n_probe_cols, n_build_cols;
  int32_t kind, pad;  // dbx_join_kind
  int64_t row_begin;  // rows [row_begin, row_begin + n_rows) of the (partitioned) probe columns
  int64_t n_rows;
  int64_t out_cap;
  unsigned long long* cursor;  // number of matches (may exceed out_cap: then the host retries)
};

__device__ __forceinline__ uint64_t load_key(const DevCol& c, int64_t row) {
  const char* base = (const char*)c.data;
  switch (c.dtype) {
    case DBX_I64: case DBX_U64: return ((const uint64_t*)base)[row];
    case DBX_I32: return (uint64_t)(int64_t)((const int32_t*)base)[row];
    case DBX_U32: return ((const uint32_t*)base)[row];
    case DBX_I16: return (uint64_t)(int64_t)((const int16_t*)base)[row];
    case DBX_U16: return ((const uint16_t*)base)[row];
    case DBX_I8: return (uint64_t)(int64_t)((const int8_t*)base)[row];
    default: return ((const uint8_t*)base)[row];
  }
}

// HashJoinHashTable::insert (hashjoin_hashtable.rs:110-141): every build row with a valid key
// claims the first free entry along its probe sequence (CAS on the row field; the key is written
// afterwards — build and probe are separated by a kernel boundary).
__device__ __forceinline__ uint64_t load_raw(const void* base, int size, int64_t row) {
  switch (size) {
    case 8: return ((const uint64_t*)base)[row];
    case 4: return ((const uint32_t*)base)[row];
    case 2: return ((const uint16_t*)base)[row];
    default: return ((const uint8_t*)base)[row];
  }
}
__global__ void join_build_kernel(const __grid_constant__ DevCol key, int64_t n_rows, int64_t row_base,
                                  const __grid_constant__ JoinTableDev t, const __grid_constant__ InlineColDev i0,
                                  const __grid_constant__ InlineColDev i1) {
  const int64_t mask = t.region - 1;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    if (key.validity && !bit_test(key.validity, key.vbit_off + r)) continue;
    const uint64_t k = load_key(key, r);
    uint64_t tag = (uint64_t)(row_base + r + 1);
    uint64_t p0 = 0, p1 = 0;
    if (i0.on) { p0 = load_raw(i0.src, i0.size, r); if (!i0.valid_bytes || i0.valid_bytes[r]) tag |= 1ULL << 62; }
    if (i1.on) { p1 = load_raw(i1.src, i1.size, r); if (!i1.valid_bytes || i1.valid_bytes[r]) tag |= 1ULL << 63; }
    int64_t rb;
    int64_t s = join_home(t, k, &rb);
    for (;;) {
      JoinEntry* e = t.entries + rb + s;
      unsigned long long old = atomicCAS((unsigned long long*)&e->row1, 0ULL, (unsigned long long)tag);
      if (old == 0ULL) { e->key = k; e->p0 = p0; e->p1 = p1; break; }
      s = (s + 1) & mask;
    }
  }
}

__device__ __forceinline__ void store_value(const JoinColDev& c, uint64_t bits, bool valid, int64_t dst_row) {
  switch (c.size) {
    case 8: ((uint64_t*)c.dst)[dst_row] = bits; break;
    case 4: ((uint32_t*)c.dst)[dst_row] = (uint32_t)bits; break;
    case 2: ((uint16_t*)c.dst)[dst_row] = (uint16_t)bits; break;
    default: ((uint8_t*)c.dst)[dst_row] = (uint8_t)bits; break;
  }
  if (c.dst_valid) c.dst_valid[dst_row] = valid ? 1 : 0;
}

__device__ __forceinline__ void copy_value(const JoinColDev& c, int64_t src_row, int64_t dst_row) {
  switch (c.size) {
    case 8: ((uint64_t*)c.dst)[dst_row] = ((const uint64_t*)c.src)[src_row]; break;
    case 4: ((uint32_t*)c.dst)[dst_row] = ((const uint32_t*)c.src)[src_row]; break;
    case 2: ((uint16_t*)c.dst)[dst_row] = ((const uint16_t*)c.src)[src_row]; break;
    default: ((uint8_t*)c.dst)[dst_row] = ((const uint8_t*)c.src)[src_row]; break;
  }
  if (c.dst_valid) c.dst_valid[dst_row] = c.src_validity ? (uint8_t)bit_test(c.src_validity, c.src_vbit_off + src_row) : 1;
}

// probe_block + InnerHashJoinStream::next fused: one thread per probe row.
// Per step of 256 rows a CTA (1) walks every row's probe sequence, counting its matches and
// keeping the first matching entry in registers, (2) reserves the step's output rows with ONE
// atomic on the global cursor (block scan of the counts; a per-warp reservation costs millions of
// same-address atomics per block and was the bottleneck), (3) writes the first match from
// registers and re-walks the sequence only for rows with several matches.
__device__ __forceinline__ JoinEntry load_entry(const JoinEntry* e) {
  JoinEntry r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(r.key), "=l"(r.row1), "=l"(r.p0), "=l"(r.p1) : "l"(e));
  return r;
}
__device__ __forceinline__ void emit_match(const JoinProbeParams& p, int64_t r, const JoinEntry& e, int64_t pos) {
  if (pos >= p.out_cap) return;
  for (int c = 0; c < p.n_probe_cols; ++c) copy_value(p.probe_cols[c], r, pos);
  for (int c = 0; c < p.n_build_cols; ++c) {
    const JoinColDev& jc = p.build_cols[c];
    if (jc.from == 0) copy_value(jc, (int64_t)(e.row1 & kRowMask) - 1, pos);
    else if (jc.from == 1) store_value(jc, e.key, true, pos);
    else if (jc.from == 2) store_value(jc, e.p0, (e.row1 >> 62) & 1, pos);
    else store_value(jc, e.p1, (e.row1 >> 63) & 1, pos);
  }
}
__global__ void __launch_bounds__(kJoinBlock) join_probe_kernel(const __grid_constant__ JoinProbeParams p) {
  __shared__ unsigned int s_warp[kJoinBlock / 32];
  __shared__ unsigned long long s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t mask = p.table.region - 1;
  const int64_t n_iter = (p.n_rows + (int64_t)gridDim.x * blockDim.x - 1) / ((int64_t)gridDim.x * blockDim.x);
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t r = p.row_begin + it * (int64_t)gridDim.x * blockDim.x + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = r < p.row_begin + p.n_rows;
    const bool live = in_range && !(p.key.validity && !bit_test(p.key.validity, p.key.vbit_off + r));
    uint64_t k = 0;
    int64_t b0 = 0, rb = 0;
    unsigned int n_match = 0;
    JoinEntry first;
    first.key = first.row1 = first.p0 = first.p1 = 0;
    if (live) {
      k = load_key(p.key, r);
      b0 = join_home(p.table, k, &rb);
      int64_t b = b0;
      for (;;) {  // an empty entry ends the probe sequence
        const JoinEntry e = load_entry(p.table.entries + rb + b);
        if (e.row1 == 0) break;
        if (e.key == k) { if (n_match == 0) first = e; ++n_match; }
        b = (b + 1) & mask;
      }
    }
    // semi / anti: the probe row itself is the output, at most once (a NULL key counts as no match)
    if (p.kind == DBX_JOIN_LEFT_SEMI) n_match = n_match ? 1u : 0u;
    else if (p.kind == DBX_JOIN_LEFT_ANTI) n_match = (in_range && n_match == 0) ? 1u : 0u;
    const bool outer_null_row = p.kind == DBX_JOIN_LEFT && in_range && n_match == 0;  // preserved row without a match
    if (outer_null_row) n_match = 1;
    // block-wide exclusive scan of the match counts -> one reservation per CTA and step
    unsigned int incl = n_match;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int tot = 0;
      for (int w = 0; w < kJoinBlock / 32; ++w) { const unsigned int c = s_warp[w]; s_warp[w] = tot; tot += c; }
      s_base = tot ? atomicAdd(p.cursor, (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    int64_t pos = (int64_t)s_base + s_warp[warp] + incl - n_match;
    __syncthreads();
    if (outer_null_row) {
      if (pos < p.out_cap) {
        for (int c = 0; c < p.n_probe_cols; ++c) copy_value(p.probe_cols[c], r, pos);
        for (int c = 0; c < p.n_build_cols; ++c) store_value(p.build_cols[c], 0, false, pos);
      }
    } else if (n_match && (p.kind == DBX_JOIN_LEFT_SEMI || p.kind == DBX_JOIN_LEFT_ANTI)) {
      if (pos < p.out_cap)
        for (int c = 0; c < p.n_probe_cols; ++c) copy_value(p.probe_cols[c], r, pos);
    } else if (n_match) {
      emit_match(p, r, first, pos++);
      if (n_match > 1) {  // duplicates of the key on the build side: walk again, skip the first
        int64_t b = b0;
        unsigned int seen = 0;
        for (;;) {
          const JoinEntry e = load_entry(p.table.entries + rb + b);
          if (e.row1 == 0) break;
          if (e.key == k && seen++ > 0) emit_match(p, r, e, pos++);
          b = (b + 1) & mask;
        }
      }
    }
  }
}

// Does any key occur twice on the build side?  One thread per slot walks the rest of the slot's
// probe sequence (short at load factor <= 0.5).  A build side without duplicates (the usual
// primary-key dimension table) lets the probe stop at its first match instead of walking on to the
// next empty entry: one dependent L2 round trip per probe row instead of two or more.
__global__ void join_dup_check_kernel(const __grid_constant__ JoinTableDev t, unsigned int* dup) {
  const int64_t mask = t.region - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += (int64_t)gridDim.x * blockDim.x) {
    const JoinEntry e = t.entries[i];
    if (e.row1 == 0) continue;
    const int64_t rb = i & ~mask;
    int64_t s = (i + 1) & mask;
    for (;;) {
      const JoinEntry f = t.entries[rb + s];
      if (f.row1 == 0) break;
      if (f.key == e.key) { *dup = 1; break; }
      s = (s + 1) & mask;
      if (rb + s == i) break;
    }
  }
}

// probe_block + JoinStream::next fused, two probe rows per thread: both rows' entry loads are in
// flight together (the probe is bound by dependent L2 round trips, not by bytes).  UNIQUE: the
// build side has no duplicate keys, so a row's walk ends at its first match.
template <bool UNIQUE>
__global__ void __launch_bounds__(kJoinBlock) join_probe2_kernel(const __grid_constant__ JoinProbeParams p) {
  __shared__ unsigned int s_warp[kJoinBlock / 32];
  __shared__ unsigned long long s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t mask = p.table.region - 1;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * 2;
  const int64_t n_iter = (p.n_rows + step - 1) / step;
  const int64_t row_end = p.row_begin + p.n_rows;
  for (int64_t it = 0; it < n_iter; ++it) {
    int64_t r[2];
    r[0] = p.row_begin + it * step + (int64_t)blockIdx.x * blockDim.x * 2 + threadIdx.x;
    r[1] = r[0] + blockDim.x;
    bool in_range[2], go[2];
    uint64_t k[2] = {0, 0};
    int64_t b0[2] = {0, 0}, rb[2] = {0, 0}, sl[2] = {0, 0};
    unsigned int n_match[2] = {0, 0};
    JoinEntry first[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      first[j].key = first[j].row1 = first[j].p0 = first[j].p1 = 0;
      in_range[j] = r[j] < row_end;
      go[j] = in_range[j] && !(p.key.validity && !bit_test(p.key.validity, p.key.vbit_off + r[j]));
      if (go[j]) {
        k[j] = load_key(p.key, r[j]);
        b0[j] = join_home(p.table, k[j], &rb[j]);
        sl[j] = b0[j];
      }
    }
    while (go[0] || go[1]) {  // an empty entry ends a probe sequence
      JoinEntry e[2];
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (go[j]) e[j] = load_entry(p.table.entries + rb[j] + sl[j]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (!go[j]) continue;
        if (e[j].row1 == 0) { go[j] = false; continue; }
        if (e[j].key == k[j]) {
          if (n_match[j] == 0) first[j] = e[j];
          ++n_match[j];
          if (UNIQUE) { go[j] = false; continue; }
        }
        sl[j] = (sl[j] + 1) & mask;
      }
    }
    bool outer_null[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      // semi / anti: the probe row itself is the output, at most once (a NULL key counts as no match)
      if (p.kind == DBX_JOIN_LEFT_SEMI) n_match[j] = n_match[j] ? 1u : 0u;
      else if (p.kind == DBX_JOIN_LEFT_ANTI) n_match[j] = (in_range[j] && n_match[j] == 0) ? 1u : 0u;
      outer_null[j] = p.kind == DBX_JOIN_LEFT && in_range[j] && n_match[j] == 0;  // preserved row without a match
      if (outer_null[j]) n_match[j] = 1;
    }
    // block-wide exclusive scan of the match counts -> one reservation per CTA and step
    const unsigned int mine = n_match[0] + n_match[1];
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int tot = 0;
      for (int w = 0; w < kJoinBlock / 32; ++w) { const unsigned int c = s_warp[w]; s_warp[w] = tot; tot += c; }
      s_base = tot ? atomicAdd(p.cursor, (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    int64_t pos = (int64_t)s_base + s_warp[warp] + incl - mine;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (outer_null[j]) {
        if (pos < p.out_cap) {
          for (int c = 0; c < p.n_probe_cols; ++c) copy_value(p.probe_cols[c], r[j], pos);
          for (int c = 0; c < p.n_build_cols; ++c) store_value(p.build_cols[c], 0, false, pos);
        }
        ++pos;
      } else if (n_match[j] && (p.kind == DBX_JOIN_LEFT_SEMI || p.kind == DBX_JOIN_LEFT_ANTI)) {
        if (pos < p.out_cap)
          for (int c = 0; c < p.n_probe_cols; ++c) copy_value(p.probe_cols[c], r[j], pos);
        ++pos;
      } else if (n_match[j]) {
        emit_match(p, r[j], first[j], pos++);
        if (!UNIQUE && n_match[j] > 1) {  // duplicates of the key on the build side: walk again, skip the first
          int64_t b = b0[j];
          unsigned int seen = 0;
          for (;;) {
            const JoinEntry e = load_entry(p.table.entries + rb[j] + b);
            if (e.row1 == 0) break;
            if (e.key == k[j] && seen++ > 0) emit_match(p, r[j], e, pos++);
            b = (b + 1) & mask;
          }
        }
      }
    }
  }
}

// Pull one table region into L2 with full-line sequential reads before it is probed: the probes
// themselves would fetch it as scattered 32-byte sectors, which HBM serves an order of magnitude
// slower than a stream.
__global__ void l2_prefetch_kernel(const char* base, int64_t bytes) {
  for (int64_t off = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 128; off < bytes; off += (int64_t)gridDim.x * blockDim.x * 128)
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + off));
}

__global__ void pack_bits_kernel(const uint8_t* bytes, int64_t n, uint8_t* bits) {
  int64_t nb = (n + 7) / 8;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    for (int k = 0; k < 8; ++k) {
      int64_t i = b * 8 + k;
      if (i < n && bytes[i]) v |= 1u << k;
    }
    bits[b] = (uint8_t)v;
  }
}

inline int64_t next_pow2_i64(int64_t x) {
  int64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}
inline int grid_rows(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + kJoinBlock - 1) / kJoinBlock, (int64_t)kNumSMs * 8)); }

// Device column that grows by appending pushed blocks (build side).
struct GrowCol {
  DevBuf data, valid_bytes;  // validity kept as one byte per row (simplifies appends at any offset)
  int64_t rows = 0;
  int size = 8;
  bool nullable = false;
};

__global__ void bits_to_bytes_kernel(const uint8_t* bits, int64_t bit_off, int64_t n, uint8_t* bytes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bytes[i] = bits ? (uint8_t)bit_test(bits, bit_off + i) : 1;
}

}  // namespace

class JoinOp : public Op {
 public:
  dbx_join_params prm;
  int n_build_cols = 0, n_probe_cols = 0;
  int build_dtype[kMaxJoinCols], probe_dtype[kMaxJoinCols];
  bool build_nullable[kMaxJoinCols], probe_nullable[kMaxJoinCols];
  Stager stager;
  std::vector<GrowCol> build;
  int64_t build_rows = 0;
  DevBuf table_buf, cursor;
  int64_t table_cap = 0;
  PinnedBuf host;
  int inline_col[2] = {-1, -1};
  int n_part = 1;
  int64_t region = 0;
  DevBuf part_counters;
  bool build_unique = false;  // no key occurs twice on the build side (checked after the build)
  DevBuf part_cols[kMaxJoinCols];
  std::vector<int64_t> part_offs;
  JoinTableDev table_view() const { return JoinTableDev{(JoinEntry*)table_buf.p, table_cap, region, n_part, 0}; }
  std::vector<std::unique_ptr<OwnedBlock>> outputs;  // joined blocks waiting to be pulled (device resident)
  size_t next_out = 0;

  // input_types = build schema (params.n_build_cols columns) followed by the probe schema.
  int32_t init(const dbx_join_params* p, const int32_t* types, int32_t n, int dev) {
    DBX_TRY(base_init(dev));
    prm = *p;
    if (p->kind < DBX_JOIN_INNER || p->kind > DBX_JOIN_LEFT) { err.set("join kind not built (INNER, LEFT, LEFT SEMI and LEFT ANTI are; right/full joins are next, SURVEY 8f.3)"); return DBX_ERR_UNSUPPORTED; }
    n_build_cols = p->n_build_cols;
    n_probe_cols = n - n_build_cols;
    if (n_build_cols <= 0 || n_probe_cols <= 0 || n_build_cols > kMaxJoinCols || n_probe_cols > kMaxJoinCols) {
      err.set("join: input_types must hold the build schema (params.n_build_cols columns) followed by the probe schema");
      return DBX_ERR_INVALID;
    }
    for (int i = 0; i < n_build_cols; ++i) { build_dtype[i] = types[i] & 0xFF; build_nullable[i] = (types[i] & DBX_NULLABLE) != 0; }
    for (int i = 0; i < n_probe_cols; ++i) { probe_dtype[i] = types[n_build_cols + i] & 0xFF; probe_nullable[i] = (types[n_build_cols + i] & DBX_NULLABLE) != 0; }
    if (p->build_key_col < 0 || p->build_key_col >= n_build_cols || p->probe_key_col < 0 || p->probe_key_col >= n_probe_cols) { err.set("join: key column outside the schema"); return DBX_ERR_INVALID; }
    auto int_key = [](int dt) { return dt != DBX_BOOL && dt != DBX_F32 && dt != DBX_F64 && dtype_size(dt) > 0; };
    if (!int_key(build_dtype[p->build_key_col]) || !int_key(probe_dtype[p->probe_key_col])) { err.set("join: keys must be integer columns"); return DBX_ERR_UNSUPPORTED; }
    for (int i = 0; i < n_build_cols; ++i) if (dtype_size(build_dtype[i]) == 0) { err.set("join: only fixed-width numeric columns are supported"); return DBX_ERR_UNSUPPORTED; }
    for (int i = 0; i < n_probe_cols; ++i) if (dtype_size(probe_dtype[i]) == 0) { err.set("join: only fixed-width numeric columns are supported"); return DBX_ERR_UNSUPPORTED; }
    // keys of different widths/signedness compare by value: both are widened to 64 bits
    // (sign-extended if signed), the common super type of the reference's key cast.  The one pair
    // with no 64-bit super type is (signed, UInt64): the widened images of -1 and 2^64-1 coincide,
    // so it is refused here (the reference's type checker casts both sides to a wider type first;
    // a caller wanting that join casts the keys before the operator, as the reference's planner does).
    {
      const int bk = build_dtype[p->build_key_col], pk = probe_dtype[p->probe_key_col];
      const bool bs = dtype_class(bk) == VC_INT, ps = dtype_class(pk) == VC_INT;
      if ((bk == DBX_U64 && ps) || (pk == DBX_U64 && bs)) { err.set("join: a signed key cannot be compared with a UInt64 key without a cast (no common 64-bit type)"); return DBX_ERR_UNSUPPORTED; }
    }
    build.resize(n_build_cols);
    for (int i = 0; i < n_build_cols; ++i) { build[i].size = dtype_size(build_dtype[i]); build[i].nullable = build_nullable[i]; }
    DBX_TRY(stager.init(dev, stream, &err));
    DBX_CUDA_TRY(err, cursor.ensure(64));
    DBX_CUDA_TRY(err, host.ensure(64));
    return DBX_OK;
  }

  // Join::add_block (build side): append the block's columns to the HBM-resident build side
  int32_t push(const dbx_block* b) override {
    if (b->num_cols != n_build_cols) { err.set("add_block: block does not match the build schema"); return DBX_ERR_INVALID; }
    const int64_t n = b->num_rows;
    if (n == 0) return DBX_OK;
    DBX_TRY(stager.begin());
    for (int c = 0; c < n_build_cols; ++c) {
      const dbx_column& col = b->cols[c];
      if (col.dtype != build_dtype[c] || col.len != n || col.is_const) { err.set("add_block: column dtype/length mismatch (const build columns unsupported)"); return DBX_ERR_INVALID; }
      DevCol dc;
      DBX_TRY(stager.stage(col, c, &dc));
      GrowCol& g = build[c];
      const size_t need = (size_t)(build_rows + n) * g.size;
      if (need > g.data.bytes) {  // grow, preserving the rows already appended (the hint avoids re-allocations)
        DevBuf nb;
        DBX_CUDA_TRY(err, nb.ensure(std::max({need, g.data.bytes * 2, (size_t)std::max<int64_t>(prm.expected_build_rows, 0) * g.size})));
        if (build_rows) DBX_CUDA_TRY(err, cudaMemcpyAsync(nb.p, g.data.p, (size_t)build_rows * g.size, cudaMemcpyDeviceToDevice, stream));
        DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
        g.data = std::move(nb);
      }
      DBX_CUDA_TRY(err, cudaMemcpyAsync((char*)g.data.p + (size_t)build_rows * g.size, dc.data, (size_t)n * g.size, cudaMemcpyDeviceToDevice, stream));
      if (g.nullable) {
        const size_t vneed = (size_t)(build_rows + n);
        if (vneed > g.valid_bytes.bytes) {
          DevBuf nb;
          DBX_CUDA_TRY(err, nb.ensure(std::max({vneed, g.valid_bytes.bytes * 2, (size_t)std::max<int64_t>(prm.expected_build_rows, 0)})));
          if (build_rows) DBX_CUDA_TRY(err, cudaMemcpyAsync(nb.p, g.valid_bytes.p, (size_t)build_rows, cudaMemcpyDeviceToDevice, stream));
          DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
          g.valid_bytes = std::move(nb);
        }
        bits_to_bytes_kernel<<<grid_rows(n), kJoinBlock, 0, stream>>>(dc.validity, dc.vbit_off, n, (uint8_t*)g.valid_bytes.p + build_rows);
        count_launch();
      }
    }
    build_rows += n;
    DBX_TRY(stager.end());
    return DBX_OK;
  }

  // Join::final_build: size the table for the build row count and insert every row
  int32_t finish() override {
    table_cap = std::max<int64_t>(next_pow2_i64(2 * std::max<int64_t>(build_rows, 1)), 1024);  // with_build_row_num
    if (build_rows >= (int64_t)kRowMask) { err.set("join: too many build rows"); return DBX_ERR_UNSUPPORTED; }
    // radix regions: cut a table that is far larger than L2 into pieces of <= 32 MB
    n_part = 1;
    {
      const int64_t bytes = table_cap * (int64_t)sizeof(JoinEntry);
      // Radix regions are OFF by default: measured on B200 (profiles/r02_ops_n1_after_rework.jsonl and
      // call F), 1e9 x 1e7 rows: 39.5 ms without regions vs 50.3 ms with 32 MB regions — once the
      // probe stops at its first match (unique build keys) one random HBM sector per row costs less
      // than the extra partition pass over the probe block.  DBX_JOIN_REGION_BYTES=<bytes> turns
      // them on (tests exercise both).
      int64_t target = 0;
      if (const char* e = getenv("DBX_JOIN_REGION_BYTES")) target = atoll(e);
      if (target > 0 && bytes > 3 * target) {
        while (n_part < kMaxParts && bytes / n_part > target) n_part *= 2;
      }
    }
    region = table_cap / n_part;
    DBX_CUDA_TRY(err, part_counters.ensure((size_t)kMaxParts * 8));
    if (n_part > 1) {  // a skewed key distribution could overfill a region: then fall back to one region
      PartParams cp;
      memset(&cp, 0, sizeof(cp));
      cp.key.data = build[prm.build_key_col].data.p;
      cp.key.dtype = build_dtype[prm.build_key_col];
      cp.n_cols = 0; cp.n_parts = n_part; cp.n_rows = build_rows;
      std::vector<int64_t> offs((size_t)n_part + 1);
      DBX_TRY(hash_partition_device(err, stream, cp, (unsigned long long*)part_counters.p, offs.data()));
      int64_t worst = 0;
      for (int i = 0; i < n_part; ++i) worst = std::max(worst, offs[i + 1] - offs[i]);
      if (worst * 10 > region * 7) { n_part = 1; region = table_cap; }
    }
    DBX_CUDA_TRY(err, table_buf.ensure((size_t)table_cap * sizeof(JoinEntry)));
    DBX_CUDA_TRY(err, cudaMemsetAsync(table_buf.p, 0, (size_t)table_cap * sizeof(JoinEntry), stream));
    if (build_rows) {
      GrowCol& kc = build[prm.build_key_col];
      DevCol key;
      memset(&key, 0, sizeof(key));
      key.data = kc.data.p;
      key.dtype = build_dtype[prm.build_key_col];
      // build-side validity is stored as bytes; expose it as a bitmap-free predicate by packing
      DevBuf kbits;
      if (kc.nullable) {
        DBX_CUDA_TRY(err, kbits.ensure((size_t)(build_rows + 7) / 8 + 8));
        pack_bits_kernel<<<grid_rows((build_rows + 7) / 8), kJoinBlock, 0, stream>>>((const uint8_t*)kc.valid_bytes.p, build_rows, (uint8_t*)kbits.p);
        count_launch();
        key.validity = (const uint8_t*)kbits.p;
      }
      JoinTableDev t = table_view();
      // the first two non-key build columns travel inside the entries
      InlineColDev ic[2];
      memset(ic, 0, sizeof(ic));
      inline_col[0] = inline_col[1] = -1;
      for (int c = 0, k = 0; c < n_build_cols && k < 2; ++c) {
        if (c == prm.build_key_col) continue;
        inline_col[k] = c;
        ic[k].src = build[c].data.p;
        ic[k].valid_bytes = build[c].nullable ? (const uint8_t*)build[c].valid_bytes.p : nullptr;
        ic[k].size = build[c].size;
        ic[k].on = 1;
        ++k;
      }
      join_build_kernel<<<grid_rows(build_rows), kJoinBlock, 0, stream>>>(key, build_rows, 0, t, ic[0], ic[1]);
      count_launch();
      DBX_CUDA_TRY(err, cudaGetLastError());
      DBX_CUDA_TRY(err, cudaMemsetAsync(cursor.p, 0, 16, stream));
      join_dup_check_kernel<<<grid_rows(table_cap), kJoinBlock, 0, stream>>>(t, (unsigned int*)cursor.p + 2);
      count_launch();
      DBX_CUDA_TRY(err, cudaGetLastError());
      DBX_CUDA_TRY(err, cudaMemcpyAsync(host.p, (unsigned int*)cursor.p + 2, 4, cudaMemcpyDeviceToHost, stream));
      DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
      build_unique = *(unsigned int*)host.p == 0 && !getenv("DBX_JOIN_NO_UNIQUE");
    } else {
      build_unique = true;
    }
    return DBX_OK;
  }

  void launch_probe(const JoinProbeParams& pp, int64_t rows) {
    static const bool old_probe = getenv("DBX_JOIN_OLD_PROBE") != nullptr;
    if (old_probe) { join_probe_kernel<<<grid_rows(rows), kJoinBlock, 0, stream>>>(pp); return; }
    const int grid = grid_rows((rows + 1) / 2);
    if (build_unique) join_probe2_kernel<true><<<grid, kJoinBlock, 0, stream>>>(pp);
    else join_probe2_kernel<false><<<grid, kJoinBlock, 0, stream>>>(pp);
  }

  // Join::probe_block: join one probe block; the joined block is queued for dbx_op_pull
  int32_t probe(const dbx_block* b) {
    if (!finished) { err.set("probe before final_build"); return DBX_ERR_STATE; }
    if (b->num_cols != n_probe_cols) { err.set("probe_block: block does not match the probe schema"); return DBX_ERR_INVALID; }
    const int64_t n = b->num_rows;
    if (n == 0) return DBX_OK;
    DevCol cols[kMaxJoinCols];
    DBX_TRY(stager.begin());
    for (int c = 0; c < n_probe_cols; ++c) {
      const dbx_column& col = b->cols[c];
      if (col.dtype != probe_dtype[c] || col.len != n || col.is_const) { err.set("probe_block: column dtype/length mismatch (const probe columns unsupported)"); return DBX_ERR_INVALID; }
      DBX_TRY(stager.stage(col, c, &cols[c]));
    }
    int64_t out_cap = n + n / 8 + 1024;  // optimistic: about one match per probe row
    DBX_TRY(timing_begin());
    // radix probe: reorder the block region by region (row order of a join result is unspecified)
    bool can_part = n_part > 1 && (n >= (1 << 16) || getenv("DBX_JOIN_REGION_BYTES"));
    for (int c = 0; c < n_probe_cols && can_part; ++c) can_part = !cols[c].validity;
    if (can_part) {
      PartParams pq;
      memset(&pq, 0, sizeof(pq));
      pq.key = cols[prm.probe_key_col];
      pq.n_cols = n_probe_cols; pq.n_parts = n_part; pq.n_rows = n;
      for (int c = 0; c < n_probe_cols; ++c) {
        const int sz = dtype_size(probe_dtype[c]);
        DBX_CUDA_TRY(err, part_cols[c].ensure((size_t)n * sz));
        pq.cols[c].src = cols[c].data; pq.cols[c].dst = part_cols[c].p; pq.cols[c].size = sz;
      }
      part_offs.assign((size_t)n_part + 1, 0);
      DBX_TRY(hash_partition_device(err, stream, pq, (unsigned long long*)part_counters.p, part_offs.data()));
      for (int c = 0; c < n_probe_cols; ++c) cols[c].data = part_cols[c].p;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
      auto ob = std::make_unique<OwnedBlock>();
    ob->stream = stream;  // freed in order behind this operator's enqueued work
      ob->device = device;
      JoinProbeParams pp;
      memset(&pp, 0, sizeof(pp));
      pp.key = cols[prm.probe_key_col];
      pp.table = table_view();
      pp.n_probe_cols = n_probe_cols;
      pp.n_build_cols = (prm.kind == DBX_JOIN_INNER || prm.kind == DBX_JOIN_LEFT) ? n_build_cols : 0;
      pp.kind = prm.kind;
      pp.n_rows = n;
      pp.out_cap = out_cap;
      pp.cursor = (unsigned long long*)cursor.p;
      std::vector<uint8_t*> valid_bytes;
      auto add_out = [&](JoinColDev& jc, int dtype, bool nullable) -> int32_t {
        void* d = nullptr;
        DBX_CUDA_TRY(err, pool_alloc(device, stream, (size_t)out_cap * dtype_size(dtype), &d));
        ob->dev_allocs.push_back(d);
        jc.dst = d;
        jc.size = dtype_size(dtype);
        uint8_t* vb = nullptr;
        if (nullable) {
          DBX_CUDA_TRY(err, pool_alloc(device, stream, (size_t)out_cap, (void**)&vb));
          ob->dev_allocs.push_back(vb);
        }
        jc.dst_valid = vb;
        valid_bytes.push_back(vb);
        dbx_column oc;
        memset(&oc, 0, sizeof(oc));
        oc.dtype = dtype; oc.mem = DBX_MEM_DEVICE; oc.data = d; oc.null_count = nullable ? -1 : 0;
        ob->cols.push_back(oc);
        return DBX_OK;
      };
      // output column order = probe projection then build projection (inner_join.rs:236-245)
      for (int c = 0; c < n_probe_cols; ++c) {
        pp.probe_cols[c].src = cols[c].data;
        pp.probe_cols[c].src_validity = cols[c].validity;
        pp.probe_cols[c].src_vbit_off = cols[c].vbit_off;
        DBX_TRY(add_out(pp.probe_cols[c], probe_dtype[c], probe_nullable[c]));
      }
      DevBuf build_bits[kMaxJoinCols];
      for (int c = 0; c < pp.n_build_cols; ++c) {
        pp.build_cols[c].src = build[c].data.p;
        pp.build_cols[c].from = c == prm.build_key_col ? 1 : (c == inline_col[0] ? 2 : (c == inline_col[1] ? 3 : 0));
        if (build[c].nullable && pp.build_cols[c].from == 0) {  // bytes -> use the byte array directly through a 1-byte "bitmap" trick: pack once
          DBX_CUDA_TRY(err, build_bits[c].ensure((size_t)(build_rows + 7) / 8 + 8));
          pack_bits_kernel<<<grid_rows((build_rows + 7) / 8), kJoinBlock, 0, stream>>>((const uint8_t*)build[c].valid_bytes.p, build_rows, (uint8_t*)build_bits[c].p);
          count_launch();
          pp.build_cols[c].src_validity = (const uint8_t*)build_bits[c].p;
        }
        DBX_TRY(add_out(pp.build_cols[c], build_dtype[c], build_nullable[c] || prm.kind == DBX_JOIN_LEFT));
      }
      DBX_CUDA_TRY(err, cudaMemsetAsync(cursor.p, 0, 8, stream));
      if (can_part) {  // region by region: stream the region into L2, then probe the rows that hash into it
        for (int part = 0; part < n_part; ++part) {
          const int64_t m = part_offs[part + 1] - part_offs[part];
          if (m == 0) continue;
          const int64_t bytes = region * (int64_t)sizeof(JoinEntry);
          l2_prefetch_kernel<<<(int)std::min<int64_t>((bytes / 128 + 255) / 256, (int64_t)kNumSMs * 8), 256, 0, stream>>>(
              (const char*)table_buf.p + (int64_t)part * bytes, bytes);
          pp.row_begin = part_offs[part];
          pp.n_rows = m;
          launch_probe(pp, m);
          count_launch(2);
        }
      } else {
        launch_probe(pp, n);
        count_launch();
      }
      DBX_CUDA_TRY(err, cudaGetLastError());
      DBX_CUDA_TRY(err, cudaMemcpyAsync(host.p, cursor.p, 8, cudaMemcpyDeviceToHost, stream));
      DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
      const int64_t matches = (int64_t)*(unsigned long long*)host.p;
      if (matches > out_cap) {  // many-to-many: retry once with the exact size
        out_cap = matches;
        continue;
      }
      for (size_t i = 0; i < ob->cols.size(); ++i) {
        ob->cols[i].len = matches;
        if (valid_bytes[i]) {
          uint8_t* bits = nullptr;
          DBX_CUDA_TRY(err, pool_alloc(device, stream, (size_t)(matches + 7) / 8 + 8, (void**)&bits));
          ob->dev_allocs.push_back(bits);
          pack_bits_kernel<<<grid_rows((matches + 7) / 8 + 1), kJoinBlock, 0, stream>>>(valid_bytes[i], matches, bits);
          count_launch();
          ob->cols[i].validity = bits;
        }
      }
      DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
      if (matches > 0) outputs.push_back(std::move(ob));
      break;
    }
    DBX_TRY(timing_end());
    DBX_TRY(stager.end());
    return DBX_OK;
  }

  // JoinStream::next
  int32_t pull(int32_t out_mem, dbx_block* out, int32_t* has_block) override {
    if (next_out >= outputs.size()) { *has_block = 0; outputs.clear(); next_out = 0; return DBX_OK; }
    std::unique_ptr<OwnedBlock> ob = std::move(outputs[next_out++]);
    *has_block = 1;
    if (out_mem == DBX_MEM_DEVICE) return fill_owned_block(ob.release(), out);
    auto hb = std::make_unique<OwnedBlock>();
    hb->device = device;
    for (const dbx_column& dc : ob->cols) {
      dbx_column c = dc;
      c.mem = DBX_MEM_HOST;
      size_t bytes = (size_t)dc.len * dtype_size(dc.dtype);
      void* hp = nullptr;
      DBX_CUDA_TRY(err, pinned_alloc(bytes, &hp));
      hb->host_allocs.push_back(hp);
      if (bytes) DBX_CUDA_TRY(err, cudaMemcpyAsync(hp, dc.data, bytes, cudaMemcpyDeviceToHost, stream));
      c.data = hp;
      if (dc.validity) {
        size_t vb = (size_t)(dc.len + 7) / 8;
        void* hv = nullptr;
        DBX_CUDA_TRY(err, pinned_alloc(vb, &hv));
        hb->host_allocs.push_back(hv);
        if (vb) DBX_CUDA_TRY(err, cudaMemcpyAsync(hv, dc.validity, vb, cudaMemcpyDeviceToHost, stream));
        c.validity = (const uint8_t*)hv;
      }
      hb->cols.push_back(c);
    }
    DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
    return fill_owned_block(hb.release(), out);
  }

  int32_t reset() override {
    build_rows = 0;
    outputs.clear();
    next_out = 0;
    return DBX_OK;
  }
};

Op* make_join_op(const dbx_join_params* p, const int32_t* types, int32_t n, int device, int32_t* st) {
  auto* op = new JoinOp();
  *st = op->init(p, types, n, device);
  if (*st != DBX_OK) { g_create_error.set(op->err.msg); delete op; return nullptr; }
  return op;
}

}  // namespace dbx

using namespace dbx;

extern "C" int32_t dbx_join_probe(dbx_op* op, const dbx_block* block) {
  if (!op || !block) return DBX_ERR_INVALID;
  Op* o = reinterpret_cast<Op*>(op);
  if (o->kind != DBX_OP_JOIN) { o->err.set("dbx_join_probe: not a join operator"); return DBX_ERR_INVALID; }
  DBX_CUDA_TRY(o->err, cudaSetDevice(o->device));
  return static_cast<JoinOp*>(o)->probe(block);
}
