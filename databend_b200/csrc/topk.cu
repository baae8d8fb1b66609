// topk.cu — DBX_OP_TOPK: `ORDER BY key [ASC|DESC] [NULLS FIRST|LAST] LIMIT k` as a streaming
// device top-k.
//
// Reference pipeline replaced (paths relative to /root/reference):
//   TransformSortPartial (per block sort + limit)   src/query/pipeline/transforms/src/processors/transforms/sorts/sort_partial.rs:24-60
//     DataBlock::sort_with_type / SortCompare       src/query/expression/src/kernels/sort.rs:91-111, sort_compare.rs:197-296
//   limit-aware merge                               sorts/sort_merge*.rs, sorts/core/merger.rs
//   fused TopN with a runtime boundary filter       src/query/service/src/pipelines/processors/transforms/top_n/transform_partial_top_n.rs:73-130
//
// B200 design: the column is read ONCE (8 B/row, 256-bit streaming loads).  Each key is mapped
// to an order-preserving u64 (OrderedFloat order: NaN greatest, -0 == +0,
// src/common/base/src/base/ordered_float.rs:147-201); a row survives only if it beats the
// current boundary (the k-th best key so far — the reference's TopN boundary filter), and
// survivors are appended to a small candidate list with one warp-aggregated atomic.  The
// candidate list is periodically cut back to k (radix sort, tiny), tightening the boundary.
// Ties are broken by ascending row id, as in the oracle.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

#include "runtime.h"

namespace dbx {

namespace {

constexpr int kTopkBlock = 256;
constexpr int64_t kMaxChunk = 1LL << 30;

struct TopkDev {
  uint64_t* ord;      // order-preserving image (smaller = earlier in the output)
  uint64_t* rowid;    // global row ordinal
  uint64_t* bits;     // original value bits (widened to 64)
  unsigned long long* count;     // appended candidates
  unsigned long long* n_null;    // NULL rows seen
  uint64_t* null_rowid;          // first rows with NULL key (up to k, by append order)
  int64_t cap;
  int64_t k;
};

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

// One pass over `n` rows of the key column.  boundary: only ord <= boundary can still be in
// the top k.  FAST: 8-byte column, 32 B aligned, no validity -> one 256-bit load per 4 rows.
template <bool FAST>
__global__ void __launch_bounds__(kTopkBlock) topk_scan_kernel(const __grid_constant__ DevCol col, int64_t n,
                                                               int64_t row_base, int cls, int asc, uint64_t boundary,
                                                               const __grid_constant__ TopkDev t) {
  const uint64_t pol = make_policy_evict_first();
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (n + kTopkBlock * 4 - 1) / (kTopkBlock * 4);
  u64x4 next_q;
  next_q.x = next_q.y = next_q.z = next_q.w = 0;
  bool have_next = false;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncwarp();
    const int64_t r0 = tile * (kTopkBlock * 4) + 4 * (int64_t)threadIdx.x;
    uint64_t v[4];
    uint32_t valid = 0, inr = 0;
    if (FAST && r0 + 4 <= n) {
      u64x4 q = have_next ? next_q : ld_stream_256((const char*)col.data + r0 * 8);
      // keep a second tile in flight: one 32-byte load per thread does not cover the HBM latency
      const int64_t rn = (tile + gridDim.x) * (kTopkBlock * 4) + 4 * (int64_t)threadIdx.x;
      have_next = tile + gridDim.x < n_tiles && rn + 4 <= n;
      if (have_next) next_q = ld_stream_256((const char*)col.data + rn * 8);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      valid = inr = 0xF;
    } else {
      have_next = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = 0;
        if (r0 + j < n) {
          inr |= 1u << j;
          bool ok = !col.validity || bit_test(col.validity, col.vbit_off + r0 + j);
          if (ok) { v[j] = load_widened(col, r0 + j, pol); valid |= 1u << j; }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = (inr >> j) & 1, ok = (valid >> j) & 1;
      const uint64_t o = key_to_ord(v[j], cls, asc != 0);
      const bool keep = in && ok && o <= boundary;
      const bool is_null = in && !ok;
      const uint32_t bal = __ballot_sync(0xffffffffu, keep);
      if (bal) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(t.count, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
          unsigned long long pos = base + __popc(bal & ((1u << lane) - 1));
          if ((int64_t)pos < t.cap) {
            t.ord[pos] = o;
            t.rowid[pos] = (uint64_t)(row_base + r0 + j);
            t.bits[pos] = v[j];
          }
        }
      }
      const uint32_t nb = __ballot_sync(0xffffffffu, is_null);
      if (nb) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(t.n_null, (unsigned long long)__popc(nb));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_null) {
          unsigned long long pos = base + __popc(nb & ((1u << lane) - 1));
          if ((int64_t)pos < t.k) t.null_rowid[pos] = (uint64_t)(row_base + r0 + j);
        }
      }
    }
  }
}

__global__ void iota_kernel(uint32_t* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}
__global__ void gather_u64_kernel(const uint64_t* src, const uint32_t* idx, uint64_t* dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
__global__ void narrow_store_kernel(const uint64_t* bits, int64_t n, int dtype, void* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t b = bits[i];
    switch (dtype) {
      case DBX_I8: case DBX_U8: ((uint8_t*)out)[i] = (uint8_t)b; break;
      case DBX_I16: case DBX_U16: ((uint16_t*)out)[i] = (uint16_t)b; break;
      case DBX_I32: case DBX_U32: ((uint32_t*)out)[i] = (uint32_t)b; break;
      case DBX_F32: ((float*)out)[i] = (float)__longlong_as_double((long long)b); break;
      default: ((uint64_t*)out)[i] = b; break;
    }
  }
}

inline int grid_1d(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8)); }

}  // namespace

class TopkOp : public Op {
 public:
  dbx_topk_params prm;
  int n_cols = 0;
  int key_dtype = 0;
  bool key_nullable = false;
  int cls = 0;
  Stager stager;
  DevBuf ord, rowid, bits, counters, null_rowid;
  DevBuf s_ord, s_rowid, s_bits, idx_a, idx_b, key_tmp, key_tmp2, cub_tmp;
  PinnedBuf host;
  int64_t cap = 0;
  int64_t n_cand = 0;        // exact candidate count (host knowledge)
  uint64_t boundary = ~0ULL; // ord of the k-th best so far
  int64_t rows_seen = 0;
  int64_t next_chunk = 0;
  int64_t n_null_seen = 0;
  std::unique_ptr<OwnedBlock> result;
  bool pulled = false;

  int32_t init(const dbx_topk_params* p, const int32_t* types, int32_t n, int dev) {
    DBX_TRY(base_init(dev));
    prm = *p;
    n_cols = n;
    if (p->key_col < 0 || p->key_col >= n) { err.set("top-k: key column outside the input schema"); return DBX_ERR_INVALID; }
    if (p->limit <= 0 || p->limit > (1 << 22)) { err.set("top-k: limit must be in [1, 4194304]"); return DBX_ERR_UNSUPPORTED; }
    key_dtype = types[p->key_col] & 0xFF;
    key_nullable = (types[p->key_col] & DBX_NULLABLE) != 0;
    if (dtype_size(key_dtype) == 0) { err.set("top-k: key must be a numeric column"); return DBX_ERR_UNSUPPORTED; }
    cls = key_dtype == DBX_U64 ? VC_UINT : (dtype_class(key_dtype) == VC_FLT ? VC_FLT : VC_INT);
    cap = std::max<int64_t>(1 << 16, 16 * p->limit);
    DBX_TRY(stager.init(dev, stream, &err));
    DBX_CUDA_TRY(err, ord.ensure(cap * 8));
    DBX_CUDA_TRY(err, rowid.ensure(cap * 8));
    DBX_CUDA_TRY(err, bits.ensure(cap * 8));
    DBX_CUDA_TRY(err, s_ord.ensure(cap * 8));
    DBX_CUDA_TRY(err, s_rowid.ensure(cap * 8));
    DBX_CUDA_TRY(err, s_bits.ensure(cap * 8));
    DBX_CUDA_TRY(err, idx_a.ensure(cap * 4));
    DBX_CUDA_TRY(err, idx_b.ensure(cap * 4));
    DBX_CUDA_TRY(err, key_tmp.ensure(cap * 8));
    DBX_CUDA_TRY(err, key_tmp2.ensure(cap * 8));
    DBX_CUDA_TRY(err, null_rowid.ensure(p->limit * 8));
    DBX_CUDA_TRY(err, counters.ensure(64));
    DBX_CUDA_TRY(err, host.ensure(64));
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)cap, 0, 64, stream);
    DBX_CUDA_TRY(err, cub_tmp.ensure(tmp + 256));
    return reset();
  }

  int32_t reset() override {
    DBX_CUDA_TRY(err, cudaMemsetAsync(counters.p, 0, 64, stream));
    n_cand = 0;
    boundary = ~0ULL;
    rows_seen = 0;
    n_null_seen = 0;
    next_chunk = std::max<int64_t>(cap / 4, 1024);
    result.reset();
    pulled = false;
    return DBX_OK;
  }

  TopkDev view() const {
    TopkDev t;
    t.ord = (uint64_t*)ord.p; t.rowid = (uint64_t*)rowid.p; t.bits = (uint64_t*)bits.p;
    t.count = (unsigned long long*)counters.p;
    t.n_null = (unsigned long long*)counters.p + 1;
    t.null_rowid = (uint64_t*)null_rowid.p;
    t.cap = cap; t.k = prm.limit;
    return t;
  }

  // Sort the first n candidates by (ord, rowid) into s_ord/s_rowid/s_bits (two stable passes).
  int32_t sort_candidates(int64_t n) {
    if (n == 0) return DBX_OK;
    size_t tmp = cub_tmp.bytes;
    iota_kernel<<<grid_1d(n), 256, 0, stream>>>((uint32_t*)idx_a.p, n);
    count_launch();
    DBX_CUDA_TRY(err, cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, (const uint64_t*)rowid.p, (uint64_t*)key_tmp.p,
                                                      (const uint32_t*)idx_a.p, (uint32_t*)idx_b.p, (int)n, 0, 64, stream));
    gather_u64_kernel<<<grid_1d(n), 256, 0, stream>>>((const uint64_t*)ord.p, (const uint32_t*)idx_b.p, (uint64_t*)key_tmp.p, n);
    count_launch();
    tmp = cub_tmp.bytes;
    DBX_CUDA_TRY(err, cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, (const uint64_t*)key_tmp.p, (uint64_t*)key_tmp2.p,
                                                      (const uint32_t*)idx_b.p, (uint32_t*)idx_a.p, (int)n, 0, 64, stream));
    gather_u64_kernel<<<grid_1d(n), 256, 0, stream>>>((const uint64_t*)ord.p, (const uint32_t*)idx_a.p, (uint64_t*)s_ord.p, n);
    gather_u64_kernel<<<grid_1d(n), 256, 0, stream>>>((const uint64_t*)rowid.p, (const uint32_t*)idx_a.p, (uint64_t*)s_rowid.p, n);
    gather_u64_kernel<<<grid_1d(n), 256, 0, stream>>>((const uint64_t*)bits.p, (const uint32_t*)idx_a.p, (uint64_t*)s_bits.p, n);
    count_launch(5);
    DBX_CUDA_TRY(err, cudaGetLastError());
    return DBX_OK;
  }

  // Cut the candidate list back to the k best and tighten the boundary.
  int32_t compact() {
    if (n_cand <= prm.limit) return DBX_OK;
    DBX_TRY(sort_candidates(n_cand));
    const int64_t k = prm.limit;
    DBX_CUDA_TRY(err, cudaMemcpyAsync(ord.p, s_ord.p, k * 8, cudaMemcpyDeviceToDevice, stream));
    DBX_CUDA_TRY(err, cudaMemcpyAsync(rowid.p, s_rowid.p, k * 8, cudaMemcpyDeviceToDevice, stream));
    DBX_CUDA_TRY(err, cudaMemcpyAsync(bits.p, s_bits.p, k * 8, cudaMemcpyDeviceToDevice, stream));
    DBX_CUDA_TRY(err, cudaMemcpyAsync(host.p, (uint64_t*)s_ord.p + (k - 1), 8, cudaMemcpyDeviceToHost, stream));
    unsigned long long kk = (unsigned long long)k;
    DBX_CUDA_TRY(err, cudaMemcpyAsync(counters.p, &kk, 8, cudaMemcpyHostToDevice, stream));
    DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
    boundary = *(uint64_t*)host.p;
    n_cand = k;
    return DBX_OK;
  }

  int32_t push(const dbx_block* b) override {
    if (b->num_cols != n_cols) { err.set("push: block column count differs from the operator's input schema"); return DBX_ERR_INVALID; }
    const dbx_column& kc = b->cols[prm.key_col];
    if (kc.dtype != key_dtype || kc.len != b->num_rows) { err.set("push: key column does not match the input schema"); return DBX_ERR_INVALID; }
    const int64_t n = b->num_rows;
    if (n == 0) return DBX_OK;
    if (kc.is_const) { err.set("top-k over a constant key column is not supported"); return DBX_ERR_UNSUPPORTED; }
    DevCol col;
    DBX_TRY(stager.begin());
    DBX_TRY(stager.stage(kc, 0, &col));
    DBX_TRY(timing_begin());
    const int esz = dtype_size(key_dtype);
    int64_t done = 0;
    while (done < n) {
      // a chunk never appends more than it has rows: keep (candidates + chunk) within the list
      int64_t room = cap - n_cand;
      if (room < cap / 4) { DBX_TRY(compact()); room = cap - n_cand; }
      int64_t m = std::min<int64_t>({n - done, next_chunk, kMaxChunk});
      const bool guaranteed = m <= room;
      const int64_t nulls_before = n_null_seen;
      DevCol c = col;
      c.data = (const char*)col.data + done * esz;
      if (c.validity) c.vbit_off += done;
      const bool fast = esz == 8 && !c.validity && ((reinterpret_cast<uintptr_t>(c.data) & 31) == 0);
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((m + 1023) / 1024, (int64_t)kNumSMs * 8));
      if (fast) topk_scan_kernel<true><<<grid, kTopkBlock, 0, stream>>>(c, m, rows_seen + done, cls, prm.asc, boundary, view());
      else topk_scan_kernel<false><<<grid, kTopkBlock, 0, stream>>>(c, m, rows_seen + done, cls, prm.asc, boundary, view());
      count_launch();
      DBX_CUDA_TRY(err, cudaGetLastError());
      DBX_CUDA_TRY(err, cudaMemcpyAsync(host.p, counters.p, 16, cudaMemcpyDeviceToHost, stream));
      DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
      const int64_t cnt = (int64_t)((unsigned long long*)host.p)[0];
      n_null_seen = (int64_t)((unsigned long long*)host.p)[1];
      if (cnt > cap) {  // more survivors than the list holds: drop this chunk's appends, tighten, retry smaller
        if (guaranteed) { err.set("internal: top-k candidate overflow"); return DBX_ERR_CUDA; }
        unsigned long long back[2] = {(unsigned long long)n_cand, (unsigned long long)nulls_before};
        DBX_CUDA_TRY(err, cudaMemcpyAsync(counters.p, back, 16, cudaMemcpyHostToDevice, stream));
        DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
        n_null_seen = nulls_before;
        DBX_TRY(compact());
        next_chunk = std::max<int64_t>(std::min<int64_t>(m / 4, cap - n_cand), 1024);
        continue;
      }
      n_cand = cnt;
      done += m;
      // the boundary tightens as rows are seen: later chunks can be geometrically larger
      // Cut back after every chunk that left more than 2k candidates: the boundary then reflects
      // every row seen so far, so the next (8x larger) chunk adds about 8k survivors instead of
      // overflowing the list and being replayed in smaller pieces.
      if (n_cand > 2 * prm.limit || n_cand > cap / 2) DBX_TRY(compact());
      next_chunk = std::min<int64_t>(next_chunk * 8, kMaxChunk);
    }
    rows_seen += n;
    DBX_TRY(timing_end());
    DBX_TRY(stager.end());
    return DBX_OK;
  }

  int32_t finish() override {
    const int64_t k = prm.limit;
    DBX_TRY(sort_candidates(n_cand));
    DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
    const int64_t n_valid = std::min<int64_t>(n_cand, k);
    const int64_t n_nulls = std::min<int64_t>(n_null_seen, k);
    int64_t take_null, take_valid;
    if (prm.nulls_first) { take_null = n_nulls; take_valid = std::min<int64_t>(k - take_null, n_valid); }
    else { take_valid = n_valid; take_null = std::min<int64_t>(k - take_valid, n_nulls); }
    const int64_t n_out = take_null + take_valid;
    // NULL row ids were appended in arbitrary order: keep the smallest ones (ties by row id)
    std::vector<uint64_t> null_ids((size_t)n_nulls);
    if (n_nulls) {
      DBX_CUDA_TRY(err, cudaMemcpy(null_ids.data(), null_rowid.p, n_nulls * 8, cudaMemcpyDeviceToHost));
      std::sort(null_ids.begin(), null_ids.end());
    }
    std::vector<uint64_t> h_rowid((size_t)n_valid), h_bits((size_t)n_valid);
    if (n_valid) {
      DBX_CUDA_TRY(err, cudaMemcpy(h_rowid.data(), s_rowid.p, n_valid * 8, cudaMemcpyDeviceToHost));
      DBX_CUDA_TRY(err, cudaMemcpy(h_bits.data(), s_bits.p, n_valid * 8, cudaMemcpyDeviceToHost));
    }
    // output block: [key (original dtype, nullable), row_id Int64], assembled in pinned memory
    auto ob = std::make_unique<OwnedBlock>();
    ob->device = device;
    const int esz = dtype_size(key_dtype);
    void *hk = nullptr, *hr = nullptr, *hv = nullptr;
    DBX_CUDA_TRY(err, pinned_alloc(std::max<int64_t>(1, n_out * esz), &hk));
    ob->host_allocs.push_back(hk);
    DBX_CUDA_TRY(err, pinned_alloc(std::max<int64_t>(1, n_out * 8), &hr));
    ob->host_allocs.push_back(hr);
    DBX_CUDA_TRY(err, pinned_alloc((size_t)(n_out + 7) / 8 + 1, &hv));
    ob->host_allocs.push_back(hv);
    memset(hv, 0, (size_t)(n_out + 7) / 8 + 1);
    auto put = [&](int64_t o, uint64_t b, uint64_t rid, bool valid) {
      switch (key_dtype) {
        case DBX_I8: case DBX_U8: ((uint8_t*)hk)[o] = (uint8_t)b; break;
        case DBX_I16: case DBX_U16: ((uint16_t*)hk)[o] = (uint16_t)b; break;
        case DBX_I32: case DBX_U32: ((uint32_t*)hk)[o] = (uint32_t)b; break;
        case DBX_F32: { double d; memcpy(&d, &b, 8); ((float*)hk)[o] = (float)d; break; }
        default: ((uint64_t*)hk)[o] = b; break;
      }
      ((int64_t*)hr)[o] = (int64_t)rid;
      if (valid) ((uint8_t*)hv)[o >> 3] |= (uint8_t)(1u << (o & 7));
    };
    int64_t o = 0;
    if (prm.nulls_first) for (int64_t i = 0; i < take_null; ++i) put(o++, 0, null_ids[i], false);
    for (int64_t i = 0; i < take_valid; ++i) put(o++, h_bits[i], h_rowid[i], true);
    if (!prm.nulls_first) for (int64_t i = 0; i < take_null; ++i) put(o++, 0, null_ids[i], false);
    dbx_column kcol;
    memset(&kcol, 0, sizeof(kcol));
    kcol.dtype = key_dtype; kcol.mem = DBX_MEM_HOST; kcol.len = n_out; kcol.data = hk;
    if (key_nullable) { kcol.validity = (const uint8_t*)hv; kcol.null_count = take_null; }
    dbx_column rcol;
    memset(&rcol, 0, sizeof(rcol));
    rcol.dtype = DBX_I64; rcol.mem = DBX_MEM_HOST; rcol.len = n_out; rcol.data = hr;
    ob->cols.push_back(kcol);
    ob->cols.push_back(rcol);
    result = std::move(ob);
    return DBX_OK;
  }

  int32_t pull(int32_t out_mem, dbx_block* out, int32_t* has_block) override {
    if (!finished) { err.set("pull before finish"); return DBX_ERR_STATE; }
    if (pulled || !result) { *has_block = 0; return DBX_OK; }
    if (out_mem != DBX_MEM_HOST) { err.set("top-k results are k rows: host output only"); return DBX_ERR_UNSUPPORTED; }
    pulled = true;
    *has_block = 1;
    return fill_owned_block(result.release(), out);
  }
};

Op* make_topk_op(const dbx_topk_params* p, const int32_t* types, int32_t n, int device, int32_t* st) {
  auto* op = new TopkOp();
  *st = op->init(p, types, n, device);
  if (*st != DBX_OK) { g_create_error.set(op->err.msg); delete op; return nullptr; }
  return op;
}

}  // namespace dbx
