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
// 16-byte entries {key, build_row + 1}; a bucket is two entries = one 32-byte sector, so one
// 256-bit load yields key AND row id of two candidates (the reference reads an 8-byte header,
// then chases the entry chain).  The probe kernel streams the probe key column once, walks
// buckets until it sees an empty entry, and writes each joined row directly into the output
// columns at a position claimed with a warp-aggregated atomic: no (probe,build) index pairs are
// materialised and no second gather pass runs (the reference does DataBlock::take +
// take_column_vec).  Output row order is therefore unspecified — like the reference's when
// several threads build the chains — and results are compared as multisets.
// NULL keys never match (fixed_keys.rs: rows with a NULL key are skipped on both sides).
#include <algorithm>

#include "runtime.h"

namespace dbx {

namespace {

constexpr int kJoinBlock = 256;
constexpr int kMaxJoinCols = 16;

struct JoinEntry {
  uint64_t key;
  uint64_t row1;  // build row + 1; 0 = empty
};

struct JoinTableDev {
  JoinEntry* entries;  // cap entries, cap = 2 * n_buckets (power of two)
  int64_t cap;
};

// One column copied into the output for every match.
struct JoinColDev {
  const void* src;
  void* dst;
  const uint8_t* src_validity;  // may be null
  int64_t src_vbit_off;
  uint8_t* dst_valid;           // one byte per output row, or null
  int32_t size;                 // bytes per value
  int32_t pad;
};

struct JoinProbeParams {
  DevCol key;
  JoinTableDev table;
  JoinColDev probe_cols[kMaxJoinCols];
  JoinColDev build_cols[kMaxJoinCols];
  int32_t n_probe_cols, n_build_cols;
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
__global__ void join_build_kernel(const __grid_constant__ DevCol key, int64_t n_rows, int64_t row_base,
                                  const __grid_constant__ JoinTableDev t) {
  const int64_t mask = t.cap - 1;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    if (key.validity && !bit_test(key.validity, key.vbit_off + r)) continue;
    const uint64_t k = load_key(key, r);
    int64_t s = (int64_t)((agg_hash_u64(k) << 1) & (uint64_t)mask);  // first entry of the home bucket
    for (;;) {
      unsigned long long old = atomicCAS((unsigned long long*)&t.entries[s].row1, 0ULL, (unsigned long long)(row_base + r + 1));
      if (old == 0ULL) { t.entries[s].key = k; break; }
      s = (s + 1) & mask;
    }
  }
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
__global__ void __launch_bounds__(kJoinBlock) join_probe_kernel(const __grid_constant__ JoinProbeParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t nb_mask = (p.table.cap >> 1) - 1;
  const int64_t n_iter = (p.n_rows + (int64_t)gridDim.x * blockDim.x - 1) / ((int64_t)gridDim.x * blockDim.x);
  for (int64_t it = 0; it < n_iter; ++it) {
    __syncwarp();
    const int64_t r = it * (int64_t)gridDim.x * blockDim.x + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = r < p.n_rows && !(p.key.validity && !bit_test(p.key.validity, p.key.vbit_off + r));
    uint64_t k = 0;
    int64_t b = 0;
    if (live) {
      k = load_key(p.key, r);
      b = (int64_t)(agg_hash_u64(k) & (uint64_t)nb_mask);
    }
    // all lanes walk their bucket sequences in lock step; a lane retires at the first empty entry
    while (__any_sync(0xffffffffu, live)) {
      uint64_t k0 = 0, r0 = 0, k1 = 0, r1 = 0;
      if (live) {
        const JoinEntry* e = p.table.entries + 2 * b;
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(k0), "=l"(r0), "=l"(k1), "=l"(r1) : "l"(e));
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint64_t ek = half ? k1 : k0, er = half ? r1 : r0;
        const bool hit = live && er != 0 && ek == k;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(p.cursor, (unsigned long long)__popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (hit) {
            const int64_t pos = (int64_t)base + __popc(bal & ((1u << lane) - 1));
            if (pos < p.out_cap) {
              for (int c = 0; c < p.n_probe_cols; ++c) copy_value(p.probe_cols[c], r, pos);
              for (int c = 0; c < p.n_build_cols; ++c) copy_value(p.build_cols[c], (int64_t)er - 1, pos);
            }
          }
        }
      }
      if (live && (r0 == 0 || r1 == 0)) live = false;  // an empty entry ends the probe sequence
      b = (b + 1) & nb_mask;
    }
  }
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
  std::vector<std::unique_ptr<OwnedBlock>> outputs;  // joined blocks waiting to be pulled (device resident)
  size_t next_out = 0;

  // input_types = build schema (params.n_build_cols columns) followed by the probe schema.
  int32_t init(const dbx_join_params* p, const int32_t* types, int32_t n, int dev) {
    DBX_TRY(base_init(dev));
    prm = *p;
    if (p->kind != DBX_JOIN_INNER) { err.set("only INNER joins are built (SURVEY 8f.3 lists left/semi/anti as next)"); return DBX_ERR_UNSUPPORTED; }
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
    // (sign-extended if signed), the common super type of the reference's key cast.
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
      if (need > g.data.bytes) {  // grow, preserving the rows already appended
        DevBuf nb;
        DBX_CUDA_TRY(err, nb.ensure(std::max(need, g.data.bytes * 2)));
        if (build_rows) DBX_CUDA_TRY(err, cudaMemcpyAsync(nb.p, g.data.p, (size_t)build_rows * g.size, cudaMemcpyDeviceToDevice, stream));
        DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
        g.data = std::move(nb);
      }
      DBX_CUDA_TRY(err, cudaMemcpyAsync((char*)g.data.p + (size_t)build_rows * g.size, dc.data, (size_t)n * g.size, cudaMemcpyDeviceToDevice, stream));
      if (g.nullable) {
        const size_t vneed = (size_t)(build_rows + n);
        if (vneed > g.valid_bytes.bytes) {
          DevBuf nb;
          DBX_CUDA_TRY(err, nb.ensure(std::max(vneed, g.valid_bytes.bytes * 2)));
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
      JoinTableDev t{(JoinEntry*)table_buf.p, table_cap};
      join_build_kernel<<<grid_rows(build_rows), kJoinBlock, 0, stream>>>(key, build_rows, 0, t);
      count_launch();
      DBX_CUDA_TRY(err, cudaGetLastError());
      DBX_CUDA_TRY(err, cudaStreamSynchronize(stream));
    }
    return DBX_OK;
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
    DBX_CUDA_TRY(err, cudaEventRecord(ev_k0, stream));
    for (int attempt = 0; attempt < 2; ++attempt) {
      auto ob = std::make_unique<OwnedBlock>();
      ob->device = device;
      JoinProbeParams pp;
      memset(&pp, 0, sizeof(pp));
      pp.key = cols[prm.probe_key_col];
      pp.table = JoinTableDev{(JoinEntry*)table_buf.p, table_cap};
      pp.n_probe_cols = n_probe_cols;
      pp.n_build_cols = n_build_cols;
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
      for (int c = 0; c < n_build_cols; ++c) {
        pp.build_cols[c].src = build[c].data.p;
        if (build[c].nullable) {  // bytes -> use the byte array directly through a 1-byte "bitmap" trick: pack once
          DBX_CUDA_TRY(err, build_bits[c].ensure((size_t)(build_rows + 7) / 8 + 8));
          pack_bits_kernel<<<grid_rows((build_rows + 7) / 8), kJoinBlock, 0, stream>>>((const uint8_t*)build[c].valid_bytes.p, build_rows, (uint8_t*)build_bits[c].p);
          count_launch();
          pp.build_cols[c].src_validity = (const uint8_t*)build_bits[c].p;
        }
        DBX_TRY(add_out(pp.build_cols[c], build_dtype[c], build_nullable[c]));
      }
      DBX_CUDA_TRY(err, cudaMemsetAsync(cursor.p, 0, 8, stream));
      join_probe_kernel<<<grid_rows(n), kJoinBlock, 0, stream>>>(pp);
      count_launch();
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
    DBX_CUDA_TRY(err, cudaEventRecord(ev_k1, stream));
    timed = true;
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
