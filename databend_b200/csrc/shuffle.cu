// shuffle.cu — dbx_shuffle: hash-partitioned row shuffle between the GPUs of one box over peer
// memory (NVLink / NVSwitch through CUDA IPC), the exchange in front of a partitioned hash join
// (BASELINE configs[2]).
//
// Reference replaced: the hash scatter of join inputs between nodes,
//   src/query/service/src/servers/flight/v1/scatter/flight_scatter_hash.rs:86-125   (hash -> target)
//   src/query/service/src/servers/flight/v1/exchange/*                              (ship the blocks)
// There the scatter produces one DataBlock per target which Arrow Flight then serialises and
// sends.  Here ONE kernel partitions a device-resident block by the owner of its key and stores
// every row straight into the owner's receive region (peer memory): per 1024-row step a CTA
// counts its rows per owner, reserves one run per owner with a single atomic, lays the step out
// owner after owner in shared memory and writes each run with consecutive stores, so NVLink sees
// full lines.  No pack pass, no count exchange, no library all-to-all.  The last CTA publishes the
// row counts and a release flag per owner; the receiver waits for the flags with a one-warp kernel.
//
// Protocol (collective, like an all-to-all): every rank alternates send(block) / recv(); the
// blocks returned by recv() are views into the receive buffer and stay valid until this rank's
// next-but-one send (regions are double-buffered by round parity).  A rank must have finished
// reading the blocks of round e before it calls send for round e + 1.
#include <algorithm>
#include <vector>

#include "runtime.h"

namespace dbx {
namespace {

constexpr int kShufMaxRanks = 16;
constexpr int kShufMaxCols = 8;
constexpr int kShufRows = 4;  // rows per thread and step

struct ShuffleHeader {
  unsigned long long count[2][kShufMaxRanks];     // [parity][source]: rows the source wrote
  unsigned long long flag[kShufMaxRanks];         // [source]: last round the source completed
  unsigned long long overflow[2][kShufMaxRanks];  // [parity][source]: round in which the region was too small
  unsigned long long pad[16];
};

struct ShuffleSendParams {
  DevCol key;
  const void* src[kShufMaxCols];
  int32_t size[kShufMaxCols];
  void* peer_base[kShufMaxRanks];
  unsigned long long* cursors;  // [n_ranks], zeroed
  unsigned int* done;           // zeroed
  int64_t n_rows, region_rows;
  int64_t col_off[kShufMaxCols];  // byte offset of a column inside a (parity, source) region
  int64_t region_bytes;           // bytes of one (parity, source) region
  unsigned long long round;
  int32_t n_cols, n_ranks, rank, parity;
};

__device__ __forceinline__ char* shuffle_region(void* base, int n_ranks, int parity, int src, int64_t region_bytes) {
  return reinterpret_cast<char*>(base) + sizeof(ShuffleHeader) + (int64_t)(parity * n_ranks + src) * region_bytes;
}
__device__ __forceinline__ uint64_t shuf_load_key(const DevCol& c, int64_t row) {
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

__global__ void __launch_bounds__(256) shuffle_send_kernel(const __grid_constant__ ShuffleSendParams p) {
  __shared__ unsigned int s_cnt[kShufMaxRanks];
  __shared__ unsigned int s_off[kShufMaxRanks + 1];
  __shared__ unsigned long long s_base[kShufMaxRanks];
  __shared__ uint64_t s_val[256 * kShufRows];
  __shared__ int s_last;
  const int64_t step_rows = 256 * kShufRows;
  const int64_t n_steps = (p.n_rows + step_rows - 1) / step_rows;
#pragma unroll 1
  for (int64_t st = blockIdx.x; st < n_steps; st += gridDim.x) {
    if (threadIdx.x < kShufMaxRanks) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i0 = st * step_rows + threadIdx.x;
    int owner[kShufRows];
    unsigned int slot[kShufRows];
#pragma unroll
    for (int j = 0; j < kShufRows; ++j) {
      const int64_t i = i0 + (int64_t)j * 256;
      owner[j] = -1;
      slot[j] = 0;
      if (i < p.n_rows) {
        int o = part_owner(shuf_load_key(p.key, i), p.n_ranks);
        asm volatile("" : "+r"(o));  // (see partition.cu: keeps ptxas 12.9 from mis-folding the scaled index)
        owner[j] = o;
        slot[j] = atomicAdd(&s_cnt[o], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int o = 0;
      for (int r = 0; r < p.n_ranks; ++r) { s_off[r] = o; o += s_cnt[r]; }
      s_off[p.n_ranks] = o;
    }
    if (threadIdx.x < p.n_ranks && s_cnt[threadIdx.x])
      s_base[threadIdx.x] = atomicAdd(&p.cursors[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kShufRows; ++j)
      if (owner[j] >= 0) slot[j] += s_off[owner[j]];
    for (int c = 0; c < p.n_cols; ++c) {
      const int sz = p.size[c];
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kShufRows; ++j) {
        if (owner[j] < 0) continue;
        const int64_t i = i0 + (int64_t)j * 256;
        uint64_t v;
        if (sz == 8) v = ((const uint64_t*)p.src[c])[i];
        else if (sz == 4) v = ((const uint32_t*)p.src[c])[i];
        else if (sz == 2) v = ((const uint16_t*)p.src[c])[i];
        else v = ((const uint8_t*)p.src[c])[i];
        s_val[slot[j]] = v;
      }
      __syncthreads();
      for (int o = 0; o < p.n_ranks; ++o) {
        const unsigned int cnt = s_cnt[o];
        if (!cnt) continue;
        const int64_t base = (int64_t)s_base[o];
        // rows beyond the region are dropped here and reported through the overflow word
        const int64_t room = p.region_rows - base;
        const int64_t n_ok = room <= 0 ? 0 : (room < (int64_t)cnt ? room : (int64_t)cnt);
        char* dst = shuffle_region(p.peer_base[o], p.n_ranks, p.parity, p.rank, p.region_bytes) + p.col_off[c] + base * sz;
        const uint64_t* from = s_val + s_off[o];
        for (int64_t t = threadIdx.x; t < n_ok; t += 256) {
          const uint64_t v = from[t];
          if (sz == 8) ((uint64_t*)dst)[t] = v;
          else if (sz == 4) ((uint32_t*)dst)[t] = (uint32_t)v;
          else if (sz == 2) ((uint16_t*)dst)[t] = (uint16_t)v;
          else ((uint8_t*)dst)[t] = (uint8_t)v;
        }
      }
    }
    __syncthreads();
  }
  // publish: the last CTA to finish writes the row counts and then the release flags
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(p.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last && threadIdx.x < p.n_ranks) {
    __threadfence_system();
    const unsigned long long cnt = atomicAdd(&p.cursors[threadIdx.x], 0ULL);
    ShuffleHeader* h = reinterpret_cast<ShuffleHeader*>(p.peer_base[threadIdx.x]);
    const bool over = (int64_t)cnt > p.region_rows;
    h->count[p.parity][p.rank] = over ? (unsigned long long)p.region_rows : cnt;
    if (over) h->overflow[p.parity][p.rank] = p.round;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&h->flag[p.rank]), "l"(p.round) : "memory");
  }
}

__device__ __forceinline__ unsigned long long shuf_globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// one warp: wait until every source has released its region of `round`; copies the counts out
__global__ void __launch_bounds__(32) shuffle_wait_kernel(void* base, int n_ranks, int parity, unsigned long long round, long long spin_limit_ns,
                                                          unsigned long long* out /* [n_ranks] counts, [16] timeout, [17] overflow, [18] wait ns */) {
  ShuffleHeader* h = reinterpret_cast<ShuffleHeader*>(base);
  const unsigned long long t0 = shuf_globaltimer_ns();
  bool fail = false;
  if ((int)threadIdx.x < n_ranks) {
    while (true) {
      unsigned long long f;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(&h->flag[threadIdx.x]) : "memory");
      if (f >= round) break;
      if ((long long)(shuf_globaltimer_ns() - t0) > spin_limit_ns) { fail = true; break; }
      __nanosleep(100);
    }
    if (!fail) {
      out[threadIdx.x] = h->count[parity][threadIdx.x];
      if (h->overflow[parity][threadIdx.x] == round) out[17] = 1;
    }
  }
  const unsigned any_fail = __ballot_sync(0xffffffffu, fail);
  if (threadIdx.x == 0) {
    out[16] = any_fail ? 1 : 0;
    out[18] = shuf_globaltimer_ns() - t0;
  }
}

}  // namespace
}  // namespace dbx

using namespace dbx;

struct dbx_shuffle {
  ErrorSink err;
  int device = 0, rank = 0, n_ranks = 1, n_cols = 0, key_col = 0;
  int32_t dtype[kShufMaxCols] = {};
  int64_t region_rows = 0, region_bytes = 0;
  int64_t col_off[kShufMaxCols] = {};
  DevBuf recv, scratch, wait_out;
  PinnedBuf host;
  void* peer_base[kShufMaxRanks] = {};
  bool peer_is_ipc[kShufMaxRanks] = {};
  bool connected = false;
  unsigned long long round = 0, received = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  long long spin_limit_ns = 5000LL * 1000 * 1000;
  float last_send_ms = 0.f;
  double last_wait_ms = 0.0;
  size_t recv_bytes() const { return sizeof(ShuffleHeader) + (size_t)2 * n_ranks * region_bytes; }
};

extern "C" {

const char* dbx_shuffle_last_error(const dbx_shuffle* s) { return s ? s->err.msg.c_str() : g_create_error.msg.c_str(); }

int32_t dbx_shuffle_create(int32_t device, int32_t rank, int32_t n_ranks, const int32_t* col_types, int32_t n_cols, int32_t key_col,
                           int64_t region_rows, dbx_shuffle** out, void* ipc_handle_out) {
  ErrorSink& err = g_create_error;
  if (!out || !col_types || n_ranks < 1 || n_ranks > kShufMaxRanks || rank < 0 || rank >= n_ranks || n_cols < 1 || n_cols > kShufMaxCols ||
      key_col < 0 || key_col >= n_cols || region_rows < 1) { err.set("dbx_shuffle_create: bad argument"); return DBX_ERR_INVALID; }
  int32_t ndev = 0;
  DBX_TRY(dbx_device_count(&ndev));
  if (device < 0 || device >= ndev) { err.set("dbx_shuffle_create: device index out of range"); return DBX_ERR_INVALID; }
  std::unique_ptr<dbx_shuffle> s(new dbx_shuffle());
  s->device = device; s->rank = rank; s->n_ranks = n_ranks; s->n_cols = n_cols; s->key_col = key_col;
  s->region_rows = region_rows;
  int64_t off = 0;
  for (int c = 0; c < n_cols; ++c) {
    const int dt = col_types[c] & 0xFF;
    if (dtype_size(dt) == 0 || (col_types[c] & DBX_NULLABLE)) { err.set("dbx_shuffle_create: columns must be non-nullable fixed-width numeric columns"); return DBX_ERR_UNSUPPORTED; }
    s->dtype[c] = dt;
    s->col_off[c] = off;
    off += ((region_rows * dtype_size(dt) + 255) / 256) * 256;
  }
  if (dtype_class(s->dtype[key_col]) == VC_FLT) { err.set("dbx_shuffle_create: the key must be an integer column"); return DBX_ERR_UNSUPPORTED; }
  s->region_bytes = off;
  if (getenv("DBX_EXCH_SPIN_MS")) s->spin_limit_ns = atoll(getenv("DBX_EXCH_SPIN_MS")) * 1000000LL;
  DBX_CUDA_TRY(err, cudaSetDevice(device));
  DBX_CUDA_TRY(err, s->recv.ensure(s->recv_bytes()));
  DBX_CUDA_TRY(err, cudaMemset(s->recv.p, 0, sizeof(ShuffleHeader)));
  DBX_CUDA_TRY(err, s->scratch.ensure(8 * (kShufMaxRanks + 2)));
  DBX_CUDA_TRY(err, s->wait_out.ensure(8 * 32));
  DBX_CUDA_TRY(err, s->host.ensure(8 * 32));
  DBX_CUDA_TRY(err, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  DBX_CUDA_TRY(err, cudaEventCreate(&s->ev0));
  DBX_CUDA_TRY(err, cudaEventCreate(&s->ev1));
  if (ipc_handle_out) {
    cudaIpcMemHandle_t hd;
    DBX_CUDA_TRY(err, cudaIpcGetMemHandle(&hd, s->recv.p));
    memcpy(ipc_handle_out, &hd, 64);
  }
  *out = s.release();
  return DBX_OK;
}

int32_t dbx_shuffle_local_buffer(dbx_shuffle* s, void** base) {
  if (!s || !base) return DBX_ERR_INVALID;
  *base = s->recv.p;
  return DBX_OK;
}

int32_t dbx_shuffle_connect(dbx_shuffle* s, const void* all_handles, void* const* same_process_ptrs) {
  if (!s || (!all_handles && !same_process_ptrs)) return DBX_ERR_INVALID;
  DBX_CUDA_TRY(s->err, cudaSetDevice(s->device));
  for (int r = 0; r < s->n_ranks; ++r) {
    if (r == s->rank) { s->peer_base[r] = s->recv.p; continue; }
    if (same_process_ptrs) { s->peer_base[r] = same_process_ptrs[r]; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char*)all_handles + (size_t)r * 64, 64);
    void* p = nullptr;
    DBX_CUDA_TRY(s->err, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    s->peer_base[r] = p;
    s->peer_is_ipc[r] = true;
  }
  s->connected = true;
  return DBX_OK;
}

/* One round: partition `block` (device-resident columns matching the schema; at most region_rows
 * rows) by the owner of its key and store every row into the owners' regions.  Enqueued on the
 * shuffle's stream, no host synchronisation. */
int32_t dbx_shuffle_send(dbx_shuffle* s, const dbx_block* block) {
  if (!s || !block) return DBX_ERR_INVALID;
  ErrorSink& err = s->err;
  if (!s->connected) { err.set("shuffle: send before connect"); return DBX_ERR_STATE; }
  if (s->round != s->received) { err.set("shuffle: send called twice without recv (rounds alternate send / recv)"); return DBX_ERR_STATE; }
  if (block->num_cols != s->n_cols) { err.set("shuffle: block does not match the schema"); return DBX_ERR_INVALID; }
  if (block->num_rows > s->region_rows) { err.set("shuffle: block larger than a receive region (send it in pieces of region_rows)"); return DBX_ERR_INVALID; }
  DBX_CUDA_TRY(err, cudaSetDevice(s->device));
  ShuffleSendParams p;
  memset(&p, 0, sizeof(p));
  for (int c = 0; c < s->n_cols; ++c) {
    const dbx_column& col = block->cols[c];
    if (col.dtype != s->dtype[c] || col.len != block->num_rows || col.is_const || col.validity || (block->num_rows && col.mem != DBX_MEM_DEVICE)) {
      err.set("shuffle: columns must be device-resident, non-nullable, non-const and match the schema");
      return DBX_ERR_INVALID;
    }
    p.src[c] = col.data;
    p.size[c] = dtype_size(col.dtype);
    p.col_off[c] = s->col_off[c];
  }
  s->round += 1;
  p.key.data = block->cols[s->key_col].data;
  p.key.dtype = s->dtype[s->key_col];
  for (int r = 0; r < s->n_ranks; ++r) p.peer_base[r] = s->peer_base[r];
  p.cursors = (unsigned long long*)s->scratch.p;
  p.done = (unsigned int*)((unsigned long long*)s->scratch.p + kShufMaxRanks);
  p.n_rows = block->num_rows; p.region_rows = s->region_rows; p.region_bytes = s->region_bytes;
  p.round = s->round; p.n_cols = s->n_cols; p.n_ranks = s->n_ranks; p.rank = s->rank; p.parity = (int)(s->round & 1);
  DBX_CUDA_TRY(err, cudaEventRecord(s->ev0, s->stream));
  DBX_CUDA_TRY(err, cudaMemsetAsync(s->scratch.p, 0, 8 * (kShufMaxRanks + 2), s->stream));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((block->num_rows + 1023) / 1024, (int64_t)kNumSMs * 8));
  shuffle_send_kernel<<<grid, 256, 0, s->stream>>>(p);
  count_launch();
  DBX_CUDA_TRY(err, cudaGetLastError());
  DBX_CUDA_TRY(err, cudaEventRecord(s->ev1, s->stream));
  return DBX_OK;
}

/* Wait (device side, then one host synchronisation for the row counts) until every source's
 * region of the current round is complete; blocks[r] / cols[r * n_cols ..] describe the rows
 * received from source r (device-resident views into the receive buffer). */
int32_t dbx_shuffle_recv(dbx_shuffle* s, dbx_block* blocks, dbx_column* cols) {
  if (!s || !blocks || !cols) return DBX_ERR_INVALID;
  ErrorSink& err = s->err;
  if (s->received >= s->round) { err.set("shuffle: recv without a send in this round"); return DBX_ERR_STATE; }
  DBX_CUDA_TRY(err, cudaSetDevice(s->device));
  const int parity = (int)(s->round & 1);
  DBX_CUDA_TRY(err, cudaMemsetAsync(s->wait_out.p, 0, 8 * 32, s->stream));
  shuffle_wait_kernel<<<1, 32, 0, s->stream>>>(s->recv.p, s->n_ranks, parity, s->round, s->spin_limit_ns, (unsigned long long*)s->wait_out.p);
  count_launch();
  DBX_CUDA_TRY(err, cudaGetLastError());
  DBX_CUDA_TRY(err, cudaMemcpyAsync(s->host.p, s->wait_out.p, 8 * 32, cudaMemcpyDeviceToHost, s->stream));
  DBX_CUDA_TRY(err, cudaStreamSynchronize(s->stream));
  cudaEventElapsedTime(&s->last_send_ms, s->ev0, s->ev1);
  const unsigned long long* h = (const unsigned long long*)s->host.p;
  s->last_wait_ms = (double)h[18] * 1e-6;
  s->received = s->round;
  if (h[16]) { err.set("shuffle: timed out waiting for a peer rank's rows"); return DBX_ERR_STATE; }
  if (h[17]) { err.set("shuffle: a peer had more rows for this rank than the receive region holds"); return DBX_ERR_OOM; }
  for (int r = 0; r < s->n_ranks; ++r) {
    const int64_t n = (int64_t)h[r];
    char* region = reinterpret_cast<char*>(s->recv.p) + sizeof(ShuffleHeader) + (int64_t)(parity * s->n_ranks + r) * s->region_bytes;
    dbx_block& b = blocks[r];
    memset(&b, 0, sizeof(b));
    b.num_rows = n; b.num_cols = s->n_cols; b.cols = cols + (size_t)r * s->n_cols;
    for (int c = 0; c < s->n_cols; ++c) {
      dbx_column& col = b.cols[c];
      memset(&col, 0, sizeof(col));
      col.dtype = s->dtype[c]; col.mem = DBX_MEM_DEVICE; col.len = n; col.data = region + s->col_off[c];
    }
  }
  return DBX_OK;
}

/* Device time of the last send kernel (ms) and the receiver's wait for the peers' flags (ms). */
int32_t dbx_shuffle_last_ms(dbx_shuffle* s, float* send_ms, float* wait_ms) {
  if (!s) return DBX_ERR_INVALID;
  if (send_ms) *send_ms = s->last_send_ms;
  if (wait_ms) *wait_ms = (float)s->last_wait_ms;
  return DBX_OK;
}

int32_t dbx_shuffle_destroy(dbx_shuffle* s) {
  if (!s) return DBX_OK;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < s->n_ranks; ++r)
    if (s->peer_is_ipc[r] && s->peer_base[r]) cudaIpcCloseMemHandle(s->peer_base[r]);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return DBX_OK;
}

}  // extern "C"
