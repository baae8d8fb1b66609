// synth.cu — deterministic counter-based synthetic columns (tests / bench only).
// The host twin is orc_synth_fill in oracle/dbx_oracle.c: both must produce identical bits.
#include "common.cuh"

namespace dbx {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

__global__ void synth_fill_kernel(int kind, uint64_t seed, int64_t a, int64_t first_row, int64_t len, void* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t row = (uint64_t)(first_row + i);
    if (kind == 4) {
      uint64_t r = splitmix64(seed + (row >> 1));
      float u1 = ((float)((r >> 40) + 1)) * (1.0f / 16777216.0f);
      float u2 = ((float)((r >> 8) & 0xFFFFFF)) * (1.0f / 16777216.0f);
      float rad = sqrtf(-2.0f * logf(u1));
      float ang = 6.28318530717958647692f * u2;
      ((float*)out)[i] = (row & 1) ? rad * sinf(ang) : rad * cosf(ang);
      continue;
    }
    uint64_t r = splitmix64(seed + row);
    switch (kind) {
      case 0: ((int64_t*)out)[i] = (int64_t)__umul64hi(r, (uint64_t)a); break;
      case 1: ((int64_t*)out)[i] = (int64_t)(int32_t)(uint32_t)(r >> 32); break;
      case 2: ((double*)out)[i] = (double)(r >> (64 - a)); break;
      case 3: ((double*)out)[i] = (double)(r >> 11) * (1.0 / 9007199254740992.0); break;
      case 5: {
        uint64_t m = a >= 64 ? ~0ULL : ((1ULL << a) - 1);
        uint64_t x = row & m;
        x = (x * 0x9E3779B97F4A7C15ULL + seed) & m;
        x ^= x >> (a / 2 + 1);
        x = (x * 0xBF58476D1CE4E5B9ULL) & m;
        x ^= x >> (a / 2 + 1);
        ((int64_t*)out)[i] = (int64_t)x;
        break;
      }
      default: break;
    }
  }
}

}  // namespace dbx

extern "C" int32_t dbx_synth_fill(int32_t device, int32_t kind, uint64_t seed, int64_t a, int64_t first_row,
                                  int64_t len, void* dev_out) {
  using namespace dbx;
  if (kind < 0 || kind > 5 || len < 0 || !dev_out) { g_create_error.set("dbx_synth_fill: bad argument"); return DBX_ERR_INVALID; }
  DBX_CUDA_TRY(g_create_error, cudaSetDevice(device));
  if (len == 0) return DBX_OK;
  int grid = (int)((len + 255) / 256 < (int64_t)kNumSMs * 16 ? (len + 255) / 256 : (int64_t)kNumSMs * 16);
  synth_fill_kernel<<<grid, 256>>>(kind, seed, a, first_row, len, dev_out);
  count_launch();
  DBX_CUDA_TRY(g_create_error, cudaGetLastError());
  DBX_CUDA_TRY(g_create_error, cudaDeviceSynchronize());
  return DBX_OK;
}
