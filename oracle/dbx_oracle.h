/* dbx_oracle.h — CPU ORACLE interface (test infrastructure only; see dbx_oracle.c). */
#ifndef DBX_ORACLE_H_
#define DBX_ORACLE_H_
#include <stdint.h>
#include "../include/dbx.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Result of a group-by: one entry per group, arbitrary group order.
 * key_bits: group-key value `as u64` (floats: canonical bits); agg_bits: the result value's
 * bit pattern in agg_dtype (DBX_I64 / DBX_U64 / DBX_F64 / argument type for min,max). */
typedef struct orc_agg_result {
  int64_t n_groups;
  int32_t n_group_cols, n_aggs;
  uint64_t* key_bits[DBX_MAX_GROUP_COLS];
  uint8_t* key_valid[DBX_MAX_GROUP_COLS];
  uint64_t* agg_bits[DBX_MAX_AGGS];
  uint8_t* agg_valid[DBX_MAX_AGGS];
  int32_t agg_dtype[DBX_MAX_AGGS];
} orc_agg_result;

int orc_filter_select(const dbx_block* blk, const dbx_predicate* pred, uint32_t* sel, int64_t* n_sel, int64_t* err_row);
int orc_take_column(const dbx_column* c, const uint32_t* sel, int64_t n_sel, void* out_data, uint8_t* out_valid);
uint64_t orc_agg_hash_u64(uint64_t x);
int orc_filter_group_agg(const dbx_block* blk, const dbx_agg_params* p, int threads, orc_agg_result* out, int64_t* err_row);
void orc_agg_result_free(orc_agg_result* r);
int orc_hash_join_inner(const dbx_column* build_key, const dbx_column* probe_key, int64_t** out_probe_idx,
                        int64_t** out_build_idx, int64_t* n_out);
int orc_hash_join(int kind, const dbx_column* build_key, const dbx_column* probe_key, int64_t** out_probe_idx,
                  int64_t** out_build_idx, int64_t* n_out);
void orc_free(void* p);
int orc_topk(const dbx_column* key, int asc, int nulls_first, int64_t k, int64_t* out_idx, int64_t* n_out);
float orc_cosine_distance(const float* a, const float* b, int64_t n);
float orc_l2_distance(const float* a, const float* b, int64_t n);
void orc_distance_rows(int kind, const float* lhs, int lhs_const, const float* rhs, int rhs_const, int64_t rows,
                       int64_t dim, float* out, int threads);
int orc_synth_fill(int kind, uint64_t seed, int64_t a, int64_t first_row, int64_t len, void* out, int threads);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
