// stubs.cu — entry points declared in include/dbx.h whose operators are not built yet.
// They fail loudly (DBX_ERR_UNSUPPORTED); nothing falls back to the CPU.
#include "runtime.h"
namespace dbx {
Op* make_filter_op(const dbx_predicate*, const int32_t*, int32_t, int, int32_t* st) { g_create_error.set("DBX_OP_FILTER is not built yet"); *st = DBX_ERR_UNSUPPORTED; return nullptr; }
}
using namespace dbx;
extern "C" {
int32_t dbx_eval_distance(int32_t, int32_t, const dbx_column*, const dbx_column*, dbx_column*) { g_create_error.set("dbx_eval_distance is not built yet"); return DBX_ERR_UNSUPPORTED; }
int32_t dbx_knn_create(int32_t, int32_t, const dbx_column*, dbx_knn**) { g_create_error.set("kNN is not built yet"); return DBX_ERR_UNSUPPORTED; }
int32_t dbx_knn_search(dbx_knn*, const dbx_column*, int32_t, int32_t, int64_t*, float*) { return DBX_ERR_UNSUPPORTED; }
int32_t dbx_knn_destroy(dbx_knn*) { return DBX_OK; }
const char* dbx_knn_last_error(const dbx_knn*) { return "kNN is not built yet"; }
}
