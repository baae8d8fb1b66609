// stubs.cu — entry points declared in include/dbx.h whose operators are not built yet.
// They fail loudly (DBX_ERR_UNSUPPORTED); nothing falls back to the CPU.
#include "runtime.h"
namespace dbx {
Op* make_filter_op(const dbx_predicate*, const int32_t*, int32_t, int, int32_t* st) { g_create_error.set("DBX_OP_FILTER is not built yet (the filter is fused into DBX_OP_AGG_PARTIAL)"); *st = DBX_ERR_UNSUPPORTED; return nullptr; }
}
