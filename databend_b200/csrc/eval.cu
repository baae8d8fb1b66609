// eval.cu — dbx_eval_scalar: Evaluator::run over a DataBlock for numeric expressions.
//
// Reference replaced (paths relative to /root/reference):
//   Evaluator::{run, partial_run, eval_common_call, run_cast}   src/query/expression/src/evaluator.rs:247-465
//   ScalarFunction::eval + passthrough_nullable                  src/query/expression/src/function.rs:103, register.rs
//   plus / minus / multiply / divide / div / modulo              src/query/functions/src/scalars/numeric_basic_arithmetic/src/numeric_basic_arithmetic.rs:255-520
//   modulo semantics                                             .../arithmetic_modulo.rs:29-97
//   result types (ResultTypeOfBinary)                            src/query/codegen/src/writes/arithmetics_type.rs:240-265
//   to_<number> casts                                            src/query/functions/src/scalars/arithmetic/src/arithmetic.rs:490-600
//   comparison / boolean functions                               src/query/functions/src/scalars/comparison.rs, boolean.rs
//
// The reference walks the Expr tree and materialises one column per node (one full memory pass
// each).  Here the expression arrives as a postfix program; the host infers every node's type with
// the reference's rules, and ONE kernel evaluates the whole program per row in registers: every
// input column is read once and one output column is written.  Per-row errors (division by zero,
// number overflowed) are collected as "first failing row" like EvalContext::set_error; NULL rows
// never raise (passthrough_nullable evaluates under the validity).
#include <algorithm>
#include <cmath>
#include <vector>

#include "runtime.h"

namespace dbx {
namespace {

constexpr int kMaxExprNodes = DBX_MAX_EXPR_NODES;
constexpr int kEvalStack = 8;

struct NodeDev {
  int32_t kind;      // dbx_expr_kind
  int32_t func;      // dbx_func
  int32_t col;       // COLUMN: input slot
  int32_t out;       // result dtype of this node (dbx_dtype)
  int32_t a_type;    // CALL / CAST: dtype of the (first) argument
  int32_t b_type;    // CALL with two arguments: dtype of the second
  int32_t m_type;    // MODULO: LeastSuper(L, R) in which the remainder is computed
  int32_t try_cast;  // CAST: 1 = try_cast (failure -> NULL)
  uint64_t c_bits;   // CONST: value image in the node's type
  int32_t c_null;
  int32_t pad;
};
struct EvalParams {
  NodeDev nodes[kMaxExprNodes];
  DevCol cols[16];
  int32_t n_nodes, n_cols;
  int64_t n_rows;
  void* out_data;          // out dtype values (BOOL: one byte per row, packed afterwards)
  uint8_t* out_valid;      // one byte per row or nullptr
  int32_t out_dtype, pad;
  unsigned long long* first_error;  // min over failing rows of (row << 8 | code); ~0 = none
};

__host__ __device__ inline bool is_float_t(int t) { return t == DBX_F32 || t == DBX_F64; }
__host__ __device__ inline bool is_signed_t(int t) { return t == DBX_I8 || t == DBX_I16 || t == DBX_I32 || t == DBX_I64; }
__host__ __device__ inline int bits_of_t(int t) {
  switch (t) {
    case DBX_I8: case DBX_U8: return 8;
    case DBX_I16: case DBX_U16: return 16;
    case DBX_I32: case DBX_U32: case DBX_F32: return 32;
    default: return 64;
  }
}
// value images: integers sign/zero-extended to 64 bits, F64 as its bits, F32 as the f64 bits of the
// (exactly widened) value; BOOL 0/1
__device__ __forceinline__ double as_f64(uint64_t v, int t) {
  if (is_float_t(t)) return __longlong_as_double((long long)v);
  return is_signed_t(t) ? (double)(int64_t)v : (double)v;
}
// narrow a 64-bit two's complement result to an integer type (Rust wrapping arithmetic in that type)
__device__ __forceinline__ uint64_t wrap_int(uint64_t v, int t) {
  switch (t) {
    case DBX_I8: return (uint64_t)(int64_t)(int8_t)v;
    case DBX_I16: return (uint64_t)(int64_t)(int16_t)v;
    case DBX_I32: return (uint64_t)(int64_t)(int32_t)v;
    case DBX_U8: return v & 0xFFu;
    case DBX_U16: return v & 0xFFFFu;
    case DBX_U32: return v & 0xFFFFFFFFu;
    default: return v;
  }
}
__device__ __forceinline__ double int_min_f(int t) { return is_signed_t(t) ? -ldexp(1.0, bits_of_t(t) - 1) : 0.0; }
__device__ __forceinline__ double int_max_p1_f(int t) { return ldexp(1.0, is_signed_t(t) ? bits_of_t(t) - 1 : bits_of_t(t)); }  // max + 1, exact
// Rust `f64 as <int>`: truncates toward zero, saturates, NaN -> 0.  Returns the 64-bit image
// (sign-extended for signed types).
__device__ __forceinline__ uint64_t f64_as_int(double d, int t) {
  if (d != d) return 0;
  const int w = bits_of_t(t);
  if (is_signed_t(t)) {
    if (d <= int_min_f(t)) return w == 64 ? 0x8000000000000000ULL : (uint64_t)(-(int64_t)(1ULL << (w - 1)));
    if (d >= int_max_p1_f(t)) return w == 64 ? 0x7FFFFFFFFFFFFFFFULL : ((1ULL << (w - 1)) - 1);
    return (uint64_t)(int64_t)trunc(d);
  }
  if (d <= 0.0) return 0;
  if (d >= int_max_p1_f(t)) return w == 64 ? ~0ULL : ((1ULL << w) - 1);
  return (uint64_t)trunc(d);
}
// Rust `x as T` between any two numeric types (lossy where Rust is)
__device__ __forceinline__ uint64_t cast_as(uint64_t v, int from, int to) {
  if (is_float_t(to)) {
    double d = as_f64(v, from);
    if (to == DBX_F32) d = (double)(float)d;
    return (uint64_t)__double_as_longlong(d);
  }
  if (is_float_t(from)) {
    return f64_as_int(__longlong_as_double((long long)v), to);
  }
  return wrap_int(v, to);  // integer to integer: two's complement truncation / reinterpretation
}
// num_traits::cast::cast (checked): false when the value is not representable in `to`
__device__ __forceinline__ bool checked_cast(uint64_t v, int from, int to, uint64_t* out) {
  if (is_float_t(to)) { *out = cast_as(v, from, to); return true; }
  if (is_float_t(from)) {
    const double d = __longlong_as_double((long long)v);
    if (d != d) return false;
    const double tr = trunc(d);
    if (!(tr >= int_min_f(to) && tr < int_max_p1_f(to))) return false;
    *out = is_signed_t(to) ? (uint64_t)(int64_t)tr : (uint64_t)tr;
    return true;
  }
  // integer -> integer: value must lie in the destination range
  if (is_signed_t(from)) {
    const int64_t x = (int64_t)v;
    if (is_signed_t(to)) {
      if (bits_of_t(to) < 64) { const int64_t lim = 1LL << (bits_of_t(to) - 1); if (x < -lim || x >= lim) return false; }
    } else {
      if (x < 0) return false;
      if (bits_of_t(to) < 64 && (uint64_t)x >= (1ULL << bits_of_t(to))) return false;
    }
  } else {
    if (is_signed_t(to)) { if (v >= (1ULL << (bits_of_t(to) - 1))) return false; }
    else if (bits_of_t(to) < 64 && v >= (1ULL << bits_of_t(to))) return false;
  }
  *out = v;
  return true;
}

__device__ __forceinline__ uint64_t load_image(const DevCol& c, int64_t r) {
  if (c.is_const) return c.const_bits;
  const char* base = (const char*)c.data;
  switch (c.dtype) {
    case DBX_I64: case DBX_U64: case DBX_F64: return ((const uint64_t*)base)[r];
    case DBX_I32: return (uint64_t)(int64_t)((const int32_t*)base)[r];
    case DBX_U32: return ((const uint32_t*)base)[r];
    case DBX_F32: return (uint64_t)__double_as_longlong((double)((const float*)base)[r]);
    case DBX_I16: return (uint64_t)(int64_t)((const int16_t*)base)[r];
    case DBX_U16: return ((const uint16_t*)base)[r];
    case DBX_I8: return (uint64_t)(int64_t)((const int8_t*)base)[r];
    case DBX_U8: return ((const uint8_t*)base)[r];
    case DBX_BOOL: return (uint64_t)bit_test((const uint8_t*)base, c.dbit_off + r);
    default: return 0;
  }
}
// three-way compare of two values of the same dtype (OrderedFloat for floats: NaN greatest and equal to itself)
__device__ __forceinline__ int cmp_same(uint64_t a, uint64_t b, int t) {
  if (is_float_t(t)) {
    const double x = __longlong_as_double((long long)a), y = __longlong_as_double((long long)b);
    const bool xn = x != x, yn = y != y;
    if (xn | yn) return xn == yn ? 0 : (xn ? 1 : -1);
    return x < y ? -1 : (x > y ? 1 : 0);
  }
  if (is_signed_t(t)) return (int64_t)a < (int64_t)b ? -1 : ((int64_t)a > (int64_t)b ? 1 : 0);
  return a < b ? -1 : (a > b ? 1 : 0);
}
enum : int { ERR_DIV_ZERO = 1, ERR_DIVIDED_BY_ZERO = 2, ERR_OVERFLOW = 3 };

__global__ void __launch_bounds__(256) eval_kernel(const __grid_constant__ EvalParams p) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < p.n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    uint64_t sv[kEvalStack];
    bool sn[kEvalStack];  // valid
    int sp = 0;
    int err = 0;
    for (int i = 0; i < p.n_nodes; ++i) {
      const NodeDev& nd = p.nodes[i];
      if (nd.kind == DBX_EXPR_COLUMN) {
        const DevCol& c = p.cols[nd.col];
        const bool ok = c.is_const ? c.is_const != 2 : (!c.validity || bit_test(c.validity, c.vbit_off + r));
        sv[sp] = ok ? load_image(c, r) : 0;
        sn[sp] = ok;
        ++sp;
      } else if (nd.kind == DBX_EXPR_CONST) {
        sv[sp] = nd.c_bits; sn[sp] = !nd.c_null; ++sp;
      } else if (nd.kind == DBX_EXPR_CAST) {
        uint64_t out = 0;
        bool ok = sn[sp - 1];
        if (ok) {
          const int from = nd.a_type, to = nd.out;
          bool fits = true;
          if (to == DBX_BOOL) out = is_float_t(from) ? (as_f64(sv[sp - 1], from) != 0.0) : (sv[sp - 1] != 0);
          else if (from == DBX_BOOL) out = cast_as(sv[sp - 1], DBX_U8, to);
          else if (is_float_t(from) && !is_float_t(to)) {  // round cast (numeric_cast_option = rounding, the default)
            const double d = round(__longlong_as_double((long long)sv[sp - 1]));
            fits = checked_cast((uint64_t)__double_as_longlong(d), DBX_F64, to, &out);
          } else {
            fits = checked_cast(sv[sp - 1], from, to, &out);
          }
          if (!fits) {
            out = 0;
            if (nd.try_cast) ok = false; else err = err ? err : ERR_OVERFLOW;
          }
        }
        sv[sp - 1] = out; sn[sp - 1] = ok;
      } else {  // CALL
        const int f = nd.func;
        if (f == DBX_FN_NOT || f == DBX_FN_NEGATE || f == DBX_FN_IS_NULL || f == DBX_FN_IS_NOT_NULL) {
          const uint64_t a = sv[sp - 1];
          const bool an = sn[sp - 1];
          if (f == DBX_FN_IS_NULL) { sv[sp - 1] = an ? 0 : 1; sn[sp - 1] = true; }
          else if (f == DBX_FN_IS_NOT_NULL) { sv[sp - 1] = an ? 1 : 0; sn[sp - 1] = true; }
          else if (f == DBX_FN_NOT) { sv[sp - 1] = a ? 0 : 1; }
          else {  // NEGATE: -(a as Negate type); floats keep their type
            // 64-bit inputs are checked (arithmetic.rs:226-276): -(i64::MIN) and -(u64 > 2^63) raise
            if (is_float_t(nd.out)) sv[sp - 1] = (uint64_t)__double_as_longlong(-as_f64(a, nd.a_type));
            else {
              if (an && ((nd.a_type == DBX_I64 && a == 0x8000000000000000ULL) || (nd.a_type == DBX_U64 && a > 0x8000000000000000ULL))) err = err ? err : ERR_OVERFLOW;
              sv[sp - 1] = wrap_int((uint64_t)0 - cast_as(a, nd.a_type, nd.out), nd.out);
            }
          }
          continue;
        }
        const uint64_t b = sv[sp - 1], a = sv[sp - 2];
        const bool bn = sn[sp - 1], an = sn[sp - 2];
        sp -= 1;
        uint64_t out = 0;
        bool ok = an && bn;
        if (f == DBX_FN_AND || f == DBX_FN_OR) {  // three-valued logic (boolean.rs: and / or on nullable booleans)
          const bool at = an && a, af = an && !a, bt = bn && b, bf = bn && !b;
          if (f == DBX_FN_AND) { if (af || bf) { out = 0; ok = true; } else if (at && bt) { out = 1; ok = true; } else ok = false; }
          else { if (at || bt) { out = 1; ok = true; } else if (af && bf) { out = 0; ok = true; } else ok = false; }
        } else if (ok) {
          const int ta = nd.a_type, tb = nd.b_type, to = nd.out;
          if (f == DBX_FN_PLUS || f == DBX_FN_MINUS || f == DBX_FN_MULTIPLY) {
            if (is_float_t(to)) {
              const double x = as_f64(a, ta), y = as_f64(b, tb);
              out = (uint64_t)__double_as_longlong(f == DBX_FN_PLUS ? x + y : (f == DBX_FN_MINUS ? x - y : x * y));
            } else {  // (a as T) op (b as T), wrapping in T
              const uint64_t x = cast_as(a, ta, to), y = cast_as(b, tb, to);
              out = wrap_int(f == DBX_FN_PLUS ? x + y : (f == DBX_FN_MINUS ? x - y : x * y), to);
            }
          } else if (f == DBX_FN_DIVIDE) {
            const double y = as_f64(b, tb);
            if (y == 0.0) err = err ? err : ERR_DIVIDED_BY_ZERO;
            else out = (uint64_t)__double_as_longlong(as_f64(a, ta) / y);
          } else if (f == DBX_FN_DIV) {
            const double y = as_f64(b, tb);
            if (y == 0.0) err = err ? err : ERR_DIVIDED_BY_ZERO;
            else out = f64_as_int(as_f64(a, ta) / y, to);
          } else if (f == DBX_FN_MODULO) {
            const bool b_zero = is_float_t(tb) ? (__longlong_as_double((long long)b) == 0.0) : (b == 0);
            if (b_zero) err = err ? err : ERR_DIV_ZERO;
            else {
              const int tm = nd.m_type;
              const uint64_t x = cast_as(a, ta, tm), y = cast_as(b, tb, tm);
              uint64_t rem;
              if (is_float_t(tm)) {
                double fr = fmod(__longlong_as_double((long long)x), __longlong_as_double((long long)y));
                if (tm == DBX_F32) fr = (double)fmodf((float)__longlong_as_double((long long)x), (float)__longlong_as_double((long long)y));
                rem = (uint64_t)__double_as_longlong(fr);
              } else if (is_signed_t(tm)) {
                const int64_t xs = (int64_t)x, ys = (int64_t)y;
                rem = (ys == -1 || ys == 0) ? 0 : (uint64_t)(xs % ys);  // MIN % -1 = 0; a divisor that WRAPS to 0 in M cannot occur for b != 0 except by truncation
                if (ys == 0) err = err ? err : ERR_DIV_ZERO;
              } else {
                rem = y == 0 ? 0 : x % y;
                if (y == 0) err = err ? err : ERR_DIV_ZERO;
              }
              out = cast_as(rem, tm, to);
            }
          } else {  // comparisons: both sides were cast to a common type by the type checker (a_type == b_type)
            const int c3 = cmp_same(a, b, ta);
            out = f == DBX_FN_EQ ? c3 == 0 : f == DBX_FN_NOTEQ ? c3 != 0 : f == DBX_FN_LT ? c3 < 0 : f == DBX_FN_LTE ? c3 <= 0 : f == DBX_FN_GT ? c3 > 0 : c3 >= 0;
          }
        }
        sv[sp - 1] = out; sn[sp - 1] = ok;
      }
    }
    const bool valid = sn[0];
    // an error is raised by the CALL whose own arguments are valid on this row (passthrough_nullable masks
    // only that call's NULL rows), whatever the validity of the final value
    if (err) atomicMin(p.first_error, ((unsigned long long)r << 8) | (unsigned long long)err);
    const uint64_t v = valid ? sv[0] : 0;
    switch (p.out_dtype) {
      case DBX_BOOL: ((uint8_t*)p.out_data)[r] = (uint8_t)(v != 0); break;
      case DBX_I8: case DBX_U8: ((uint8_t*)p.out_data)[r] = (uint8_t)v; break;
      case DBX_I16: case DBX_U16: ((uint16_t*)p.out_data)[r] = (uint16_t)v; break;
      case DBX_I32: case DBX_U32: ((uint32_t*)p.out_data)[r] = (uint32_t)v; break;
      case DBX_F32: ((float*)p.out_data)[r] = (float)__longlong_as_double((long long)v); break;
      default: ((uint64_t*)p.out_data)[r] = v; break;
    }
    if (p.out_valid) p.out_valid[r] = valid ? 1 : 0;
  }
}
__global__ void eval_pack_bits_kernel(const uint8_t* bytes, int64_t n, uint8_t* bits) {
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

// ---- the reference's type rules (arithmetics_type.rs codegen)
inline int make_type(int bits, bool is_signed, bool is_float) {
  if (is_float) return bits <= 32 ? DBX_F32 : DBX_F64;
  switch (bits) {
    case 8: return is_signed ? DBX_I8 : DBX_U8;
    case 16: return is_signed ? DBX_I16 : DBX_U16;
    case 32: return is_signed ? DBX_I32 : DBX_U32;
    default: return is_signed ? DBX_I64 : DBX_U64;
  }
}
inline int next_bits(int b) { return b >= 64 ? 64 : b * 2; }
inline int type_add_mul(int a, int b) { return make_type(next_bits(std::max(bits_of_t(a), bits_of_t(b))), is_signed_t(a) || is_signed_t(b), is_float_t(a) || is_float_t(b)); }
inline int type_minus(int a, int b) { return make_type(next_bits(std::max(bits_of_t(a), bits_of_t(b))), true, is_float_t(a) || is_float_t(b)); }
inline int type_intdiv(int a, int b) {  // NumberDataType::is_signed counts the float types as signed (number.rs:392-404)
  return make_type(std::max(bits_of_t(a), bits_of_t(b)), is_signed_t(a) || is_signed_t(b) || is_float_t(a) || is_float_t(b), false);
}
inline int type_super(int a, int b) { return make_type(std::max(bits_of_t(a), bits_of_t(b)), is_signed_t(a) || is_signed_t(b), is_float_t(a) || is_float_t(b)); }
inline int type_modulo(int a, int b) {
  if (is_float_t(a) || is_float_t(b)) return DBX_F64;
  const bool s = is_signed_t(a);
  return make_type(s ? next_bits(bits_of_t(b)) : bits_of_t(b), s, false);
}
inline int type_negate(int a) { return is_float_t(a) ? a : make_type(next_bits(bits_of_t(a)), true, false); }

}  // namespace
}  // namespace dbx

using namespace dbx;

extern "C" int32_t dbx_eval_scalar(int32_t device, const dbx_expr* expr, const dbx_block* block, int32_t out_mem, dbx_block* out,
                                   int32_t* out_dtype, int64_t* first_error_row) {
  ErrorSink& err = g_create_error;
  if (!expr || !block || !out || expr->n_nodes < 1 || expr->n_nodes > kMaxExprNodes || block->num_cols > 16) { err.set("dbx_eval_scalar: bad argument"); return DBX_ERR_INVALID; }
  if (first_error_row) *first_error_row = -1;
  int32_t ndev = 0;
  DBX_TRY(dbx_device_count(&ndev));
  if (device < 0 || device >= ndev) { err.set("dbx_eval_scalar: device index out of range"); return DBX_ERR_INVALID; }
  DBX_CUDA_TRY(err, cudaSetDevice(device));
  const int64_t n = block->num_rows;
  // ---- type inference over the postfix program
  EvalParams p;
  memset(&p, 0, sizeof(p));
  int tstack[kEvalStack];
  bool nstack[kEvalStack];  // nullable
  int sp = 0;
  auto numeric = [](int t) { return t != DBX_BOOL && t != DBX_VEC_F32 && dtype_size(t) > 0; };
  for (int i = 0; i < expr->n_nodes; ++i) {
    const dbx_expr_node& in = expr->nodes[i];
    NodeDev& nd = p.nodes[i];
    nd.kind = in.kind; nd.func = in.func;
    if (in.kind == DBX_EXPR_COLUMN) {
      if (in.col < 0 || in.col >= block->num_cols) { err.set("eval: column index outside the block"); return DBX_ERR_INVALID; }
      const dbx_column& c = block->cols[in.col];
      if (c.dtype == DBX_VEC_F32 || (c.dtype != DBX_BOOL && dtype_size(c.dtype) == 0)) { err.set("eval: only numeric and boolean columns"); return DBX_ERR_UNSUPPORTED; }
      if (sp >= kEvalStack) { err.set("eval: expression too deep"); return DBX_ERR_UNSUPPORTED; }
      nd.col = in.col; nd.out = c.dtype;
      tstack[sp] = c.dtype; nstack[sp] = c.validity != nullptr || (c.is_const && c.konst.is_null); ++sp;
    } else if (in.kind == DBX_EXPR_CONST) {
      if (sp >= kEvalStack) { err.set("eval: expression too deep"); return DBX_ERR_UNSUPPORTED; }
      const int t = in.c.dtype;
      nd.out = t; nd.c_null = in.c.is_null;
      if (t == DBX_F32) { const double d = (double)(float)in.c.v.f64; memcpy(&nd.c_bits, &d, 8); }
      else nd.c_bits = in.c.v.u64;
      tstack[sp] = t; nstack[sp] = in.c.is_null != 0; ++sp;
    } else if (in.kind == DBX_EXPR_CAST) {
      if (sp < 1) { err.set("eval: malformed postfix program"); return DBX_ERR_INVALID; }
      const int to = in.cast_to;
      if (to != DBX_BOOL && !numeric(to)) { err.set("eval: cast target must be numeric or boolean"); return DBX_ERR_UNSUPPORTED; }
      nd.a_type = tstack[sp - 1]; nd.out = to; nd.try_cast = in.try_cast;
      tstack[sp - 1] = to; nstack[sp - 1] = nstack[sp - 1] || in.try_cast;
    } else if (in.kind == DBX_EXPR_CALL) {
      const int f = in.func;
      const bool unary = f == DBX_FN_NOT || f == DBX_FN_NEGATE || f == DBX_FN_IS_NULL || f == DBX_FN_IS_NOT_NULL;
      if (sp < (unary ? 1 : 2)) { err.set("eval: malformed postfix program"); return DBX_ERR_INVALID; }
      if (unary) {
        const int ta = tstack[sp - 1];
        nd.a_type = ta;
        if (f == DBX_FN_NOT) { if (ta != DBX_BOOL) { err.set("eval: not() needs a Boolean argument"); return DBX_ERR_INVALID; } nd.out = DBX_BOOL; }
        else if (f == DBX_FN_NEGATE) { if (!numeric(ta)) { err.set("eval: minus() needs a numeric argument"); return DBX_ERR_INVALID; } nd.out = type_negate(ta); }
        else { nd.out = DBX_BOOL; nstack[sp - 1] = false; }
        tstack[sp - 1] = nd.out;
        continue;
      }
      const int ta = tstack[sp - 2], tb = tstack[sp - 1];
      nd.a_type = ta; nd.b_type = tb;
      const bool nullable = nstack[sp - 2] || nstack[sp - 1];
      int to;
      switch (f) {
        case DBX_FN_PLUS: case DBX_FN_MULTIPLY: case DBX_FN_MINUS: case DBX_FN_DIVIDE: case DBX_FN_DIV: case DBX_FN_MODULO:
          if (!numeric(ta) || !numeric(tb)) { err.set("eval: arithmetic needs numeric arguments"); return DBX_ERR_INVALID; }
          to = f == DBX_FN_MINUS ? type_minus(ta, tb) : f == DBX_FN_DIVIDE ? DBX_F64 : f == DBX_FN_DIV ? type_intdiv(ta, tb)
               : f == DBX_FN_MODULO ? type_modulo(ta, tb) : type_add_mul(ta, tb);
          nd.m_type = type_super(ta, tb);
          break;
        case DBX_FN_EQ: case DBX_FN_NOTEQ: case DBX_FN_LT: case DBX_FN_LTE: case DBX_FN_GT: case DBX_FN_GTE:
          if (ta != tb) { err.set("eval: comparison arguments must have one type (the type checker casts both sides to their common super type: add DBX_EXPR_CAST nodes)"); return DBX_ERR_INVALID; }
          to = DBX_BOOL;
          break;
        case DBX_FN_AND: case DBX_FN_OR:
          if (ta != DBX_BOOL || tb != DBX_BOOL) { err.set("eval: and / or need Boolean arguments"); return DBX_ERR_INVALID; }
          to = DBX_BOOL;
          break;
        default: err.set("eval: unknown function"); return DBX_ERR_INVALID;
      }
      nd.out = to;
      sp -= 1;
      tstack[sp - 1] = to; nstack[sp - 1] = nullable;
    } else { err.set("eval: unknown node kind"); return DBX_ERR_INVALID; }
  }
  if (sp != 1) { err.set("eval: postfix program does not reduce to one value"); return DBX_ERR_INVALID; }
  const int ot = tstack[0];
  const bool o_nullable = nstack[0];
  if (out_dtype) *out_dtype = ot | (o_nullable ? DBX_NULLABLE : 0);

  cudaStream_t st = nullptr;
  DBX_CUDA_TRY(err, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
  // ---- inputs on the device
  std::vector<DevBuf> owned;
  p.n_cols = block->num_cols;
  for (int c = 0; c < block->num_cols; ++c) {
    const dbx_column& col = block->cols[c];
    DevCol& dc = p.cols[c];
    memset(&dc, 0, sizeof(dc));
    dc.dtype = col.dtype;
    if (col.dtype == DBX_VEC_F32 || (col.dtype != DBX_BOOL && dtype_size(col.dtype) == 0)) continue;  // never referenced (checked above)
    if (col.len != n) { err.set("eval: column length differs from num_rows"); return DBX_ERR_INVALID; }
    if (col.is_const) {
      dc.is_const = col.konst.is_null ? 2 : 1;
      if (col.dtype == DBX_F32 || col.dtype == DBX_F64) { const double d = col.dtype == DBX_F32 ? (double)(float)col.konst.v.f64 : col.konst.v.f64; memcpy(&dc.const_bits, &d, 8); }
      else dc.const_bits = col.konst.v.u64;
      continue;
    }
    if (col.mem == DBX_MEM_DEVICE) { dc.data = col.data; dc.validity = col.validity; dc.vbit_off = col.validity_bit_offset; dc.dbit_off = col.data_bit_offset; continue; }
    const bool is_bool = col.dtype == DBX_BOOL;
    const int64_t b0 = is_bool ? col.data_bit_offset >> 3 : 0;
    const size_t bytes = is_bool ? (size_t)(((col.data_bit_offset + n + 7) >> 3) - b0) : (size_t)n * dtype_size(col.dtype);
    owned.emplace_back();
    DBX_CUDA_TRY(err, owned.back().ensure(bytes ? bytes : 1));
    if (bytes) DBX_CUDA_TRY(err, cudaMemcpyAsync(owned.back().p, (const char*)col.data + b0, bytes, cudaMemcpyHostToDevice, st));
    dc.data = owned.back().p;
    dc.dbit_off = is_bool ? (col.data_bit_offset & 7) : 0;
    if (col.validity) {
      const int64_t v0 = col.validity_bit_offset >> 3, v1 = (col.validity_bit_offset + n + 7) >> 3;
      owned.emplace_back();
      DBX_CUDA_TRY(err, owned.back().ensure((size_t)std::max<int64_t>(v1 - v0, 1)));
      if (v1 > v0) DBX_CUDA_TRY(err, cudaMemcpyAsync(owned.back().p, col.validity + v0, (size_t)(v1 - v0), cudaMemcpyHostToDevice, st));
      dc.validity = (const uint8_t*)owned.back().p;
      dc.vbit_off = col.validity_bit_offset & 7;
    }
  }
  // ---- output
  auto ob = std::make_unique<OwnedBlock>();
  ob->device = device;
  const size_t esz = ot == DBX_BOOL ? 1 : dtype_size(ot);
  void *od = nullptr, *ov = nullptr, *ferr = nullptr;
  DBX_CUDA_TRY(err, pool_alloc(device, st, (size_t)std::max<int64_t>(n, 1) * esz, &od));
  ob->dev_allocs.push_back(od);
  if (o_nullable) { DBX_CUDA_TRY(err, pool_alloc(device, st, (size_t)std::max<int64_t>(n, 1), &ov)); ob->dev_allocs.push_back(ov); }
  DBX_CUDA_TRY(err, pool_alloc(device, st, 8, &ferr));
  ob->dev_allocs.push_back(ferr);
  DBX_CUDA_TRY(err, cudaMemsetAsync(ferr, 0xFF, 8, st));
  p.n_nodes = expr->n_nodes; p.n_rows = n; p.out_data = od; p.out_valid = (uint8_t*)ov; p.out_dtype = ot;
  p.first_error = (unsigned long long*)ferr;
  if (n) {
    eval_kernel<<<(int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8)), 256, 0, st>>>(p);
    count_launch();
    DBX_CUDA_TRY(err, cudaGetLastError());
  }
  dbx_column oc;
  memset(&oc, 0, sizeof(oc));
  oc.dtype = ot; oc.mem = DBX_MEM_DEVICE; oc.len = n; oc.data = od;
  auto pack = [&](const void* bytes, const void** dst) -> int32_t {
    void* bits = nullptr;
    DBX_CUDA_TRY(err, pool_alloc(device, st, (size_t)(n + 7) / 8 + 8, &bits));
    ob->dev_allocs.push_back(bits);
    if (n) { eval_pack_bits_kernel<<<(int)std::max<int64_t>(1, std::min<int64_t>(((n + 7) / 8 + 255) / 256, (int64_t)kNumSMs * 8)), 256, 0, st>>>((const uint8_t*)bytes, n, (uint8_t*)bits); count_launch(); }
    *dst = bits;
    return DBX_OK;
  };
  if (ot == DBX_BOOL) DBX_TRY(pack(od, &oc.data));
  if (o_nullable) { const void* vb = nullptr; DBX_TRY(pack(ov, &vb)); oc.validity = (const uint8_t*)vb; oc.null_count = -1; }
  unsigned long long herr = ~0ULL;
  DBX_CUDA_TRY(err, cudaMemcpyAsync(&herr, ferr, 8, cudaMemcpyDeviceToHost, st));
  DBX_CUDA_TRY(err, cudaStreamSynchronize(st));
  if (herr != ~0ULL) {  // EvalContext::render_error: "<message>, during run expr" with the first failing row
    const int code = (int)(herr & 0xFF);
    const int64_t row = (int64_t)(herr >> 8);
    if (first_error_row) *first_error_row = row;
    const char* msg = code == ERR_DIV_ZERO ? "Division by zero" : code == ERR_DIVIDED_BY_ZERO ? "divided by zero" : "number overflowed";
    err.set(std::string(msg) + " while evaluating the expression (first failing row " + std::to_string(row) + ")");
    return DBX_ERR_BAD_ARGUMENTS;
  }
  ob->cols.push_back(oc);
  int32_t rc = pull_owned_block(ob, device, st, err, out_mem, out);
  if (rc == DBX_OK) out->num_rows = n;
  return rc;
}
