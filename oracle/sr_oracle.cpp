// sr_oracle.cpp -- CPU restatement of the StarRocks BE hot path.  TEST INFRASTRUCTURE ONLY
// (see sr_oracle.h).  Each section cites the reference file:line whose algorithm it restates.
// Nothing here is copied from the reference; it is written from the behaviour of those files.

#include "sr_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_set>
#include <vector>

typedef __int128 i128;
typedef unsigned __int128 u128;

static thread_local std::string g_err;
static int32_t fail(int32_t code, const std::string& msg) {
    g_err = msg;
    return code;
}
extern "C" const char* orc_last_error(void) {
    return g_err.c_str();
}

// ---------------------------------------------------------------------------------------
// types
// ---------------------------------------------------------------------------------------
static inline int type_width(int32_t t) {
    switch (t) {
    case SR_TYPE_BOOLEAN:
    case SR_TYPE_TINYINT:
        return 1;
    case SR_TYPE_SMALLINT:
        return 2;
    case SR_TYPE_INT:
    case SR_TYPE_FLOAT:
    case SR_TYPE_DATE:
    case SR_TYPE_DECIMAL32:
        return 4;
    case SR_TYPE_BIGINT:
    case SR_TYPE_DOUBLE:
    case SR_TYPE_DATETIME:
    case SR_TYPE_DECIMAL64:
        return 8;
    case SR_TYPE_LARGEINT:
    case SR_TYPE_DECIMAL128:
        return 16;
    default:
        return 0;
    }
}
static inline bool is_float_class(int32_t t) {
    return t == SR_TYPE_FLOAT || t == SR_TYPE_DOUBLE;
}
static inline bool is_decimal(int32_t t) {
    return t == SR_TYPE_DECIMAL32 || t == SR_TYPE_DECIMAL64 || t == SR_TYPE_DECIMAL128;
}

static inline int64_t load_int(const void* data, int32_t type, int64_t i) {
    switch (type) {
    case SR_TYPE_BOOLEAN:
        return ((const uint8_t*)data)[i];
    case SR_TYPE_TINYINT:
        return ((const int8_t*)data)[i];
    case SR_TYPE_SMALLINT:
        return ((const int16_t*)data)[i];
    case SR_TYPE_INT:
    case SR_TYPE_DATE:
    case SR_TYPE_DECIMAL32:
        return ((const int32_t*)data)[i];
    case SR_TYPE_BIGINT:
    case SR_TYPE_DATETIME:
    case SR_TYPE_DECIMAL64:
        return ((const int64_t*)data)[i];
    case SR_TYPE_LARGEINT:
    case SR_TYPE_DECIMAL128:
        return (int64_t)((const i128*)data)[i];
    default:
        return 0;
    }
}
static inline double load_double(const void* data, int32_t type, int64_t i) {
    if (type == SR_TYPE_FLOAT) return ((const float*)data)[i];
    if (type == SR_TYPE_DOUBLE) return ((const double*)data)[i];
    return (double)load_int(data, type, i);
}

static const sr_col_view* find_col(const sr_chunk_view* c, int32_t slot) {
    for (int k = 0; k < c->num_cols; k++)
        if (c->cols[k].slot_id == slot) return &c->cols[k];
    return nullptr;
}

// ---------------------------------------------------------------------------------------
// hash functions
// ---------------------------------------------------------------------------------------
// JoinKeyHash<T,4>: be/src/exec/join/join_hash_map_helper.h:35-43
extern "C" uint32_t orc_join_key_hash32(uint32_t v, uint32_t num_log_buckets) {
    const uint32_t a = 2654435761u;
    v ^= v >> (32 - num_log_buckets);
    const uint32_t fraction = v * a;
    return fraction >> (32 - num_log_buckets);
}
// JoinKeyHash<T,8>: be/src/exec/join/join_hash_map_helper.h:45-54
extern "C" uint32_t orc_join_key_hash64(uint64_t v, uint32_t num_log_buckets) {
    const uint64_t a = 11400714819323198485ull;
    v ^= v >> (64 - num_log_buckets);
    const uint64_t fraction = v * a;
    return (uint32_t)(fraction >> (64 - num_log_buckets));
}
// JoinHashMapHelper::calc_bucket_size: :70-77 with phmap NormalizeCapacity
// (be/src/base/phmap/phmap.h:485: n ? ~size_t{} >> clz(n) : 1)
extern "C" uint32_t orc_calc_bucket_size(uint32_t size) {
    const uint64_t MAX_BUCKET_SIZE = 1ull << 31;
    uint64_t expect = (uint64_t)size + (size - 1) / 4;
    if (expect >= MAX_BUCKET_SIZE) return (uint32_t)MAX_BUCKET_SIZE;
    uint64_t norm = expect ? (~0ull >> __builtin_clzll(expect)) : 1;
    return (uint32_t)(norm + 1);
}

// CRC32C (Castagnoli, reflected 0x82F63B78) -- what _mm_crc32_u32 / _mm_crc32_u8 compute.
static uint32_t g_crc32c_tab[256];
static uint32_t g_zcrc_tab[256];
static std::atomic<bool> g_tabs_ready{false};
static void init_tabs() {
    if (g_tabs_ready.load(std::memory_order_acquire)) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i, z = i;
        for (int k = 0; k < 8; k++) {
            c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            z = (z & 1) ? (z >> 1) ^ 0xEDB88320u : z >> 1;
        }
        g_crc32c_tab[i] = c;
        g_zcrc_tab[i] = z;
    }
    g_tabs_ready.store(true, std::memory_order_release);
}
// crc_hash_32: be/src/base/hash/hash.h:96-130.  The SSE4.2 instruction has no pre/post
// inversion, and processing a 32-bit word equals processing its 4 bytes in LE order.
extern "C" uint32_t orc_crc32c(const void* data, int32_t bytes, uint32_t hash) {
    init_tabs();
    const uint8_t* p = (const uint8_t*)data;
    while (bytes-- > 0) hash = g_crc32c_tab[(hash ^ *p++) & 0xFF] ^ (hash >> 8);
    return hash;
}
// zlib crc32(seed, data, len): pre/post inverted.
extern "C" uint32_t orc_zlib_crc32(const void* data, int32_t bytes, uint32_t seed) {
    init_tabs();
    const uint8_t* p = (const uint8_t*)data;
    uint32_t c = seed ^ 0xFFFFFFFFu;
    while (bytes-- > 0) c = g_zcrc_tab[(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
// crc_hash_32 (be/src/base/hash/hash.h:96-130) = CRC32C of the bytes, then phmap_mix<4>
// (:25-33: l = a * 0xcc9e2d51; l ^ (l >> 32)) because the low half of a CRC has poor uniformity.
extern "C" uint32_t orc_crc_hash_32(const void* data, int32_t bytes, uint32_t seed) {
    const uint64_t a = orc_crc32c(data, bytes, seed);
    const uint64_t l = a * 0xcc9e2d51ull;
    return (uint32_t)(l ^ (l >> 32));
}
extern "C" uint32_t orc_join_key_hash_slice(const void* data, int32_t size, uint32_t num_buckets) {
    return orc_crc_hash_32(data, size, 0x811C9DC5u) & (num_buckets - 1);
}
// HashUtil::fnv_hash: be/src/base/hash/hash_util.hpp:127-134 (hash = (byte ^ hash) * prime)
extern "C" uint32_t orc_fnv_hash(const void* data, int32_t bytes, uint32_t hash) {
    const uint8_t* p = (const uint8_t*)data;
    while (bytes-- > 0) {
        hash = (*p ^ hash) * 0x01000193u;
        ++p;
    }
    return hash;
}
extern "C" uint32_t orc_xorshift32(uint32_t x) {
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    return x;
}
extern "C" uint32_t orc_reduce_op(uint32_t l, uint32_t r) {
    return (uint32_t)(((uint64_t)l * (uint64_t)r) >> 32);
}

// ---------------------------------------------------------------------------------------
// filter: column_filter_range.cpp:39-148 (semantics: order-preserving compaction)
// ---------------------------------------------------------------------------------------
extern "C" int64_t orc_filter_range(const uint8_t* filter, void* data, int32_t elem_size, int64_t from, int64_t to) {
    uint8_t* d = (uint8_t*)data;
    int64_t out = from;
    for (int64_t i = from; i < to; i++) {
        if (filter[i]) {
            if (out != i) memcpy(d + out * elem_size, d + i * elem_size, elem_size);
            out++;
        }
    }
    return out;
}

// ---------------------------------------------------------------------------------------
// predicates: column_operator_predicate.h:41-111 (`sel[i] = op(v[i], const)`, nullable:
// `!null[i] && op`), merged by AND (chunk_predicate_evaluator.cpp:84-149).
// ---------------------------------------------------------------------------------------
template <typename V>
static inline bool pred_cmp(int32_t op, V v, V lo, V hi) {
    switch (op) {
    case SR_PRED_EQ:
        return v == lo;
    case SR_PRED_NE:
        return v != lo;
    case SR_PRED_LT:
        return v < lo;
    case SR_PRED_LE:
        return v <= lo;
    case SR_PRED_GT:
        return v > lo;
    case SR_PRED_GE:
        return v >= lo;
    case SR_PRED_BETWEEN:
        return v >= lo && v <= hi;
    default:
        return false;
    }
}

static int32_t eval_pred_range(const sr_pred& p, const sr_chunk_view* in, int64_t r0, int64_t n, uint8_t* sel,
                               bool and_merge) {
    const sr_col_view* c = find_col(in, p.slot_id);
    if (!c) return fail(SR_ERR_INVALID_ARGUMENT, "predicate slot not in chunk");
    const bool isf = is_float_class(c->type);
    for (int64_t k = 0; k < n; k++) {
        const int64_t i = r0 + k;
        const bool is_null = c->nulls && c->nulls[i];
        bool r;
        if (p.op == SR_PRED_IS_NULL) {
            r = is_null;
        } else if (p.op == SR_PRED_IS_NOT_NULL) {
            r = !is_null;
        } else if (is_null) {
            r = false;
        } else if (p.op == SR_PRED_IN || p.op == SR_PRED_NOT_IN) {
            const int64_t v = load_int(c->data, c->type, i);
            bool found = false;
            for (int q = 0; q < p.in_count; q++) found |= (p.in_list[q] == v);
            r = (p.op == SR_PRED_IN) ? found : !found;
        } else if (isf) {
            r = pred_cmp<double>(p.op, load_double(c->data, c->type, i), p.dlo, p.dhi);
        } else {
            r = pred_cmp<int64_t>(p.op, load_int(c->data, c->type, i), p.ilo, p.ihi);
        }
        sel[k] = and_merge ? (uint8_t)(sel[k] & (uint8_t)r) : (uint8_t)r;
    }
    return SR_OK;
}

// ---------------------------------------------------------------------------------------
// expressions: column-at-a-time evaluation like the vectorized Expr tree
// (exprs/binary_predicate.cpp:134-143, compound_predicate.cpp, arithmetic_expr.cpp).
// ---------------------------------------------------------------------------------------
struct ExprVal {
    bool is_double = false;
    bool is_bool = false;
    std::vector<int64_t> iv;
    std::vector<double> dv;
    std::vector<uint8_t> nul;
};

static int32_t eval_expr_range(const sr_expr* e, const sr_chunk_view* in, int64_t r0, int64_t n, ExprVal* out) {
    std::vector<ExprVal> st;
    st.reserve(SR_EXPR_STACK);
    for (int k = 0; k < e->num_nodes; k++) {
        const sr_expr_node& nd = e->nodes[k];
        switch (nd.op) {
        case SR_EX_COL: {
            const sr_col_view* c = find_col(in, nd.slot_id);
            if (!c) return fail(SR_ERR_INVALID_ARGUMENT, "expr slot not in chunk");
            ExprVal v;
            v.is_double = is_float_class(c->type);
            v.nul.assign(n, 0);
            if (v.is_double) {
                v.dv.resize(n);
                for (int64_t i = 0; i < n; i++) v.dv[i] = load_double(c->data, c->type, r0 + i);
            } else {
                v.iv.resize(n);
                for (int64_t i = 0; i < n; i++) v.iv[i] = load_int(c->data, c->type, r0 + i);
            }
            if (c->nulls)
                for (int64_t i = 0; i < n; i++) v.nul[i] = c->nulls[r0 + i] ? 1 : 0;
            st.push_back(std::move(v));
            break;
        }
        case SR_EX_ICONST: {
            ExprVal v;
            v.iv.assign(n, nd.ival);
            v.nul.assign(n, 0);
            st.push_back(std::move(v));
            break;
        }
        case SR_EX_DCONST: {
            ExprVal v;
            v.is_double = true;
            v.dv.assign(n, nd.dval);
            v.nul.assign(n, 0);
            st.push_back(std::move(v));
            break;
        }
        case SR_EX_TO_DOUBLE: {
            if (st.empty()) return fail(SR_ERR_INVALID_ARGUMENT, "expr stack underflow");
            ExprVal& v = st.back();
            if (!v.is_double) {
                v.dv.resize(n);
                for (int64_t i = 0; i < n; i++) v.dv[i] = (double)v.iv[i];
                v.iv.clear();
                v.is_double = true;
            }
            break;
        }
        case SR_EX_NOT:
        case SR_EX_IS_NULL: {
            if (st.empty()) return fail(SR_ERR_INVALID_ARGUMENT, "expr stack underflow");
            ExprVal& v = st.back();
            if (nd.op == SR_EX_IS_NULL) {
                v.iv.resize(n);
                for (int64_t i = 0; i < n; i++) {
                    v.iv[i] = v.nul[i];
                    v.nul[i] = 0;
                }
                v.dv.clear();
                v.is_double = false;
            } else {
                if (v.is_double) return fail(SR_ERR_INVALID_ARGUMENT, "NOT on double");
                for (int64_t i = 0; i < n; i++) v.iv[i] = v.iv[i] ? 0 : 1; // NULL stays NULL
            }
            v.is_bool = true;
            break;
        }
        default: {
            if (st.size() < 2) return fail(SR_ERR_INVALID_ARGUMENT, "expr stack underflow");
            ExprVal b = std::move(st.back());
            st.pop_back();
            ExprVal a = std::move(st.back());
            st.pop_back();
            ExprVal r;
            r.nul.resize(n);
            if (nd.op == SR_EX_AND || nd.op == SR_EX_OR) {
                // SQL three-valued logic (compound_predicate.cpp): FALSE AND NULL = FALSE, TRUE OR NULL = TRUE
                r.iv.resize(n);
                r.is_bool = true;
                for (int64_t i = 0; i < n; i++) {
                    const bool an = a.nul[i], bn = b.nul[i];
                    const bool av = !an && a.iv[i] != 0, bv = !bn && b.iv[i] != 0;
                    if (nd.op == SR_EX_AND) {
                        const bool af = !an && !av, bf = !bn && !bv;
                        if (af || bf) {
                            r.iv[i] = 0;
                            r.nul[i] = 0;
                        } else if (an || bn) {
                            r.iv[i] = 0;
                            r.nul[i] = 1;
                        } else {
                            r.iv[i] = 1;
                            r.nul[i] = 0;
                        }
                    } else {
                        if (av || bv) {
                            r.iv[i] = 1;
                            r.nul[i] = 0;
                        } else if (an || bn) {
                            r.iv[i] = 0;
                            r.nul[i] = 1;
                        } else {
                            r.iv[i] = 0;
                            r.nul[i] = 0;
                        }
                    }
                }
                st.push_back(std::move(r));
                break;
            }
            const bool dbl = a.is_double || b.is_double || nd.op == SR_EX_DIV;
            if (dbl) {
                if (!a.is_double) {
                    a.dv.resize(n);
                    for (int64_t i = 0; i < n; i++) a.dv[i] = (double)a.iv[i];
                }
                if (!b.is_double) {
                    b.dv.resize(n);
                    for (int64_t i = 0; i < n; i++) b.dv[i] = (double)b.iv[i];
                }
            }
            for (int64_t i = 0; i < n; i++) r.nul[i] = a.nul[i] | b.nul[i];
            const bool is_cmp = nd.op >= SR_EX_EQ && nd.op <= SR_EX_GE;
            if (is_cmp) {
                r.iv.resize(n);
                r.is_bool = true;
                for (int64_t i = 0; i < n; i++) {
                    bool c;
                    if (dbl) {
                        const double x = a.dv[i], y = b.dv[i];
                        c = nd.op == SR_EX_EQ ? x == y : nd.op == SR_EX_NE ? x != y : nd.op == SR_EX_LT ? x < y
                            : nd.op == SR_EX_LE ? x <= y : nd.op == SR_EX_GT ? x > y : x >= y;
                    } else {
                        const int64_t x = a.iv[i], y = b.iv[i];
                        c = nd.op == SR_EX_EQ ? x == y : nd.op == SR_EX_NE ? x != y : nd.op == SR_EX_LT ? x < y
                            : nd.op == SR_EX_LE ? x <= y : nd.op == SR_EX_GT ? x > y : x >= y;
                    }
                    r.iv[i] = c ? 1 : 0;
                }
            } else if (dbl) {
                r.is_double = true;
                r.dv.resize(n);
                for (int64_t i = 0; i < n; i++) {
                    const double x = a.dv[i], y = b.dv[i];
                    r.dv[i] = nd.op == SR_EX_ADD ? x + y : nd.op == SR_EX_SUB ? x - y : nd.op == SR_EX_MUL ? x * y : x / y;
                    // VectorizedDiv runs under ArithmeticRightZeroCheck (be/src/exprs/arithmetic_operation.h:638,
                    // arithmetic_expr.cpp VectorizedDivArithmeticExpr): a zero divisor yields NULL, for DOUBLE too
                    if (nd.op == SR_EX_DIV && y == 0.0) {
                        r.nul[i] = 1;
                        r.dv[i] = 0.0;
                    }
                }
            } else {
                r.iv.resize(n);
                for (int64_t i = 0; i < n; i++) {
                    const uint64_t x = (uint64_t)a.iv[i], y = (uint64_t)b.iv[i];
                    r.iv[i] = (int64_t)(nd.op == SR_EX_ADD ? x + y : nd.op == SR_EX_SUB ? x - y : x * y);
                }
            }
            st.push_back(std::move(r));
            break;
        }
        }
        if (st.size() > SR_EXPR_STACK) return fail(SR_ERR_INVALID_ARGUMENT, "expr stack overflow");
    }
    if (st.size() != 1) return fail(SR_ERR_INVALID_ARGUMENT, "expr does not reduce to one value");
    *out = std::move(st.back());
    return SR_OK;
}

extern "C" int32_t orc_eval_expr(const sr_expr* e, const sr_chunk_view* in, int64_t* out_i, double* out_d,
                                 uint8_t* out_null, int32_t* is_double) {
    for (int64_t r0 = 0; r0 < in->num_rows || r0 == 0; r0 += ORC_CHUNK_SIZE) {
        const int64_t n = std::min<int64_t>(ORC_CHUNK_SIZE, in->num_rows - r0);
        ExprVal v;
        int32_t rc = eval_expr_range(e, in, r0, n, &v);
        if (rc) return rc;
        *is_double = v.is_double;
        for (int64_t i = 0; i < n; i++) {
            if (v.is_double) {
                if (out_d) out_d[r0 + i] = v.dv[i];
            } else if (out_i) {
                out_i[r0 + i] = v.iv[i];
            }
            if (out_null) out_null[r0 + i] = v.nul[i];
        }
        if (in->num_rows == 0) break;
    }
    return SR_OK;
}

// selection vector for rows [r0, r0+n): ColumnPredicate conjuncts then generic conjuncts;
// a conjunct counts only when true and not null (count_true_with_notnull).
static int32_t scan_select_range(const sr_scan_desc* d, const sr_chunk_view* in, int64_t r0, int64_t n, uint8_t* sel) {
    memset(sel, 1, n);
    for (int k = 0; k < d->num_preds; k++) {
        int32_t rc = eval_pred_range(d->preds[k], in, r0, n, sel, true);
        if (rc) return rc;
    }
    for (int k = 0; k < d->num_filter_exprs; k++) {
        ExprVal v;
        int32_t rc = eval_expr_range(&d->filter_exprs[k], in, r0, n, &v);
        if (rc) return rc;
        if (v.is_double) return fail(SR_ERR_INVALID_ARGUMENT, "filter expr is not boolean");
        for (int64_t i = 0; i < n; i++) sel[i] &= (uint8_t)(!v.nul[i] && v.iv[i] != 0);
    }
    return SR_OK;
}

extern "C" int32_t orc_scan_evaluate(const sr_scan_desc* desc, const sr_chunk_view* in, uint8_t* selection) {
    for (int64_t r0 = 0; r0 < in->num_rows; r0 += ORC_CHUNK_SIZE) {
        const int64_t n = std::min<int64_t>(ORC_CHUNK_SIZE, in->num_rows - r0);
        int32_t rc = scan_select_range(desc, in, r0, n, selection + r0);
        if (rc) return rc;
    }
    return SR_OK;
}

extern "C" int64_t orc_scan_filter(const sr_scan_desc* desc, const sr_chunk_view* in, void** out_data,
                                   uint8_t** out_nulls) {
    std::vector<uint8_t> sel(ORC_CHUNK_SIZE);
    int64_t out_rows = 0;
    for (int64_t r0 = 0; r0 < in->num_rows; r0 += ORC_CHUNK_SIZE) {
        const int64_t n = std::min<int64_t>(ORC_CHUNK_SIZE, in->num_rows - r0);
        int32_t rc = scan_select_range(desc, in, r0, n, sel.data());
        if (rc) return rc;
        int64_t kept = 0;
        for (int k = 0; k < desc->num_out_slots; k++) {
            const sr_col_view* c = find_col(in, desc->out_slots[k]);
            if (!c) return fail(SR_ERR_INVALID_ARGUMENT, "out slot not in chunk");
            const int w = type_width(c->type);
            // Chunk::filter -> Column::filter_range: copy the chunk then compact in place
            uint8_t* dst = (uint8_t*)out_data[k] + out_rows * w;
            memcpy(dst, (const uint8_t*)c->data + r0 * w, n * w);
            kept = orc_filter_range(sel.data(), dst, w, 0, n);
            if (c->nulls && out_nulls && out_nulls[k]) {
                uint8_t* nd = out_nulls[k] + out_rows;
                memcpy(nd, c->nulls + r0, n);
                orc_filter_range(sel.data(), nd, 1, 0, n);
            }
        }
        if (desc->num_out_slots == 0)
            for (int64_t i = 0; i < n; i++) kept += sel[i];
        out_rows += kept;
    }
    return out_rows;
}

// ---------------------------------------------------------------------------------------
// hash join
// ---------------------------------------------------------------------------------------
struct OwnedCol {
    int32_t type = 0;
    int32_t slot = 0;
    bool nullable = false;
    bool has_null = false;
    std::vector<uint8_t> data; // rows * width, row 0 = sentinel
    std::vector<uint8_t> nulls;
};

struct DenseGroup {
    uint32_t start_index = 0;
    uint32_t bitset = 0;
};

// HashTableProbeState (join_hash_table_descriptor.h:207-330)
struct ProbeState {
    std::vector<uint32_t> p_next, p_buckets;
    std::vector<int64_t> p_keys;
    std::vector<uint8_t> p_nulls, p_match_filter;
    bool p_has_null = false;
    uint32_t probe_row_count = 0;
    uint32_t cur_probe_index = 0, cur_build_index = 0, cur_row_match_count = 0;
    bool has_remain = false;
    void prepare(uint32_t chunk_size) { // JoinHashMap::probe_prepare (join_hash_map.hpp:36-52)
        p_next.assign(chunk_size, 0);
        p_buckets.assign(chunk_size, 0);
        p_keys.assign(chunk_size, 0);
        p_nulls.assign(chunk_size, 0);
        p_match_filter.assign(chunk_size, 0);
    }
};

struct orc_join {
    sr_join_desc desc;
    orc_join_options opt;
    int32_t chunk_size = ORC_CHUNK_SIZE;
    // JoinHashTableItems (join_hash_table_descriptor.h:105-205)
    std::vector<OwnedCol> build_cols; // build_chunk, row 0 reserved
    uint32_t row_count = 0;
    std::vector<int64_t> keys; // packed build keys, index 0 = sentinel (value 0)
    std::vector<uint8_t> key_nulls;
    bool keys_have_null = false;
    int key_bytes = 0; // width the hash is computed on (1,2,4,8)
    int32_t method = 0;
    uint32_t bucket_size = 0;
    uint32_t log_bucket_size = 0;
    int64_t min_value = 0, max_value = 0;
    std::vector<uint32_t> first, next;
    std::vector<uint8_t> key_bitset;
    std::vector<DenseGroup> dense_groups;
    bool built = false;
    ProbeState ps; // the builder's own probe state; probers of a fragment clone one each
    // HashTableProbeState::build_match_index (join_hash_map.hpp:47-48): 1 = the build row found a probe partner.  Index 0 is
    // the sentinel row, always 1.  One array for the whole join: the reference merges the probers' arrays before POST_PROBE.
    mutable std::vector<uint8_t> build_match;
    bool wide = false;
    std::map<std::pair<uint64_t, uint64_t>, int64_t> wide_ids; // 128-bit build key -> id
};

static const int FP_BITS = 7; // join_hash_map_method.h:137-147

static inline uint32_t key_hash(const orc_join* j, int64_t key, uint32_t log_buckets) {
    // calc_bucket_num<CppType>: 4-byte and 8-byte keys use multiplicative hashing; narrower
    // keys are only ever DIRECT_MAPPING (join_hash_table.cpp:230).
    if (j->key_bytes == 8) return orc_join_key_hash64((uint64_t)key, log_buckets);
    return orc_join_key_hash32((uint32_t)(int32_t)key, log_buckets);
}

extern "C" orc_join* orc_join_create(const sr_join_desc* desc, const orc_join_options* opt) {
    if (desc->num_keys < 1 || desc->num_keys > SR_MAX_JOIN_KEYS) {
        fail(SR_ERR_INVALID_ARGUMENT, "num_keys");
        return nullptr;
    }
    int total = 0;
    for (int k = 0; k < desc->num_keys; k++) {
        if (is_float_class(desc->key_types[k]) || type_width(desc->key_types[k]) > 8 ||
            type_width(desc->key_types[k]) == 0) {
            fail(SR_ERR_NOT_SUPPORTED, "join key type");
            return nullptr;
        }
        total += type_width(desc->key_types[k]);
    }
    if (total > 16) {
        fail(SR_ERR_NOT_SUPPORTED, "packed join key wider than 16 bytes");
        return nullptr;
    }
    auto* j = new orc_join();
    j->desc = *desc;
    j->wide = total > 8;
    if (opt) {
        j->opt = *opt;
    } else {
        j->opt.enable_range_direct_mapping = desc->enable_range_direct_mapping;
        j->opt.enable_linear_chained = 1;
        j->opt.l2_cache_size = 1 << 20;
        j->opt.l3_cache_size = 32 << 20;
        j->opt.force_method = 0;
        j->opt.chunk_size = 0;
    }
    if (j->opt.chunk_size > 0) j->chunk_size = j->opt.chunk_size;
    // ONE_KEY keeps the native width; several keys are serialized into the smallest of
    // int32/int64 that fits (SERIALIZED_FIXED_SIZE_INT/BIGINT, join_hash_table.cpp:225-250)
    j->key_bytes = desc->num_keys == 1 ? type_width(desc->key_types[0]) : (total <= 4 ? 4 : 8);
    // 9..16 bytes: SERIALIZED_FIXED_SIZE_LARGEINT (join_hash_table.cpp:221-222).  The table machinery below works on 64-bit
    // keys; a 128-bit key is INTERNED -- every distinct build key gets the next id, the id is the table key.  The pair
    // sequence of a probe (probe order, chain by descending build index) does not depend on which hash placed the key.
    j->keys.push_back(0);
    j->key_nulls.push_back(0);
    return j;
}
extern "C" void orc_join_destroy(orc_join* j) {
    delete j;
}

// pack key columns of `c` rows [r0,r0+n) into int64 (little-endian concatenation, the
// SERIALIZED_FIXED_SIZE layout, join_key_constructor.hpp) + null flags.
static int32_t pack_keys(const orc_join* j, const sr_chunk_view* c, const int32_t* slots, int64_t r0, int64_t n,
                         int64_t* out, uint8_t* out_null, bool* any_null, std::map<std::pair<uint64_t, uint64_t>, int64_t>* intern = nullptr) {
    const sr_join_desc& d = j->desc;
    *any_null = false;
    if (j->wide) {
        std::vector<uint64_t> lo(n, 0), hi(n, 0);
        for (int64_t i = 0; i < n; i++) out_null[i] = 0;
        int shift = 0;
        for (int k = 0; k < d.num_keys; k++) {
            const sr_col_view* col = find_col(c, slots[k]);
            if (!col) return fail(SR_ERR_INVALID_ARGUMENT, "join key slot not in chunk");
            const int w = type_width(col->type);
            const uint64_t mask = w == 8 ? ~0ull : ((1ull << (8 * w)) - 1);
            for (int64_t i = 0; i < n; i++) {
                const uint64_t v = (uint64_t)load_int(col->data, col->type, r0 + i) & mask;
                if (shift < 64) {
                    lo[i] |= v << shift;
                    if (shift + 8 * w > 64) hi[i] |= v >> (64 - shift);
                } else {
                    hi[i] |= v << (shift - 64);
                }
                if (col->nulls && col->nulls[r0 + i]) {
                    out_null[i] = 1;
                    *any_null = true;
                }
            }
            shift += 8 * w;
        }
        for (int64_t i = 0; i < n; i++) {
            const auto key = std::make_pair(lo[i], hi[i]);
            if (intern) { // build side: a new key gets the next id
                auto it = intern->find(key);
                if (it == intern->end()) it = intern->emplace(key, (int64_t)intern->size()).first;
                out[i] = it->second;
            } else { // probe side: a key the build side never had gets an id no build row carries
                auto it = j->wide_ids.find(key);
                out[i] = it == j->wide_ids.end() ? -1 - i : it->second;
            }
        }
        return SR_OK;
    }
    if (d.num_keys == 1) {
        const sr_col_view* col = find_col(c, slots[0]);
        if (!col) return fail(SR_ERR_INVALID_ARGUMENT, "join key slot not in chunk");
        for (int64_t i = 0; i < n; i++) {
            out[i] = load_int(col->data, col->type, r0 + i);
            const uint8_t nu = col->nulls ? (col->nulls[r0 + i] ? 1 : 0) : 0;
            out_null[i] = nu;
            *any_null |= nu != 0;
        }
        return SR_OK;
    }
    for (int64_t i = 0; i < n; i++) {
        out[i] = 0;
        out_null[i] = 0;
    }
    int shift = 0;
    for (int k = 0; k < d.num_keys; k++) {
        const sr_col_view* col = find_col(c, slots[k]);
        if (!col) return fail(SR_ERR_INVALID_ARGUMENT, "join key slot not in chunk");
        const int w = type_width(col->type);
        const uint64_t mask = w == 8 ? ~0ull : ((1ull << (8 * w)) - 1);
        for (int64_t i = 0; i < n; i++) {
            const uint64_t v = (uint64_t)load_int(col->data, col->type, r0 + i) & mask;
            out[i] |= (int64_t)(v << shift);
            if (col->nulls && col->nulls[r0 + i]) {
                out_null[i] = 1;
                *any_null = true;
            }
        }
        shift += 8 * w;
    }
    return SR_OK;
}

// JoinHashTable::append_chunk: join_hash_table.cpp:712-752 (row 0 of every build column is a
// default value; nullable columns get NULL there)
extern "C" int32_t orc_join_append_build(orc_join* j, const sr_chunk_view* chunk) {
    if (j->built) return fail(SR_ERR_STATE, "append after build");
    if (j->build_cols.empty()) {
        for (int k = 0; k < chunk->num_cols; k++) {
            OwnedCol oc;
            oc.type = chunk->cols[k].type;
            oc.slot = chunk->cols[k].slot_id;
            oc.nullable = chunk->cols[k].nulls != nullptr;
            oc.data.assign(type_width(oc.type), 0);
            oc.nulls.assign(1, oc.nullable ? 1 : 0);
            j->build_cols.push_back(std::move(oc));
        }
    }
    const int64_t n = chunk->num_rows;
    for (auto& oc : j->build_cols) {
        const sr_col_view* c = find_col(chunk, oc.slot);
        if (!c) return fail(SR_ERR_INVALID_ARGUMENT, "build chunk misses a column");
        const int w = type_width(oc.type);
        const size_t old = oc.data.size();
        oc.data.resize(old + n * w);
        memcpy(oc.data.data() + old, c->data, n * w);
        const size_t oldn = oc.nulls.size();
        oc.nulls.resize(oldn + n, 0);
        if (c->nulls) {
            // upgrade to nullable when a later chunk carries a null column (:726-742)
            oc.nullable = true;
            for (int64_t i = 0; i < n; i++) {
                oc.nulls[oldn + i] = c->nulls[i] ? 1 : 0;
                oc.has_null |= c->nulls[i] != 0;
            }
        }
    }
    const size_t oldk = j->keys.size();
    j->keys.resize(oldk + n);
    j->key_nulls.resize(oldk + n);
    bool any = false;
    int32_t rc = pack_keys(j, chunk, j->desc.build_key_slots, 0, n, j->keys.data() + oldk, j->key_nulls.data() + oldk, &any, j->wide ? &j->wide_ids : nullptr);
    if (rc) return rc;
    j->keys_have_null |= any;
    j->row_count += (uint32_t)n;
    return SR_OK;
}

static void build_bucket_chained(orc_join* j) {
    // BucketChainedJoinHashMap::build_prepare + construct_hash_table: join_hash_map_method.hpp:29-86
    j->bucket_size = orc_calc_bucket_size(j->row_count + 1);
    j->log_bucket_size = __builtin_ctz(j->bucket_size);
    j->first.assign(j->bucket_size, 0);
    j->next.assign(j->row_count + 1, 0);
    const uint32_t num_rows = 1 + j->row_count;
    for (uint32_t i = 1; i < num_rows; i++) {
        if (j->keys_have_null && j->key_nulls[i]) {
            j->next[i] = 0;
            continue;
        }
        const uint32_t b = key_hash(j, j->keys[i], j->log_bucket_size);
        j->next[i] = j->first[b];
        j->first[b] = i;
    }
}

static void build_linear_chained(orc_join* j, bool need_chain) {
    // TLinearChainedJoinHashMap: join_hash_map_method.hpp:133-260; fingerprint packing
    // first = (fp << 24) | index, join_hash_map_method.h:137-147
    j->bucket_size = orc_calc_bucket_size(j->row_count + 1);
    j->log_bucket_size = __builtin_ctz(j->bucket_size);
    j->first.assign(j->bucket_size, 0);
    j->next.assign(j->row_count + 1, 0);
    const uint32_t mask = j->bucket_size - 1;
    const uint32_t num_rows = 1 + j->row_count;
    for (uint32_t i = 1; i < num_rows; i++) {
        if (j->keys_have_null && j->key_nulls[i]) {
            j->next[i] = 0;
            continue;
        }
        const uint32_t hash = key_hash(j, j->keys[i], j->log_bucket_size + FP_BITS);
        const uint32_t fp = (hash & 0x7F) | 0x80;
        uint32_t b = hash >> FP_BITS;
        uint32_t probe_times = 1;
        while (true) {
            if (j->first[b] == 0) {
                j->next[i] = 0;
                j->first[b] = (fp << 24) | i;
                break;
            }
            const uint32_t cur_fp = j->first[b] >> 24, cur = j->first[b] & 0xFFFFFF;
            if (fp == cur_fp && j->keys[i] == j->keys[cur]) {
                if (need_chain) {
                    j->next[i] = cur;
                    j->first[b] = (fp << 24) | i;
                }
                break;
            }
            b = (b + probe_times) & mask;
            probe_times++;
        }
    }
}

static void build_direct(orc_join* j, int64_t min_value, uint64_t interval) {
    // DirectMappingJoinHashMap / RangeDirectMappingJoinHashMap::construct_hash_table:
    // join_hash_map_method.hpp:542-590, 620-674
    j->bucket_size = (uint32_t)interval;
    j->first.assign(interval, 0);
    j->next.assign(j->row_count + 1, 0);
    const uint32_t num_rows = 1 + j->row_count;
    for (uint32_t i = 1; i < num_rows; i++) {
        if (j->keys_have_null && j->key_nulls[i]) continue;
        const uint64_t b = (uint64_t)(j->keys[i] - min_value);
        j->next[i] = j->first[b];
        j->first[b] = i;
    }
}

static void build_range_set(orc_join* j) {
    // RangeDirectMappingJoinHashSet: join_hash_map_method.hpp:720-746
    const uint64_t interval = (uint64_t)j->max_value - j->min_value + 1;
    j->bucket_size = (uint32_t)((interval + 7) / 8);
    j->key_bitset.assign(j->bucket_size, 0);
    for (uint32_t i = 1; i < 1 + j->row_count; i++) {
        const uint64_t b = (uint64_t)(j->keys[i] - j->min_value);
        const bool ok = !(j->keys_have_null && j->key_nulls[i]);
        j->key_bitset[b / 8] |= (uint8_t)((ok ? 1 : 0) << (b % 8));
    }
}

static inline uint32_t dense_slot(const orc_join* j, int64_t key) {
    const uint32_t b = (uint32_t)(key - j->min_value);
    const DenseGroup& g = j->dense_groups[b / 32];
    return g.start_index + __builtin_popcount(g.bitset & ((1u << (b % 32)) - 1));
}

static void build_dense_range(orc_join* j) {
    // DenseRangeDirectMappingJoinHashMap: join_hash_map_method.hpp:790-905
    const uint64_t interval = (uint64_t)j->max_value - j->min_value + 1;
    j->bucket_size = j->row_count + 1;
    j->dense_groups.assign((interval + 31) / 32, DenseGroup());
    j->first.assign(j->row_count + 1, 0);
    j->next.assign(j->row_count + 1, 0);
    const uint32_t num_rows = j->row_count + 1;
    for (uint32_t r = 1; r < num_rows; r++) {
        if (j->keys_have_null && j->key_nulls[r]) continue;
        const uint32_t b = (uint32_t)(j->keys[r] - j->min_value);
        j->dense_groups[b / 32].bitset |= 1u << (b % 32);
    }
    uint32_t start = 0;
    for (auto& g : j->dense_groups) {
        g.start_index = start;
        start += __builtin_popcount(g.bitset);
    }
    for (uint32_t r = 1; r < num_rows; r++) {
        if (j->keys_have_null && j->key_nulls[r]) continue;
        const uint32_t idx = dense_slot(j, j->keys[r]);
        j->next[r] = j->first[idx];
        j->first[idx] = r;
    }
}

// JoinHashMapSelector::_determine_hash_map_method: join_hash_table.cpp:225-350
static int32_t select_method(orc_join* j) {
    if (j->opt.force_method) {
        // tests force a family; the range families still need min/max of the build keys
        if (j->row_count > 0) {
            int64_t mn = std::numeric_limits<int64_t>::max(), mx = std::numeric_limits<int64_t>::min();
            for (uint32_t i = 1; i <= j->row_count; i++) {
                mn = std::min(mn, j->keys[i]);
                mx = std::max(mx, j->keys[i]);
            }
            j->min_value = mn;
            j->max_value = mx;
        }
        return j->opt.force_method;
    }
    // the key-set methods keep no build rows: they are only chosen when no other-join conjunct has to read them
    // (JoinHashTableItems::with_other_conjunct, join_hash_table.cpp:259-300)
    const bool semi_anti = (j->desc.join_type == SR_JOIN_LEFT_SEMI || j->desc.join_type == SR_JOIN_LEFT_ANTI) && j->desc.other_conjunct.num_nodes == 0;
    const bool one_key = j->desc.num_keys == 1;
    if (one_key && j->key_bytes <= 2) return ORC_DIRECT_MAPPING;
    if (one_key && j->opt.enable_range_direct_mapping && j->row_count > 0) {
        int64_t mn = std::numeric_limits<int64_t>::max(), mx = std::numeric_limits<int64_t>::min();
        for (uint32_t i = 1; i <= j->row_count; i++) { // min/max over all rows incl. null rows' payload value
            mn = std::min(mn, j->keys[i]);
            mx = std::max(mx, j->keys[i]);
        }
        if (!(mn == std::numeric_limits<int64_t>::min() && mx == std::numeric_limits<int64_t>::max())) {
            const uint64_t interval = (uint64_t)mx - (uint64_t)mn + 1;
            if (interval < std::numeric_limits<uint32_t>::max()) {
                j->min_value = mn;
                j->max_value = mx;
                const uint64_t bucket_size = orc_calc_bucket_size(j->row_count + 1);
                const uint64_t row_count = j->row_count;
                if (semi_anti) {
                    const uint64_t mem = (interval + 7) / 8;
                    if (mem <= bucket_size * 64 || mem <= (uint64_t)j->opt.l3_cache_size / 2)
                        return ORC_RANGE_DIRECT_MAPPING_SET;
                } else {
                    if (interval <= bucket_size || interval <= (uint64_t)j->opt.l2_cache_size)
                        return ORC_RANGE_DIRECT_MAPPING;
                    if (interval / 4 + row_count * 4 <= (bucket_size + bucket_size / 10) * 4)
                        return ORC_DENSE_RANGE_DIRECT_MAPPING;
                }
            }
        }
    }
    if (j->opt.enable_linear_chained) {
        const uint64_t bucket_size = orc_calc_bucket_size(j->row_count + 1);
        if (bucket_size <= (1u << 24)) return semi_anti ? ORC_LINEAR_CHAINED_SET : ORC_LINEAR_CHAINED;
    }
    return ORC_BUCKET_CHAINED;
}

extern "C" int32_t orc_join_build(orc_join* j) {
    if (j->built) return SR_OK;
    j->method = select_method(j);
    switch (j->method) {
    case ORC_BUCKET_CHAINED:
        build_bucket_chained(j);
        break;
    case ORC_LINEAR_CHAINED:
        build_linear_chained(j, true);
        break;
    case ORC_LINEAR_CHAINED_SET:
        build_linear_chained(j, false);
        break;
    case ORC_DIRECT_MAPPING: {
        const int64_t mn = j->key_bytes == 1 ? (j->desc.key_types[0] == SR_TYPE_BOOLEAN ? 0 : -128) : -32768;
        const uint64_t interval = j->key_bytes == 1 ? (j->desc.key_types[0] == SR_TYPE_BOOLEAN ? 2 : 256) : 65536;
        j->min_value = mn;
        j->max_value = mn + (int64_t)interval - 1;
        build_direct(j, mn, interval);
        break;
    }
    case ORC_RANGE_DIRECT_MAPPING:
        build_direct(j, j->min_value, (uint64_t)j->max_value - j->min_value + 1);
        break;
    case ORC_RANGE_DIRECT_MAPPING_SET:
        build_range_set(j);
        break;
    case ORC_DENSE_RANGE_DIRECT_MAPPING:
        build_dense_range(j);
        break;
    default:
        return fail(SR_ERR_INVALID_ARGUMENT, "bad join method");
    }
    j->ps.prepare(j->chunk_size);
    j->build_match.assign((size_t)j->row_count + 1, 0);
    j->build_match[0] = 1;
    j->built = true;
    return SR_OK;
}

extern "C" int32_t orc_join_method(const orc_join* j) {
    return j->method;
}
extern "C" int64_t orc_join_build_rows(const orc_join* j) {
    return j->row_count;
}
extern "C" int64_t orc_join_bucket_size(const orc_join* j) {
    return j->bucket_size;
}
extern "C" int64_t orc_join_min_value(const orc_join* j) {
    return j->min_value;
}
extern "C" int64_t orc_join_max_value(const orc_join* j) {
    return j->max_value;
}
extern "C" const uint32_t* orc_join_first(const orc_join* j) {
    return j->first.data();
}
extern "C" const uint32_t* orc_join_next(const orc_join* j) {
    return j->next.data();
}

// lookup_init of each method (join_hash_map_method.hpp:87-126, 300-370, 592-618, 676-706,
// 748-783, 907-957): probe_state.next[i] = head build row of probe row i (or 0)
static void lookup_init(const orc_join* j, ProbeState& ps) {
    const uint32_t n = ps.probe_row_count;
    const bool hn = ps.p_has_null;
    switch (j->method) {
    case ORC_BUCKET_CHAINED:
        for (uint32_t i = 0; i < n; i++) {
            ps.p_buckets[i] = key_hash(j, ps.p_keys[i], j->log_bucket_size);
            ps.p_next[i] = (hn && ps.p_nulls[i]) ? 0 : j->first[ps.p_buckets[i]];
        }
        break;
    case ORC_LINEAR_CHAINED:
    case ORC_LINEAR_CHAINED_SET: {
        const uint32_t mask = j->bucket_size - 1;
        for (uint32_t i = 0; i < n; i++) {
            if (hn && ps.p_nulls[i]) {
                ps.p_next[i] = 0;
                continue;
            }
            const uint32_t hash = key_hash(j, ps.p_keys[i], j->log_bucket_size + FP_BITS);
            const uint32_t fp = (hash & 0x7F) | 0x80;
            uint32_t b = hash >> FP_BITS, probe_times = 1;
            while (true) {
                if (j->first[b] == 0) {
                    ps.p_next[i] = 0;
                    break;
                }
                const uint32_t cur_fp = j->first[b] >> 24, cur = j->first[b] & 0xFFFFFF;
                if (fp == cur_fp && ps.p_keys[i] == j->keys[cur]) {
                    ps.p_next[i] = j->method == ORC_LINEAR_CHAINED ? cur : 1;
                    break;
                }
                b = (b + probe_times) & mask;
                probe_times++;
            }
        }
        break;
    }
    case ORC_DIRECT_MAPPING:
    case ORC_RANGE_DIRECT_MAPPING:
        for (uint32_t i = 0; i < n; i++) {
            const int64_t k = ps.p_keys[i];
            const bool ok = !(hn && ps.p_nulls[i]) && k >= j->min_value && k <= j->max_value;
            ps.p_next[i] = ok ? j->first[(uint64_t)(k - j->min_value)] : 0;
        }
        break;
    case ORC_RANGE_DIRECT_MAPPING_SET:
        for (uint32_t i = 0; i < n; i++) {
            const int64_t k = ps.p_keys[i];
            const bool ok = !(hn && ps.p_nulls[i]) && k >= j->min_value && k <= j->max_value;
            if (ok) {
                const uint64_t idx = (uint64_t)(k - j->min_value);
                ps.p_next[i] = (j->key_bitset[idx / 8] & (1 << (idx % 8))) != 0;
            } else {
                ps.p_next[i] = 0;
            }
        }
        break;
    case ORC_DENSE_RANGE_DIRECT_MAPPING:
        for (uint32_t i = 0; i < n; i++) {
            const int64_t k = ps.p_keys[i];
            const bool ok = !(hn && ps.p_nulls[i]) && k >= j->min_value && k <= j->max_value;
            uint32_t r = 0;
            if (ok) {
                const uint64_t b = (uint64_t)(k - j->min_value);
                const DenseGroup& g = j->dense_groups[b / 32];
                if (g.bitset & (1u << (b % 32)))
                    r = j->first[g.start_index + __builtin_popcount(g.bitset & ((1u << (b % 32)) - 1))];
            }
            ps.p_next[i] = r;
        }
        break;
    }
}

static inline bool is_set_method(int32_t m) {
    return m == ORC_RANGE_DIRECT_MAPPING_SET || m == ORC_LINEAR_CHAINED_SET;
}

static inline uint32_t count_zero(const uint8_t* p, uint32_t n) {
    uint32_t z = 0;
    for (uint32_t i = 0; i < n; i++) z += p[i] == 0;
    return z;
}

// _contains_probe_row: join_hash_map.hpp:1168-1184
static inline bool contains_probe_row(const orc_join* j, const ProbeState& ps, uint32_t i) {
    uint32_t idx = ps.p_next[i];
    if (idx == 0) return false;
    if (is_set_method(j->method)) return true;
    do {
        if (j->keys[idx] == ps.p_keys[i]) return true;
        idx = j->next[idx];
    } while (idx != 0);
    return false;
}

static int32_t probe_chunk_impl(const orc_join* j, ProbeState& ps, const int32_t* probe_key_slots,
                                const sr_chunk_view* probe, int32_t first_probe, uint32_t* probe_index,
                                uint32_t* build_index, orc_probe_result* res, int32_t join_type_override = -1) {
    if (!j->built) return fail(SR_ERR_STATE, "probe before build");
    if (probe->num_rows > j->chunk_size) return fail(SR_ERR_INVALID_ARGUMENT, "probe chunk larger than chunk_size");
    const uint32_t chunk_size = j->chunk_size;
    if (first_probe) {
        // _search_ht: build probe keys, lookup_init (join_hash_map.hpp:381-420)
        ps.probe_row_count = (uint32_t)probe->num_rows;
        bool any = false;
        int32_t rc = pack_keys(j, probe, probe_key_slots, 0, probe->num_rows, ps.p_keys.data(), ps.p_nulls.data(), &any);
        if (rc) return rc;
        ps.p_has_null = any;
        ps.cur_probe_index = 0;
        ps.cur_build_index = 0;
        ps.cur_row_match_count = 0;
        ps.has_remain = false;
        lookup_init(j, ps);
    } else if (!ps.has_remain) {
        return fail(SR_ERR_STATE, "no remaining probe state");
    }
    int32_t match_flag = 0;
    size_t match_count = 0;
    bool one_to_many = false;
    const uint32_t n = ps.probe_row_count;
    auto over = [&]() { // PROBE_OVER
        ps.has_remain = false;
        ps.cur_probe_index = 0;
        ps.cur_build_index = 0;
        res->count = (int64_t)match_count;
        ps.cur_row_match_count = 0;
    };
    const int32_t jt = join_type_override >= 0 ? join_type_override : j->desc.join_type;
    // joins with a POST_PROBE phase mark every build row they match (join_hash_map.hpp:1352 right outer, :1492 right anti,
    // :1600 full outer)
    const bool marks = jt == SR_JOIN_RIGHT_OUTER || jt == SR_JOIN_FULL_OUTER || jt == SR_JOIN_RIGHT_SEMI || jt == SR_JOIN_RIGHT_ANTI;
    if (jt == SR_JOIN_RIGHT_SEMI || jt == SR_JOIN_RIGHT_ANTI) {
        // _probe_from_ht_for_right_anti_join (join_hash_map.hpp:1480-1500): mark, emit nothing.  RIGHT SEMI marks the same
        // way here and emits its (matched) build rows from probe_remain in build order; the reference emits each of them
        // during the probe at its first match (:1405-1440) -- the same rows in another order.
        for (uint32_t i = 0; i < n; i++)
            for (uint32_t b = ps.p_next[i]; b != 0; b = j->next[b])
                if (j->keys[b] == ps.p_keys[i]) j->build_match[b] = 1;
        over();
    } else if (jt == SR_JOIN_LEFT_SEMI || jt == SR_JOIN_LEFT_ANTI) {
        // _probe_from_ht_for_left_semi_join / left_anti_join: join_hash_map.hpp:1186-1255
        for (uint32_t i = 0; i < n; i++) {
            const bool c = contains_probe_row(j, ps, i);
            if (c == (jt == SR_JOIN_LEFT_SEMI)) {
                probe_index[match_count] = i;
                build_index[match_count] = 0;
                match_count++;
            }
        }
        if (match_count == n)
            match_flag = 1;
        else if (match_count * 2 >= n)
            match_flag = 2;
        over();
    } else {
        // INNER: _probe_from_ht (join_hash_map.hpp:718-795); LEFT OUTER:
        // _probe_from_ht_for_left_outer_join (:950-1030).  Chunk-full resume via
        // RETURN_IF_CHUNK_FULL2 (:593-602).
        // RIGHT OUTER probes like INNER, FULL OUTER like LEFT OUTER (_probe_from_ht_for_right_outer_join :1330-1375,
        // _probe_from_ht_for_full_outer_join :1560-1640), both marking the build rows they emit
        const bool outer = jt == SR_JOIN_LEFT_OUTER || jt == SR_JOIN_FULL_OUTER;
        size_t i = ps.cur_probe_index;
        if (!first_probe) {
            probe_index[0] = ps.cur_probe_index;
            build_index[0] = ps.cur_build_index;
            match_count = 1;
            if (ps.p_next[i] == 0) {
                i++;
                ps.cur_row_match_count = 0;
            }
        }
        if (first_probe && !outer) memset(ps.p_match_filter.data(), 0, n);
        uint32_t cur_row_match_count = ps.cur_row_match_count;
        bool returned = false;
        for (; i < n && !returned; i++) {
            uint32_t b = ps.p_next[i];
            if (b == 0) {
                if (outer) {
                    probe_index[match_count] = (uint32_t)i;
                    build_index[match_count] = 0;
                    match_count++;
                    if (match_count > chunk_size) {
                        ps.p_next[i] = j->next[0];
                        ps.cur_probe_index = (uint32_t)i;
                        ps.cur_build_index = 0;
                        ps.has_remain = true;
                        res->count = chunk_size;
                        ps.cur_row_match_count = cur_row_match_count;
                        returned = true;
                    }
                }
                continue;
            }
            do {
                if (j->keys[b] == ps.p_keys[i]) {
                    probe_index[match_count] = (uint32_t)i;
                    build_index[match_count] = b;
                    if (marks) j->build_match[b] = 1;
                    match_count++;
                    if (first_probe || outer) cur_row_match_count++;
                    if (first_probe && !outer) ps.p_match_filter[i] = 1;
                    if (match_count > chunk_size) {
                        ps.p_next[i] = j->next[b];
                        ps.cur_probe_index = (uint32_t)i;
                        ps.cur_build_index = b;
                        ps.has_remain = true;
                        res->count = chunk_size;
                        ps.cur_row_match_count = cur_row_match_count;
                        returned = true;
                        break;
                    }
                }
                b = j->next[b];
            } while (b != 0);
            if (returned) break;
            if (outer && cur_row_match_count <= 0) {
                probe_index[match_count] = (uint32_t)i;
                build_index[match_count] = 0;
                match_count++;
                if (match_count > chunk_size) {
                    ps.p_next[i] = 0;
                    ps.cur_probe_index = (uint32_t)i;
                    ps.cur_build_index = 0;
                    ps.has_remain = true;
                    res->count = chunk_size;
                    ps.cur_row_match_count = cur_row_match_count;
                    returned = true;
                    break;
                }
            } else if (first_probe) {
                one_to_many |= cur_row_match_count > 1;
            }
            if (first_probe || outer) cur_row_match_count = 0;
        }
        if (!returned) {
            ps.cur_row_match_count = cur_row_match_count;
            if (first_probe && match_count > 0 && !one_to_many) {
                if (outer) {
                    match_flag = 1; // CHECK_ALL_MATCH
                } else {          // CHECK_MATCH
                    const uint32_t z = count_zero(ps.p_match_filter.data(), n);
                    if (z == 0)
                        match_flag = 1;
                    else if (z < n - z)
                        match_flag = 2;
                }
            }
            over();
        }
    }
    res->has_remain = ps.has_remain;
    res->match_flag = match_flag;
    res->cur_probe_index = (int32_t)ps.cur_probe_index;
    res->cur_row_match_count = (int32_t)ps.cur_row_match_count;
    return SR_OK;
}

extern "C" int32_t orc_join_probe_chunk(orc_join* j, const sr_chunk_view* probe, int32_t first_probe,
                                        uint32_t* probe_index, uint32_t* build_index, orc_probe_result* res) {
    return probe_chunk_impl(j, j->ps, j->desc.probe_key_slots, probe, first_probe, probe_index, build_index, res);
}

// Other-join conjunct (HashJoiner::_other_join_conjunct_ctxs, exec/hash_joiner.h:314-329; hash_joiner.cpp
// _process_other_conjunct / _process_outer_join_with_other_conjunct / _process_semi_join_with_other_conjunct /
// _process_right_anti_join_with_other_conjunct ...): the key match yields candidate pairs (the INNER probe), the conjunct
// is evaluated over each pair's probe row and build row, and the join type decides what a probe row emits: its passing
// pairs (INNER, RIGHT OUTER), those or one NULL-padded row when none passes (LEFT / FULL OUTER), itself once when some /
// no pair passes (LEFT SEMI / ANTI), nothing (RIGHT SEMI / ANTI); RIGHT / FULL joins mark the build rows of passing pairs.
static int64_t probe_all_with_conjunct(orc_join* j, const sr_chunk_view* probe, uint32_t* probe_index, uint32_t* build_index, int64_t cap) {
    std::vector<uint32_t> cpi, cbi;
    {
        std::vector<uint32_t> pi(j->chunk_size + 8), bi(j->chunk_size + 8);
        std::vector<sr_col_view> cols(probe->num_cols);
        for (int64_t r0 = 0; r0 < probe->num_rows; r0 += j->chunk_size) {
            const int64_t n = std::min<int64_t>(j->chunk_size, probe->num_rows - r0);
            for (int k = 0; k < probe->num_cols; k++) {
                cols[k] = probe->cols[k];
                cols[k].data = (const uint8_t*)cols[k].data + r0 * type_width(cols[k].type);
                if (cols[k].nulls) cols[k].nulls += r0;
            }
            sr_chunk_view sub{cols.data(), probe->num_cols, SR_MEM_HOST, n};
            int32_t first = 1;
            while (true) {
                orc_probe_result res{};
                int32_t rc = probe_chunk_impl(j, j->ps, j->desc.probe_key_slots, &sub, first, pi.data(), bi.data(), &res, SR_JOIN_INNER);
                if (rc) return rc;
                for (int64_t q = 0; q < res.count; q++) {
                    cpi.push_back((uint32_t)(pi[q] + r0));
                    cbi.push_back(bi[q]);
                }
                if (!res.has_remain) break;
                first = 0;
            }
        }
    }
    const int64_t nc = (int64_t)cpi.size();
    // the candidate pairs as a chunk of the columns the conjunct reads
    const sr_expr& e = j->desc.other_conjunct;
    std::vector<std::vector<uint8_t>> data, nulls;
    std::vector<sr_col_view> pcols;
    for (int k = 0; k < e.num_nodes; k++) {
        if (e.nodes[k].op != SR_EX_COL) continue;
        const int32_t slot = e.nodes[k].slot_id;
        bool seen = false;
        for (auto& c : pcols) seen |= c.slot_id == slot;
        if (seen) continue;
        const sr_col_view* pc = find_col(probe, slot);
        const OwnedCol* oc = nullptr;
        for (auto& c : j->build_cols)
            if (c.slot == slot) oc = &c;
        if (!pc && !oc) return fail(SR_ERR_INVALID_ARGUMENT, "other-join conjunct slot is in neither chunk");
        const int32_t type = pc ? pc->type : oc->type;
        const int w = type_width(type);
        data.emplace_back((size_t)std::max<int64_t>(nc, 1) * w);
        nulls.emplace_back((size_t)std::max<int64_t>(nc, 1), 0);
        for (int64_t c = 0; c < nc; c++) {
            if (pc) {
                memcpy(data.back().data() + c * w, (const uint8_t*)pc->data + (size_t)cpi[c] * w, w);
                nulls.back()[c] = pc->nulls ? pc->nulls[cpi[c]] : 0;
            } else {
                memcpy(data.back().data() + c * w, oc->data.data() + (size_t)cbi[c] * w, w);
                nulls.back()[c] = oc->nullable ? oc->nulls[cbi[c]] : 0;
            }
        }
        sr_col_view v;
        v.data = nullptr; // set below: the vectors may still move
        v.nulls = nullptr;
        v.type = type;
        v.slot_id = slot;
        pcols.push_back(v);
    }
    for (size_t k = 0; k < pcols.size(); k++) {
        pcols[k].data = data[k].data();
        pcols[k].nulls = nulls[k].data();
    }
    std::vector<uint8_t> pass((size_t)nc, 0);
    if (nc > 0) {
        sr_chunk_view pairs{pcols.data(), (int32_t)pcols.size(), SR_MEM_HOST, nc};
        ExprVal v;
        int32_t rc = eval_expr_range(&e, &pairs, 0, nc, &v);
        if (rc) return rc;
        if (v.is_double) return fail(SR_ERR_INVALID_ARGUMENT, "the other-join conjunct is not boolean");
        for (int64_t c = 0; c < nc; c++) pass[c] = !v.nul[c] && v.iv[c] != 0;
    }
    const int32_t jt = j->desc.join_type;
    const bool marks = jt == SR_JOIN_RIGHT_OUTER || jt == SR_JOIN_FULL_OUTER || jt == SR_JOIN_RIGHT_SEMI || jt == SR_JOIN_RIGHT_ANTI;
    const bool pairs_out = jt == SR_JOIN_INNER || jt == SR_JOIN_LEFT_OUTER || jt == SR_JOIN_RIGHT_OUTER || jt == SR_JOIN_FULL_OUTER;
    int64_t total = 0;
    auto emit = [&](uint32_t p, uint32_t b) {
        if (total < cap) {
            probe_index[total] = p;
            build_index[total] = b;
        }
        total++;
    };
    int64_t c = 0;
    for (int64_t i = 0; i < probe->num_rows; i++) {
        bool any = false;
        for (; c < nc && (int64_t)cpi[c] == i; c++) {
            if (!pass[c]) continue;
            any = true;
            if (marks) j->build_match[cbi[c]] = 1;
            if (pairs_out) emit((uint32_t)i, cbi[c]);
        }
        if ((jt == SR_JOIN_LEFT_OUTER || jt == SR_JOIN_FULL_OUTER || jt == SR_JOIN_LEFT_ANTI) && !any) emit((uint32_t)i, 0);
        if (jt == SR_JOIN_LEFT_SEMI && any) emit((uint32_t)i, 0);
    }
    return total > cap ? -total : total;
}

extern "C" int64_t orc_join_probe_all(orc_join* j, const sr_chunk_view* probe, uint32_t* probe_index,
                                      uint32_t* build_index, int64_t cap) {
    if (j->desc.other_conjunct.num_nodes > 0) return probe_all_with_conjunct(j, probe, probe_index, build_index, cap);
    std::vector<uint32_t> pi(j->chunk_size + 8), bi(j->chunk_size + 8);
    std::vector<sr_col_view> cols(probe->num_cols);
    int64_t total = 0;
    bool overflow = false;
    for (int64_t r0 = 0; r0 < probe->num_rows; r0 += j->chunk_size) {
        const int64_t n = std::min<int64_t>(j->chunk_size, probe->num_rows - r0);
        for (int k = 0; k < probe->num_cols; k++) {
            cols[k] = probe->cols[k];
            cols[k].data = (const uint8_t*)cols[k].data + r0 * type_width(cols[k].type);
            if (cols[k].nulls) cols[k].nulls += r0;
        }
        sr_chunk_view sub{cols.data(), probe->num_cols, SR_MEM_HOST, n};
        int32_t first = 1;
        while (true) {
            orc_probe_result res{};
            int32_t rc = orc_join_probe_chunk(j, &sub, first, pi.data(), bi.data(), &res);
            if (rc) return rc;
            for (int64_t q = 0; q < res.count; q++) {
                if (total + q < cap) {
                    probe_index[total + q] = (uint32_t)(pi[q] + r0);
                    build_index[total + q] = bi[q];
                } else {
                    overflow = true;
                }
            }
            total += res.count;
            if (!res.has_remain) break;
            first = 0;
        }
    }
    return overflow ? -total : total;
}

static int32_t join_output_slots(orc_join* j, const sr_chunk_view* probe, int64_t n, const uint32_t* probe_index,
                                 const uint32_t* build_index, const int32_t* probe_slots, int np,
                                 const int32_t* build_slots, int nb, void** out_data, uint8_t** out_nulls) {
    // _probe_output: Column::append_selective by probe_index (join_hash_map.hpp:163-181)
    for (int k = 0; k < np; k++) {
        const sr_col_view* c = find_col(probe, probe_slots[k]);
        if (!c) return fail(SR_ERR_INVALID_ARGUMENT, "probe out slot not in chunk");
        const int w = type_width(c->type);
        uint8_t* dst = (uint8_t*)out_data[k];
        for (int64_t q = 0; q < n; q++) memcpy(dst + q * w, (const uint8_t*)c->data + (int64_t)probe_index[q] * w, w);
        if (out_nulls && out_nulls[k]) {
            for (int64_t q = 0; q < n; q++) out_nulls[k][q] = c->nulls ? c->nulls[probe_index[q]] : 0;
        }
    }
    // _build_output: gather build columns by build_index; build_index 0 -> NULL for outer
    // joins (join_hash_map.hpp:252-269, 325-376)
    for (int k = 0; k < nb; k++) {
        const OwnedCol* oc = nullptr;
        for (auto& c : j->build_cols)
            if (c.slot == build_slots[k]) oc = &c;
        if (!oc) return fail(SR_ERR_INVALID_ARGUMENT, "build out slot not in build chunk");
        const int w = type_width(oc->type);
        uint8_t* dst = (uint8_t*)out_data[np + k];
        for (int64_t q = 0; q < n; q++) memcpy(dst + q * w, oc->data.data() + (size_t)build_index[q] * w, w);
        if (out_nulls && out_nulls[np + k]) {
            for (int64_t q = 0; q < n; q++)
                out_nulls[np + k][q] = build_index[q] == 0 ? 1 : (oc->nullable ? oc->nulls[build_index[q]] : 0);
        }
    }
    return SR_OK;
}

// POST_PROBE: _search_ht_remain (join_hash_map.hpp:420-457) -- build rows in order whose mark is 0 (RIGHT OUTER, FULL OUTER,
// RIGHT ANTI) or 1 (RIGHT SEMI, see probe_chunk_impl).  Returns the number of rows (negated when cap is too small).
extern "C" int64_t orc_join_probe_remain(orc_join* j, uint32_t* build_index, int64_t cap) {
    if (!j->built) return fail(SR_ERR_STATE, "probe_remain before build");
    const int32_t jt = j->desc.join_type;
    if (jt != SR_JOIN_RIGHT_OUTER && jt != SR_JOIN_FULL_OUTER && jt != SR_JOIN_RIGHT_SEMI && jt != SR_JOIN_RIGHT_ANTI)
        return fail(SR_ERR_INVALID_ARGUMENT, "join type has no post-probe phase");
    const uint8_t want = jt == SR_JOIN_RIGHT_SEMI ? 1 : 0;
    int64_t n = 0;
    for (uint32_t i = 1; i <= j->row_count; i++)
        if (j->build_match[i] == want) {
            if (n < cap) build_index[n] = i;
            n++;
        }
    return n > cap ? -n : n;
}

// output of the POST_PROBE rows: probe_out columns all NULL (_probe_null_output, join_hash_map.hpp:206-232; absent for RIGHT
// SEMI / ANTI), then the build_out columns gathered by build_index.  probe_types: sr_type of every probe_out slot.
extern "C" int32_t orc_join_output_remain(orc_join* j, int64_t n, const uint32_t* build_index, const int32_t* probe_types, void** out_data,
                                          uint8_t** out_nulls) {
    const int32_t jt = j->desc.join_type;
    const int np = (jt == SR_JOIN_RIGHT_OUTER || jt == SR_JOIN_FULL_OUTER) ? j->desc.num_probe_out : 0;
    for (int k = 0; k < np; k++) {
        memset(out_data[k], 0, (size_t)n * type_width(probe_types[k]));
        if (out_nulls && out_nulls[k]) memset(out_nulls[k], 1, (size_t)n);
    }
    sr_chunk_view none{nullptr, 0, SR_MEM_HOST, 0};
    return join_output_slots(j, &none, n, nullptr, build_index, nullptr, 0, j->desc.build_out_slots, j->desc.num_build_out, out_data + np,
                             out_nulls ? out_nulls + np : nullptr);
}

extern "C" int32_t orc_join_output(orc_join* j, const sr_chunk_view* probe, int64_t n, const uint32_t* probe_index,
                                   const uint32_t* build_index, void** out_data, uint8_t** out_nulls) {
    const int nb = (j->desc.join_type == SR_JOIN_LEFT_SEMI || j->desc.join_type == SR_JOIN_LEFT_ANTI)
                           ? 0
                           : j->desc.num_build_out;
    return join_output_slots(j, probe, n, probe_index, build_index, j->desc.probe_out_slots, j->desc.num_probe_out,
                             j->desc.build_out_slots, nb, out_data, out_nulls);
}

// ---------------------------------------------------------------------------------------
// hash aggregate
// ---------------------------------------------------------------------------------------
struct FnState {
    i128 isum = 0;   // SUM(int-class) / MIN / MAX (int) / decimal sums
    double dsum = 0; // SUM(double-class) / AVG / MIN / MAX (double)
    int64_t count = 0;
    bool has = false; // NullableAggregateFunctionState::is_null == false
};

struct Key128 {
    uint64_t lo = 0, hi = 0;
    uint32_t nul = 0; // bit k: group key k is NULL (kept apart so that 16 full key bytes fit)
    bool operator==(const Key128& o) const { return lo == o.lo && hi == o.hi && nul == o.nul; }
};

struct orc_agg {
    sr_agg_desc desc;
    int key_width[SR_MAX_GROUP_KEYS];
    int key_off_bits[SR_MAX_GROUP_KEYS];
    // open addressing map key -> group index (the AggHashMap, agg_hash_map.h:221-443)
    std::vector<Key128> slot_key;
    std::vector<int32_t> slot_group; // -1 empty
    uint64_t cap_mask = 0;
    // state arena in insertion order (aggregator.cpp:57-87,1718-1724)
    std::vector<Key128> group_keys;
    std::vector<FnState> states; // group * num_fns + f
    // COUNT(DISTINCT) states: DistinctAggregateState<LT>::set (distinct.h:51-62), one hash set per (group, function)
    std::vector<std::unordered_set<int64_t>> dsets; // group * num_fns + f; sized only when the desc has such a function
    bool has_distinct = false;
    int64_t num_groups = 0;
};

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

static int32_t agg_result_type(const sr_agg_fn& f) {
    switch (f.kind) {
    case SR_AGG_COUNT:
    case SR_AGG_COUNT_STAR:
    case SR_AGG_COUNT_DISTINCT: // TDistinctAggregateFunction<..., AggDistinctType::COUNT>::finalize_to_column (distinct.h:590-594)
        return SR_TYPE_BIGINT;
    case SR_AGG_AVG:
    case SR_AGG_AVG_MERGE:
        return SR_TYPE_DOUBLE; // AvgResultTrait<arithmetic> (avg.h:33-36)
    case SR_AGG_SUM:           // SumResultLT (sum.h:24-34)
        if (is_float_class(f.input_type)) return SR_TYPE_DOUBLE;
        if (is_decimal(f.input_type)) return SR_TYPE_DECIMAL128;
        if (f.input_type == SR_TYPE_LARGEINT) return SR_TYPE_LARGEINT;
        return SR_TYPE_BIGINT;
    default:
        return f.input_type; // MIN / MAX keep the input type (maxmin.h)
    }
}

extern "C" orc_agg* orc_agg_create(const sr_agg_desc* desc) {
    if (desc->num_group_keys < 0 || desc->num_group_keys > SR_MAX_GROUP_KEYS || desc->num_fns < 0 ||
        desc->num_fns > SR_MAX_AGG_FNS) {
        fail(SR_ERR_INVALID_ARGUMENT, "agg desc");
        return nullptr;
    }
    int bits = 0;
    auto* a = new orc_agg();
    a->desc = *desc;
    for (int k = 0; k < desc->num_group_keys; k++) {
        const int w = type_width(desc->group_types[k]);
        if (w == 0 || w > 8 || is_float_class(desc->group_types[k])) {
            delete a;
            fail(SR_ERR_NOT_SUPPORTED, "group key type");
            return nullptr;
        }
        a->key_width[k] = w;
        a->key_off_bits[k] = bits;
        bits += 8 * w;
    }
    if (bits > 128) {
        delete a;
        fail(SR_ERR_NOT_SUPPORTED, "group key wider than 16 bytes");
        return nullptr;
    }
    for (int f = 0; f < desc->num_fns; f++) {
        const sr_agg_fn& fn = desc->fns[f];
        if (fn.kind == SR_AGG_AVG && (is_decimal(fn.input_type) || type_width(fn.input_type) > 8)) {
            delete a;
            fail(SR_ERR_NOT_SUPPORTED, "avg on decimal/largeint");
            return nullptr;
        }
        if (fn.kind == SR_AGG_COUNT_DISTINCT) a->has_distinct = true;
    }
    a->slot_key.assign(1024, Key128());
    a->slot_group.assign(1024, -1);
    a->cap_mask = 1023;
    if (desc->num_group_keys == 0) { // single state (compute_single_agg_state, aggregator.cpp:882-905)
        a->num_groups = 1;
        a->group_keys.push_back(Key128());
        a->states.resize(std::max(1, desc->num_fns));
        if (a->has_distinct) a->dsets.resize(a->states.size());
    }
    return a;
}
extern "C" void orc_agg_destroy(orc_agg* a) {
    delete a;
}

static void agg_grow(orc_agg* a) {
    const size_t ncap = a->slot_key.size() * 2;
    std::vector<Key128> nk(ncap);
    std::vector<int32_t> ng(ncap, -1);
    const uint64_t mask = ncap - 1;
    for (int64_t g = 0; g < a->num_groups; g++) {
        uint64_t s = mix64(a->group_keys[g].lo ^ mix64(a->group_keys[g].hi + a->group_keys[g].nul)) & mask;
        while (ng[s] >= 0) s = (s + 1) & mask;
        ng[s] = (int32_t)g;
        nk[s] = a->group_keys[g];
    }
    a->slot_key.swap(nk);
    a->slot_group.swap(ng);
    a->cap_mask = mask;
}

static inline int32_t agg_find_or_insert(orc_agg* a, const Key128& k) {
    uint64_t s = mix64(k.lo ^ mix64(k.hi + k.nul)) & a->cap_mask;
    while (true) {
        const int32_t g = a->slot_group[s];
        if (g < 0) break;
        if (a->slot_key[s] == k) return g;
        s = (s + 1) & a->cap_mask;
    }
    const int32_t g = (int32_t)a->num_groups++;
    a->slot_group[s] = g;
    a->slot_key[s] = k;
    a->group_keys.push_back(k);
    a->states.resize((size_t)a->num_groups * std::max(1, a->desc.num_fns));
    if (a->has_distinct) a->dsets.resize(a->states.size());
    if ((uint64_t)a->num_groups * 2 > a->cap_mask) agg_grow(a);
    return g;
}

static inline void key_put(Key128* k, int off_bits, int w, uint64_t v) {
    const uint64_t mask = w == 8 ? ~0ull : ((1ull << (8 * w)) - 1);
    v &= mask;
    if (off_bits < 64) {
        k->lo |= v << off_bits;
        if (off_bits + 8 * w > 64) k->hi |= v >> (64 - off_bits);
    } else {
        k->hi |= v << (off_bits - 64);
    }
}
static inline uint64_t key_get(const Key128& k, int off_bits, int w) {
    const uint64_t mask = w == 8 ? ~0ull : ((1ull << (8 * w)) - 1);
    uint64_t v;
    if (off_bits < 64) {
        v = k.lo >> off_bits;
        if (off_bits + 8 * w > 64) v |= k.hi << (64 - off_bits);
    } else {
        v = k.hi >> (off_bits - 64);
    }
    return v & mask;
}

// null flags of group keys live in the top bits of `hi` (bit 127-k): a NULL key is its own
// group (agg_hash_map.h:363-395 keeps a dedicated null-key state).
static inline void key_set_null(Key128* k, int idx) {
    k->nul |= 1u << idx;
}
static inline bool key_is_null(const Key128& k, int idx) {
    return (k.nul >> idx) & 1;
}

static inline void fn_update_int(FnState& s, int32_t kind, int64_t v) {
    switch (kind) {
    case SR_AGG_SUM:
        s.isum += v; // SumAggregateFunction::update (sum.h:58-63)
        break;
    case SR_AGG_AVG:
        s.dsum += (double)v; // AvgAggregateState<double> (avg.h:62-66,84-103)
        s.count++;
        break;
    case SR_AGG_COUNT:
        s.count++;
        break;
    case SR_AGG_MIN:
        s.isum = s.has ? std::min<i128>(s.isum, v) : (i128)v;
        break;
    case SR_AGG_MAX:
        s.isum = s.has ? std::max<i128>(s.isum, v) : (i128)v;
        break;
    }
    s.has = true;
}
static inline void fn_update_double(FnState& s, int32_t kind, double v) {
    switch (kind) {
    case SR_AGG_SUM:
        s.dsum += v;
        break;
    case SR_AGG_AVG:
        s.dsum += v;
        s.count++;
        break;
    case SR_AGG_COUNT:
        s.count++;
        break;
    case SR_AGG_MIN:
        s.dsum = s.has ? std::min(s.dsum, v) : v;
        break;
    case SR_AGG_MAX:
        s.dsum = s.has ? std::max(s.dsum, v) : v;
        break;
    }
    s.has = true;
}

static int32_t agg_push_range(orc_agg* a, const sr_chunk_view* c, int64_t r0, int64_t n, std::vector<int32_t>& gidx) {
    const sr_agg_desc& d = a->desc;
    const int nf = std::max(1, d.num_fns);
    // evaluate_groupby_exprs + build_hash_map (aggregator.cpp:1346, 1616-1640)
    gidx.resize(n);
    if (d.num_group_keys == 0) {
        std::fill(gidx.begin(), gidx.end(), 0);
    } else {
        const sr_col_view* kc[SR_MAX_GROUP_KEYS];
        for (int k = 0; k < d.num_group_keys; k++) {
            kc[k] = find_col(c, d.group_slots[k]);
            if (!kc[k]) return fail(SR_ERR_INVALID_ARGUMENT, "group slot not in chunk");
        }
        for (int64_t i = 0; i < n; i++) {
            Key128 key;
            for (int k = 0; k < d.num_group_keys; k++) {
                if (kc[k]->nulls && kc[k]->nulls[r0 + i]) {
                    key_set_null(&key, k);
                } else {
                    key_put(&key, a->key_off_bits[k], a->key_width[k], (uint64_t)load_int(kc[k]->data, kc[k]->type, r0 + i));
                }
            }
            gidx[i] = agg_find_or_insert(a, key);
        }
    }
    // compute_batch_agg_states: per function, evaluate the input column then update row by
    // row through the state pointers (aggregator.cpp:907-929, aggregate.h:407-412)
    for (int f = 0; f < d.num_fns; f++) {
        const sr_agg_fn& fn = d.fns[f];
        if (fn.kind == SR_AGG_COUNT_STAR) {
            for (int64_t i = 0; i < n; i++) {
                FnState& s = a->states[(size_t)gidx[i] * nf + f];
                s.count++;
                s.has = true;
            }
            continue;
        }
        ExprVal v;
        int32_t rc = eval_expr_range(&fn.input, c, r0, n, &v);
        if (rc) return rc;
        if (fn.kind == SR_AGG_AVG_MERGE) { // AvgAggregateFunction::merge (avg.h:105-118): sum += sum state, count += count state
            const sr_col_view* cc = find_col(c, fn.reserved);
            if (!cc || !v.is_double) return fail(SR_ERR_INVALID_ARGUMENT, "AVG_MERGE needs a DOUBLE sum state and a count state column");
            for (int64_t i = 0; i < n; i++) {
                if (v.nul[i] || (cc->nulls && cc->nulls[r0 + i])) continue;
                const int64_t cnt = load_int(cc->data, cc->type, r0 + i);
                if (cnt == 0) continue;
                FnState& s = a->states[(size_t)gidx[i] * nf + f];
                s.dsum += v.dv[i];
                s.count += cnt;
                s.has = true;
            }
            continue;
        }
        if (fn.kind == SR_AGG_COUNT_DISTINCT) { // TDistinctAggregateFunction::update -> set.insert(key) (distinct.h:56,417-424)
            if (v.is_double) return fail(SR_ERR_NOT_SUPPORTED, "COUNT(DISTINCT) on a floating type");
            for (int64_t i = 0; i < n; i++) {
                if (v.nul[i]) continue;
                a->dsets[(size_t)gidx[i] * nf + f].insert(v.iv[i]);
                a->states[(size_t)gidx[i] * nf + f].has = true;
            }
            continue;
        }
        for (int64_t i = 0; i < n; i++) {
            if (v.nul[i]) continue; // NullableAggregateFunction skips NULL inputs
            FnState& s = a->states[(size_t)gidx[i] * nf + f];
            if (v.is_double)
                fn_update_double(s, fn.kind, v.dv[i]);
            else
                fn_update_int(s, fn.kind, v.iv[i]);
        }
    }
    return SR_OK;
}

// Aggregator::build_hash_map_with_selection (aggregator.cpp, agg_hash_map.h:303-361 compute_agg_states with
// allocate_and_compute_state = false): selection[i] = 1 when row i's group is NOT in the hash map (the row will be streamed),
// 0 when it is; the map is not changed.
extern "C" int32_t orc_agg_streaming_selection(const orc_agg* a, const sr_chunk_view* c, uint8_t* selection) {
    const sr_agg_desc& d = a->desc;
    const sr_col_view* kc[SR_MAX_GROUP_KEYS];
    for (int k = 0; k < d.num_group_keys; k++) {
        kc[k] = find_col(c, d.group_slots[k]);
        if (!kc[k]) return fail(SR_ERR_INVALID_ARGUMENT, "group slot not in chunk");
    }
    for (int64_t i = 0; i < c->num_rows; i++) {
        if (d.num_group_keys == 0) {
            selection[i] = 0;
            continue;
        }
        Key128 key;
        for (int k = 0; k < d.num_group_keys; k++) {
            if (kc[k]->nulls && kc[k]->nulls[i])
                key_set_null(&key, k);
            else
                key_put(&key, a->key_off_bits[k], a->key_width[k], (uint64_t)load_int(kc[k]->data, kc[k]->type, i));
        }
        uint64_t s = mix64(key.lo ^ mix64(key.hi + key.nul)) & a->cap_mask;
        bool found = false;
        while (a->slot_group[s] >= 0) {
            if (a->slot_key[s] == key) {
                found = true;
                break;
            }
            s = (s + 1) & a->cap_mask;
        }
        selection[i] = found ? 0 : 1;
    }
    return SR_OK;
}

extern "C" int32_t orc_agg_push(orc_agg* a, const sr_chunk_view* chunk) {
    std::vector<int32_t> gidx;
    for (int64_t r0 = 0; r0 < chunk->num_rows; r0 += ORC_CHUNK_SIZE) {
        const int64_t n = std::min<int64_t>(ORC_CHUNK_SIZE, chunk->num_rows - r0);
        int32_t rc = agg_push_range(a, chunk, r0, n, gidx);
        if (rc) return rc;
    }
    return SR_OK;
}
extern "C" int64_t orc_agg_num_groups(const orc_agg* a) {
    return a->num_groups;
}
extern "C" int32_t orc_agg_num_out_cols(const orc_agg* a) {
    return a->desc.num_group_keys + a->desc.num_fns;
}
extern "C" int32_t orc_agg_out_type(const orc_agg* a, int32_t k) {
    if (k < a->desc.num_group_keys) return a->desc.group_types[k];
    return agg_result_type(a->desc.fns[k - a->desc.num_group_keys]);
}

static void store_int(void* dst, int32_t type, int64_t row, i128 v) {
    switch (type_width(type)) {
    case 1:
        ((int8_t*)dst)[row] = (int8_t)v;
        break;
    case 2:
        ((int16_t*)dst)[row] = (int16_t)v;
        break;
    case 4:
        ((int32_t*)dst)[row] = (int32_t)v;
        break;
    case 8:
        ((int64_t*)dst)[row] = (int64_t)v;
        break;
    case 16:
        memcpy((uint8_t*)dst + row * 16, &v, 16);
        break;
    }
}

extern "C" int32_t orc_agg_output(orc_agg* a, void** out_data, uint8_t** out_nulls) {
    const sr_agg_desc& d = a->desc;
    const int nf = std::max(1, d.num_fns);
    for (int64_t g = 0; g < a->num_groups; g++) {
        for (int k = 0; k < d.num_group_keys; k++) {
            const bool nul = key_is_null(a->group_keys[g], k);
            uint64_t raw = key_get(a->group_keys[g], a->key_off_bits[k], a->key_width[k]);
            // sign extend
            const int w = a->key_width[k];
            int64_t v = w == 8 ? (int64_t)raw : (int64_t)(raw << (64 - 8 * w)) >> (64 - 8 * w);
            if (d.group_types[k] == SR_TYPE_BOOLEAN) v = (int64_t)raw;
            store_int(out_data[k], d.group_types[k], g, nul ? 0 : v);
            if (out_nulls && out_nulls[k]) out_nulls[k][g] = nul ? 1 : 0;
        }
        for (int f = 0; f < d.num_fns; f++) {
            const sr_agg_fn& fn = d.fns[f];
            const FnState& s = a->states[(size_t)g * nf + f];
            const int32_t rt = agg_result_type(fn);
            const int col = d.num_group_keys + f;
            bool nul = false;
            if (fn.kind == SR_AGG_COUNT || fn.kind == SR_AGG_COUNT_STAR) {
                ((int64_t*)out_data[col])[g] = s.count;
            } else if (fn.kind == SR_AGG_COUNT_DISTINCT) { // distinct_count() = set.size() (distinct.h:62,593)
                ((int64_t*)out_data[col])[g] = (int64_t)a->dsets[(size_t)g * nf + f].size();
            } else if (fn.kind == SR_AGG_AVG || fn.kind == SR_AGG_AVG_MERGE) {
                nul = !s.has || s.count == 0;
                ((double*)out_data[col])[g] = nul ? 0.0 : s.dsum / (double)s.count; // avg.h:218-236
            } else if (rt == SR_TYPE_DOUBLE) {
                nul = !s.has;
                ((double*)out_data[col])[g] = nul ? 0.0 : s.dsum;
            } else if (rt == SR_TYPE_FLOAT) {
                nul = !s.has;
                ((float*)out_data[col])[g] = nul ? 0.0f : (float)s.dsum;
            } else {
                nul = !s.has;
                store_int(out_data[col], rt, g, nul ? (i128)0 : s.isum);
            }
            if (out_nulls && out_nulls[col]) out_nulls[col][g] = nul ? 1 : 0;
        }
    }
    return SR_OK;
}

// Aggregator::output_chunk_by_streaming (aggregator.cpp:1071-1120) -> AggregateFunction::convert_to_serialize_format:
// every input row becomes one intermediate row.  SUM (sum.h:118-126) / MIN / MAX (maxmin.h) states are the input value in
// the result type (NULL stays NULL), COUNT (count.h:84-99) is 1 for a non-NULL input and 0 otherwise, COUNT(*) is 1.
// out_data / out_nulls: one buffer per function (chunk->num_rows elements of the result type); the group-by columns are
// the input columns themselves.
extern "C" int32_t orc_agg_convert_to_states(const sr_agg_desc* d, const sr_chunk_view* c, void** out_data, uint8_t** out_nulls) {
    const int64_t n = c->num_rows;
    for (int f = 0; f < d->num_fns; f++) {
        const sr_agg_fn& fn = d->fns[f];
        if (fn.kind == SR_AGG_AVG || fn.kind == SR_AGG_AVG_MERGE) return fail(SR_ERR_INVALID_ARGUMENT, "first-phase desc expected (AVG is split into SUM + COUNT)");
        const int32_t rt = agg_result_type(fn);
        if (type_width(rt) > 8) return fail(SR_ERR_NOT_SUPPORTED, "128-bit states in the intermediate format");
        if (fn.kind == SR_AGG_COUNT_STAR) {
            for (int64_t i = 0; i < n; i++) ((int64_t*)out_data[f])[i] = 1;
            continue;
        }
        for (int64_t r0 = 0; r0 < n; r0 += ORC_CHUNK_SIZE) {
            const int64_t m = std::min<int64_t>(ORC_CHUNK_SIZE, n - r0);
            ExprVal v;
            int32_t rc = eval_expr_range(&fn.input, c, r0, m, &v);
            if (rc) return rc;
            for (int64_t i = 0; i < m; i++) {
                if (fn.kind == SR_AGG_COUNT) {
                    ((int64_t*)out_data[f])[r0 + i] = v.nul[i] ? 0 : 1;
                    continue;
                }
                if (out_nulls && out_nulls[f]) out_nulls[f][r0 + i] = v.nul[i] ? 1 : 0;
                if (v.is_double) {
                    const double x = v.nul[i] ? 0.0 : v.dv[i];
                    if (rt == SR_TYPE_FLOAT)
                        ((float*)out_data[f])[r0 + i] = (float)x;
                    else
                        ((double*)out_data[f])[r0 + i] = x;
                } else {
                    store_int(out_data[f], rt, r0 + i, v.nul[i] ? (i128)0 : (i128)v.iv[i]);
                }
            }
        }
    }
    return SR_OK;
}

extern "C" int32_t orc_agg_merge(orc_agg* a, const orc_agg* o) {
    const sr_agg_desc& d = a->desc;
    const int nf = std::max(1, d.num_fns);
    for (int64_t g = 0; g < o->num_groups; g++) {
        const int32_t t = d.num_group_keys == 0 ? 0 : agg_find_or_insert(a, o->group_keys[g]);
        for (int f = 0; f < d.num_fns; f++) {
            const FnState& s = o->states[(size_t)g * nf + f];
            FnState& r = a->states[(size_t)t * nf + f];
            if (!s.has) continue;
            const sr_agg_fn& fn = d.fns[f];
            if (fn.kind == SR_AGG_COUNT_DISTINCT) { // merge = insert the other state's keys (distinct.h:64-75 deserialize_and_merge)
                const auto& os = o->dsets[(size_t)g * nf + f];
                a->dsets[(size_t)t * nf + f].insert(os.begin(), os.end());
                r.has = true;
                continue;
            }
            const bool dbl = is_float_class(fn.input_type);
            switch (fn.kind) {
            case SR_AGG_SUM:
                r.isum += s.isum;
                r.dsum += s.dsum;
                break;
            case SR_AGG_AVG:
            case SR_AGG_AVG_MERGE:
                r.dsum += s.dsum;
                r.count += s.count;
                break;
            case SR_AGG_COUNT:
            case SR_AGG_COUNT_STAR:
                r.count += s.count;
                break;
            case SR_AGG_MIN:
                if (dbl)
                    r.dsum = r.has ? std::min(r.dsum, s.dsum) : s.dsum;
                else
                    r.isum = r.has ? std::min(r.isum, s.isum) : s.isum;
                break;
            case SR_AGG_MAX:
                if (dbl)
                    r.dsum = r.has ? std::max(r.dsum, s.dsum) : s.dsum;
                else
                    r.isum = r.has ? std::max(r.isum, s.isum) : s.isum;
                break;
            }
            r.has = true;
        }
    }
    return SR_OK;
}

// ---------------------------------------------------------------------------------------
// exchange wire format: protobuf_serde.cpp:88-140 (encode_fixed32_le(buff, 1); encode_fixed32_le(buff + 4, num_rows);
// columns), column_array_serde.cpp:228-238 (write_little_endian_32(size) + write_raw), :768-772 (null column, data column)
// ---------------------------------------------------------------------------------------
extern "C" int64_t orc_chunk_serialize(const sr_chunk_view* c, int64_t r0, int64_t r1, uint8_t* dst, int64_t cap) {
    if (r0 < 0 || r1 < r0 || r1 > c->num_rows) return fail(SR_ERR_INVALID_ARGUMENT, "row range");
    const int64_t rows = r1 - r0;
    int64_t total = 8;
    for (int k = 0; k < c->num_cols; k++) total += (c->cols[k].nulls ? 4 + rows : 0) + 4 + rows * type_width(c->cols[k].type);
    if (!dst) return total;
    if (total > cap) return fail(SR_ERR_INVALID_ARGUMENT, "serialize: destination too small");
    uint8_t* p = dst;
    auto put32 = [&](uint32_t v) {
        for (int b = 0; b < 4; b++) *p++ = (uint8_t)(v >> (8 * b));
    };
    put32(1);
    put32((uint32_t)rows);
    for (int k = 0; k < c->num_cols; k++) {
        const sr_col_view& col = c->cols[k];
        const int w = type_width(col.type);
        if (col.nulls) {
            put32((uint32_t)rows);
            memcpy(p, col.nulls + r0, (size_t)rows);
            p += rows;
        }
        put32((uint32_t)(rows * w));
        memcpy(p, (const uint8_t*)col.data + r0 * w, (size_t)(rows * w));
        p += rows * w;
    }
    return total;
}

// ---------------------------------------------------------------------------------------
// XXH3 64-bit, short inputs (be/src/base/hash/xxhash.h: XXH3_len_1to3_64b, XXH3_len_4to8_64b, XXH3_len_9to16_64b,
// XXH3_rrmxmx, XXH3_avalanche, XXH64_avalanche, kSecret)
// ---------------------------------------------------------------------------------------
static const uint8_t kXxh3Secret[64] = {0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
                                         0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
                                         0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
                                         0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c};
static inline uint64_t xxh_r64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}
static inline uint32_t xxh_r32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}
static inline uint64_t xxh_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
extern "C" uint64_t orc_xxh3_64(const void* data, int32_t bytes, uint64_t seed) {
    const uint8_t* in = (const uint8_t*)data;
    const uint64_t len = (uint64_t)bytes;
    if (bytes >= 1 && bytes <= 3) {
        const uint32_t combined = ((uint32_t)in[0] << 16) | ((uint32_t)in[bytes >> 1] << 24) | (uint32_t)in[bytes - 1] | ((uint32_t)bytes << 8);
        const uint64_t bitflip = (uint64_t)(xxh_r32(kXxh3Secret) ^ xxh_r32(kXxh3Secret + 4)) + seed;
        uint64_t h = (uint64_t)combined ^ bitflip; // XXH64_avalanche
        h ^= h >> 33;
        h *= 0xC2B2AE3D27D4EB4FULL;
        h ^= h >> 29;
        h *= 0x165667B19E3779F9ULL;
        h ^= h >> 32;
        return h;
    }
    if (bytes >= 4 && bytes <= 8) {
        seed ^= (uint64_t)__builtin_bswap32((uint32_t)seed) << 32;
        const uint32_t in1 = xxh_r32(in), in2 = xxh_r32(in + bytes - 4);
        const uint64_t bitflip = (xxh_r64(kXxh3Secret + 8) ^ xxh_r64(kXxh3Secret + 16)) - seed;
        uint64_t h = ((uint64_t)in2 + ((uint64_t)in1 << 32)) ^ bitflip; // XXH3_rrmxmx
        h ^= xxh_rotl64(h, 49) ^ xxh_rotl64(h, 24);
        h *= 0x9FB21C651E98DF25ULL;
        h ^= (h >> 35) + len;
        h *= 0x9FB21C651E98DF25ULL;
        return h ^ (h >> 28);
    }
    if (bytes >= 9 && bytes <= 16) {
        const uint64_t bitflip1 = (xxh_r64(kXxh3Secret + 24) ^ xxh_r64(kXxh3Secret + 32)) + seed;
        const uint64_t bitflip2 = (xxh_r64(kXxh3Secret + 40) ^ xxh_r64(kXxh3Secret + 48)) - seed;
        const uint64_t lo = xxh_r64(in) ^ bitflip1, hi = xxh_r64(in + bytes - 8) ^ bitflip2;
        const unsigned __int128 prod = (unsigned __int128)lo * hi;
        uint64_t h = len + __builtin_bswap64(lo) + hi + ((uint64_t)prod ^ (uint64_t)(prod >> 64)); // XXH3_avalanche
        h ^= h >> 37;
        h *= 0x165667919E3779F9ULL;
        h ^= h >> 32;
        return h;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// exchange partitioning: exchange_sink_operator.cpp:586-637, shuffler.h:72-89,
// column_hash.cpp:138-300 (per-type byte feeding, NULL mixing)
// ---------------------------------------------------------------------------------------
extern "C" int32_t orc_hash_partition(const sr_part_desc* d, const sr_chunk_view* in, uint32_t* hash_values,
                                      uint32_t* channel_ids, uint32_t* row_indexes, int64_t* channel_starts) {
    const int64_t n = in->num_rows;
    // seeds: HashUtil::FNV_SEED / XXH3_SEED_32 (exchange_sink_operator.cpp:597-606), 0 for the CRC32 bucket shuffle
    const uint32_t seed = d->hash_fn == SR_HASH_FNV ? 0x811C9DC5u : d->hash_fn == SR_HASH_XXH3 ? 0x9E3779B1u : 0u;
    for (int64_t i = 0; i < n; i++) hash_values[i] = seed;
    for (int k = 0; k < d->num_part_slots; k++) {
        const sr_col_view* c = find_col(in, d->part_slots[k]);
        if (!c) return fail(SR_ERR_INVALID_ARGUMENT, "partition slot not in chunk");
        const int w = type_width(c->type);
        for (int64_t i = 0; i < n; i++) {
            uint32_t h = hash_values[i];
            if (c->nulls && c->nulls[i]) {
                if (d->hash_fn == SR_HASH_CRC32) {
                    const uint32_t zero = 0; // CRC32: NULL hashed as int 0 (column_hash.cpp:262-266)
                    h = orc_zlib_crc32(&zero, 4, h);
                } else {
                    h = h ^ (0x9e3779b9u + (h << 6) + (h >> 2)); // :270
                }
            } else {
                const uint8_t* p = (const uint8_t*)c->data + i * w;
                h = d->hash_fn == SR_HASH_FNV ? orc_fnv_hash(p, w, h) : d->hash_fn == SR_HASH_XXH3 ? (uint32_t)orc_xxh3_64(p, w, h) : orc_zlib_crc32(p, w, h);
            }
            hash_values[i] = h;
        }
    }
    const uint32_t nch = (uint32_t)d->num_channels;
    for (int64_t i = 0; i < n; i++)
        channel_ids[i] = d->reduce_op == SR_REDUCE_MULHI ? orc_reduce_op(hash_values[i], nch) : hash_values[i] % nch;
    // counting sort of row indexes per channel, stable (exchange_sink_operator.cpp:621-637)
    std::vector<int64_t> starts(nch + 1, 0);
    for (int64_t i = 0; i < n; i++) starts[channel_ids[i]]++;
    for (uint32_t c = 1; c <= nch; c++) starts[c] += starts[c - 1];
    for (int64_t i = n - 1; i >= 0; --i) {
        row_indexes[starts[channel_ids[i]] - 1] = (uint32_t)i;
        starts[channel_ids[i]]--;
    }
    // after the backward fill starts[c] is the begin offset of channel c
    for (uint32_t c = 0; c < nch; c++) channel_starts[c] = starts[c];
    channel_starts[nch] = n;
    return SR_OK;
}

// ---------------------------------------------------------------------------------------
// whole pipeline: one pipeline driver per thread, chunk at a time (pipeline_driver.cpp:270-500)
// ---------------------------------------------------------------------------------------
struct LocalCol {
    int32_t type = 0, slot = 0;
    bool has_nulls = false;
    std::vector<uint8_t> data, nulls;
};
struct LocalChunk {
    std::vector<LocalCol> cols;
    int64_t rows = 0;
    void view(std::vector<sr_col_view>& cv, sr_chunk_view* v) const {
        cv.resize(cols.size());
        for (size_t k = 0; k < cols.size(); k++)
            cv[k] = sr_col_view{cols[k].data.data(), cols[k].has_nulls ? cols[k].nulls.data() : nullptr, cols[k].type,
                                cols[k].slot};
        v->cols = cv.data();
        v->num_cols = (int32_t)cols.size();
        v->mem = SR_MEM_HOST;
        v->num_rows = rows;
    }
};

static void collect_expr_slots(const sr_expr& e, std::vector<int32_t>& s) {
    for (int k = 0; k < e.num_nodes; k++)
        if (e.nodes[k].op == SR_EX_COL) s.push_back(e.nodes[k].slot_id);
}

struct FragPlan {
    // slots of the incoming chunk that stage k must forward (needed by later stages)
    std::vector<std::vector<int32_t>> forward; // [num_joins + 1]; forward[0] = scan output
};

static FragPlan make_plan(const orc_fragment_desc* d) {
    FragPlan p;
    std::vector<int32_t> agg_slots;
    for (int k = 0; k < d->agg.num_group_keys; k++) agg_slots.push_back(d->agg.group_slots[k]);
    for (int f = 0; f < d->agg.num_fns; f++)
        if (d->agg.fns[f].kind != SR_AGG_COUNT_STAR) collect_expr_slots(d->agg.fns[f].input, agg_slots);
    p.forward.resize(d->num_joins + 1);
    // needed after stage k (k = 0: scan): join keys of joins >= k and agg slots, minus payloads
    // produced by joins >= k
    for (int k = 0; k <= d->num_joins; k++) {
        std::vector<int32_t> need = agg_slots;
        for (int q = k; q < d->num_joins; q++) need.push_back(d->joins[q].probe_key_slot);
        std::sort(need.begin(), need.end());
        need.erase(std::unique(need.begin(), need.end()), need.end());
        for (int q = k; q < d->num_joins; q++)
            for (int pl = 0; pl < d->joins[q].num_payload; pl++)
                need.erase(std::remove(need.begin(), need.end(), d->joins[q].payload_build_slots[pl]), need.end());
        p.forward[k] = need;
    }
    return p;
}

struct Driver {
    const orc_fragment_desc* d;
    const FragPlan* plan;
    std::vector<ProbeState> probers; // clone_readable_table: shared items, own probe state
    orc_agg* agg;
    int64_t rows_passed = 0;
    std::vector<uint32_t> pi, bi;
    std::vector<int32_t> gidx;
    std::vector<uint8_t> sel;
    int32_t err = 0;
    std::string errmsg;
};

static int32_t run_stage(Driver& dr, int stage, const sr_chunk_view* in);

static int32_t run_join_stage(Driver& dr, int stage, const sr_chunk_view* in) {
    const orc_fragment_desc* d = dr.d;
    const orc_frag_join& fj = d->joins[stage];
    orc_join* j = fj.join;
    ProbeState& ps = dr.probers[stage];
    const int32_t probe_slots[SR_MAX_JOIN_KEYS] = {fj.probe_key_slot, 0};
    const std::vector<int32_t>& fwd = dr.plan->forward[stage + 1];
    int32_t first = 1;
    while (true) {
        orc_probe_result res{};
        int32_t rc = probe_chunk_impl(j, ps, probe_slots, in, first, dr.pi.data(), dr.bi.data(), &res);
        if (rc) return rc;
        if (res.count > 0) {
            LocalChunk out;
            out.rows = res.count;
            std::vector<void*> od;
            std::vector<uint8_t*> on;
            std::vector<int32_t> pslots;
            for (int32_t s : fwd) {
                const sr_col_view* c = find_col(in, s);
                if (!c) continue; // produced later or by this join
                LocalCol lc;
                lc.type = c->type;
                lc.slot = s;
                lc.has_nulls = c->nulls != nullptr;
                lc.data.resize(res.count * type_width(c->type));
                if (lc.has_nulls) lc.nulls.resize(res.count);
                out.cols.push_back(std::move(lc));
                pslots.push_back(s);
            }
            const bool outer = j->desc.join_type == SR_JOIN_LEFT_OUTER;
            std::vector<int32_t> bslots;
            for (int pl = 0; pl < fj.num_payload; pl++) {
                const OwnedCol* oc = nullptr;
                for (auto& c : j->build_cols)
                    if (c.slot == fj.payload_build_slots[pl]) oc = &c;
                if (!oc) return fail(SR_ERR_INVALID_ARGUMENT, "payload slot not in build chunk");
                LocalCol lc;
                lc.type = oc->type;
                lc.slot = oc->slot;
                lc.has_nulls = outer || oc->nullable;
                lc.data.resize(res.count * type_width(oc->type));
                if (lc.has_nulls) lc.nulls.resize(res.count);
                out.cols.push_back(std::move(lc));
                bslots.push_back(oc->slot);
            }
            for (auto& lc : out.cols) {
                od.push_back(lc.data.data());
                on.push_back(lc.has_nulls ? lc.nulls.data() : nullptr);
            }
            rc = join_output_slots(j, in, res.count, dr.pi.data(), dr.bi.data(), pslots.data(), (int)pslots.size(),
                                   bslots.data(), (int)bslots.size(), od.data(), on.data());
            if (rc) return rc;
            std::vector<sr_col_view> cv;
            sr_chunk_view v;
            out.view(cv, &v);
            rc = run_stage(dr, stage + 1, &v);
            if (rc) return rc;
        }
        if (!res.has_remain) break;
        first = 0;
    }
    return SR_OK;
}

static int32_t run_stage(Driver& dr, int stage, const sr_chunk_view* in) {
    if (in->num_rows == 0) return SR_OK;
    if (stage < dr.d->num_joins) return run_join_stage(dr, stage, in);
    dr.rows_passed += in->num_rows;
    return agg_push_range(dr.agg, in, 0, in->num_rows, dr.gidx);
}

static void driver_loop(Driver& dr, const sr_chunk_view* fact, std::atomic<int64_t>* next_morsel) {
    const orc_fragment_desc* d = dr.d;
    const int64_t MORSEL = 64 * ORC_CHUNK_SIZE; // 64 chunks per IO task (scan_operator.h:119)
    dr.sel.resize(ORC_CHUNK_SIZE);
    const bool has_filter = d->scan.num_preds > 0 || d->scan.num_filter_exprs > 0;
    const std::vector<int32_t>& fwd = dr.plan->forward[0];
    while (true) {
        const int64_t m0 = next_morsel->fetch_add(MORSEL);
        if (m0 >= fact->num_rows) break;
        const int64_t m1 = std::min(fact->num_rows, m0 + MORSEL);
        for (int64_t r0 = m0; r0 < m1; r0 += ORC_CHUNK_SIZE) {
            const int64_t n = std::min<int64_t>(ORC_CHUNK_SIZE, m1 - r0);
            int32_t rc;
            if (!has_filter) {
                // no conjuncts: the scanned chunk is forwarded as is (zero copy view)
                std::vector<sr_col_view> cv;
                for (int32_t s : fwd) {
                    const sr_col_view* c = find_col(fact, s);
                    if (!c) {
                        dr.err = fail(SR_ERR_INVALID_ARGUMENT, "fact slot missing");
                        dr.errmsg = g_err;
                        return;
                    }
                    sr_col_view v = *c;
                    v.data = (const uint8_t*)c->data + r0 * type_width(c->type);
                    if (v.nulls) v.nulls += r0;
                    cv.push_back(v);
                }
                sr_chunk_view v{cv.data(), (int32_t)cv.size(), SR_MEM_HOST, n};
                rc = run_stage(dr, 0, &v);
            } else {
                rc = scan_select_range(&d->scan, fact, r0, n, dr.sel.data());
                if (!rc) {
                    LocalChunk out;
                    int64_t kept = 0;
                    for (int64_t i = 0; i < n; i++) kept += dr.sel[i];
                    out.rows = kept;
                    if (kept > 0) {
                        for (int32_t s : fwd) {
                            const sr_col_view* c = find_col(fact, s);
                            if (!c) {
                                rc = fail(SR_ERR_INVALID_ARGUMENT, "fact slot missing");
                                break;
                            }
                            const int w = type_width(c->type);
                            LocalCol lc;
                            lc.type = c->type;
                            lc.slot = s;
                            lc.data.resize(n * w);
                            memcpy(lc.data.data(), (const uint8_t*)c->data + r0 * w, n * w);
                            orc_filter_range(dr.sel.data(), lc.data.data(), w, 0, n);
                            if (c->nulls) {
                                lc.has_nulls = true;
                                lc.nulls.assign(c->nulls + r0, c->nulls + r0 + n);
                                orc_filter_range(dr.sel.data(), lc.nulls.data(), 1, 0, n);
                            }
                            out.cols.push_back(std::move(lc));
                        }
                        if (!rc) {
                            std::vector<sr_col_view> cv;
                            sr_chunk_view v;
                            out.view(cv, &v);
                            rc = run_stage(dr, 0, &v);
                        }
                    }
                }
            }
            if (rc) {
                dr.err = rc;
                dr.errmsg = g_err;
                return;
            }
        }
    }
}

extern "C" int32_t orc_fragment_run(const orc_fragment_desc* desc, const sr_chunk_view* fact, int32_t num_threads,
                                    orc_agg* result, int64_t* rows_passed) {
    if (num_threads < 1) num_threads = 1;
    for (int k = 0; k < desc->num_joins; k++)
        if (!desc->joins[k].join || !desc->joins[k].join->built) return fail(SR_ERR_STATE, "join not built");
    FragPlan plan = make_plan(desc);
    std::vector<Driver> drivers(num_threads);
    std::vector<std::unique_ptr<orc_agg>> aggs;
    for (int t = 0; t < num_threads; t++) {
        Driver& dr = drivers[t];
        dr.d = desc;
        dr.plan = &plan;
        for (int k = 0; k < desc->num_joins; k++) {
            // HashJoinProbeOperator::_reference_builder_hash_table_once -> clone_readable_table
            // (hash_join_probe_operator.cpp:104-117): every prober gets its own probe state.
            dr.probers.emplace_back();
            dr.probers.back().prepare(desc->joins[k].join->chunk_size);
        }
        aggs.emplace_back(orc_agg_create(&desc->agg));
        if (!aggs.back()) return SR_ERR_INVALID_ARGUMENT;
        dr.agg = aggs.back().get();
        dr.pi.resize(ORC_CHUNK_SIZE + 8);
        dr.bi.resize(ORC_CHUNK_SIZE + 8);
    }
    std::atomic<int64_t> next_morsel{0};
    std::vector<std::thread> th;
    for (int t = 1; t < num_threads; t++) th.emplace_back([&, t]() { driver_loop(drivers[t], fact, &next_morsel); });
    driver_loop(drivers[0], fact, &next_morsel);
    for (auto& t : th) t.join();
    int64_t passed = 0;
    for (auto& dr : drivers) {
        if (dr.err) return fail(dr.err, dr.errmsg);
        passed += dr.rows_passed;
    }
    for (int t = 0; t < num_threads; t++) {
        int32_t rc = orc_agg_merge(result, drivers[t].agg);
        if (rc) return rc;
    }
    if (rows_passed) *rows_passed = passed;
    return SR_OK;
}

// ---------------------------------------------------------------------------------------
// runtime filter: min/max + has_null (MinMaxRuntimeFilter, runtime_filter.h:584-) and the split-block bloom filter
// (SimdBlockFilter, runtime_filter.h:79-240; init runtime_filter.cpp:26-35; make_mask :114-123; SALT :58-59)
// ---------------------------------------------------------------------------------------
static const uint32_t RF_SALT[8] = {0x47b6137b, 0x44974d91, 0x8824ad5b, 0xa2b7289d, 0x705495c7, 0x2df1424b, 0x9efc4947, 0x5c6bfb31};

struct orc_rf {
    int32_t key_type = 0;
    int32_t log_num_buckets = 0; // 0: no bloom part
    uint64_t directory_mask = 0;
    std::vector<uint32_t> directory; // 8 words per bucket
    int64_t min_value = INT64_MAX, max_value = INT64_MIN;
    int64_t num_inserted = 0;
    bool has_null = false;
    // runtime IN filter (HashJoiner::_create_runtime_in_filters, hash_joiner.cpp:563-609): build sides of at most
    // max_pushdown_conditions_per_column = 1024 rows publish their distinct keys as an IN predicate -- exact membership
    bool in_enabled = false;
    int64_t in_rows = 0;
    std::set<int64_t> in_values;
};

extern "C" orc_rf* orc_rf_create(int32_t key_type, int64_t expected_rows, int32_t with_bloom) {
    if (type_width(key_type) == 0 || type_width(key_type) > 8 || is_float_class(key_type)) {
        fail(SR_ERR_NOT_SUPPORTED, "runtime filter key type");
        return nullptr;
    }
    auto* rf = new orc_rf();
    rf->key_type = key_type;
    rf->in_enabled = with_bloom && expected_rows <= SR_RF_IN_FILTER_ROW_LIMIT; // a pure MinMaxRuntimeFilter (with_bloom = 0) stays a range test
    if (with_bloom) { // SimdBlockFilter::init
        const uint64_t nums = (uint64_t)std::max<int64_t>(1, expected_rows);
        const int log_heap_space = (int)std::ceil(std::log2((double)nums));
        rf->log_num_buckets = std::max(1, log_heap_space - 5);
        rf->directory_mask = (1ull << std::min(63, rf->log_num_buckets)) - 1;
        rf->directory.assign((size_t)8 << rf->log_num_buckets, 0u);
    }
    return rf;
}
extern "C" void orc_rf_destroy(orc_rf* rf) {
    delete rf;
}
static inline void rf_make_mask(uint32_t key, uint32_t* masks) {
    for (int i = 0; i < 8; i++) masks[i] = 1u << ((key * RF_SALT[i]) >> 27);
}
extern "C" void orc_rf_insert_hash(orc_rf* rf, uint64_t hash) {
    if (rf->log_num_buckets == 0) return;
    uint32_t masks[8];
    rf_make_mask((uint32_t)(hash >> rf->log_num_buckets), masks);
    uint32_t* b = rf->directory.data() + 8 * (hash & rf->directory_mask);
    for (int i = 0; i < 8; i++) b[i] |= masks[i];
}
extern "C" int32_t orc_rf_test_hash(const orc_rf* rf, uint64_t hash) {
    if (rf->log_num_buckets == 0) return 1;
    uint32_t masks[8];
    rf_make_mask((uint32_t)(hash >> rf->log_num_buckets), masks);
    const uint32_t* b = rf->directory.data() + 8 * (hash & rf->directory_mask);
    for (int i = 0; i < 8; i++)
        if ((b[i] & masks[i]) == 0) return 0;
    return 1;
}
// phmap_mix<8> (base/phmap/phmap_utils.h:86-95) over std::hash<integer> (identity on the value converted to size_t)
extern "C" uint64_t orc_rf_value_hash(int64_t value) {
    const unsigned __int128 p = (unsigned __int128)(uint64_t)value * 0xde5fb9d2630458e9ull;
    return (uint64_t)(p >> 64) + (uint64_t)p;
}
extern "C" int32_t orc_rf_insert(orc_rf* rf, const sr_chunk_view* in, int32_t slot_id, int32_t insert_nulls) {
    const sr_col_view* c = find_col(in, slot_id);
    if (!c) return fail(SR_ERR_INVALID_ARGUMENT, "runtime filter slot not in chunk");
    if (type_width(c->type) != type_width(rf->key_type) || is_float_class(c->type)) return fail(SR_ERR_INVALID_ARGUMENT, "runtime filter key type differs");
    if (rf->in_enabled && in->num_rows > 0) {
        rf->in_rows += in->num_rows;
        if (rf->in_rows > SR_RF_IN_FILTER_ROW_LIMIT) {
            rf->in_enabled = false;
            rf->in_values.clear();
        }
    }
    for (int64_t i = 0; i < in->num_rows; i++) {
        if (c->nulls && c->nulls[i]) {
            if (insert_nulls) rf->has_null = true;
            continue;
        }
        const int64_t v = load_int(c->data, c->type, i);
        if (rf->in_enabled) rf->in_values.insert(v);
        rf->min_value = std::min(rf->min_value, v);
        rf->max_value = std::max(rf->max_value, v);
        rf->num_inserted++;
        orc_rf_insert_hash(rf, orc_rf_value_hash(v));
    }
    return SR_OK;
}
extern "C" int32_t orc_rf_merge(orc_rf* rf, const orc_rf* other) {
    if (rf->log_num_buckets != other->log_num_buckets) return fail(SR_ERR_INVALID_ARGUMENT, "bloom directories of different size");
    for (size_t i = 0; i < rf->directory.size(); i++) rf->directory[i] |= other->directory[i];
    rf->min_value = std::min(rf->min_value, other->min_value);
    rf->max_value = std::max(rf->max_value, other->max_value);
    rf->num_inserted += other->num_inserted;
    rf->has_null = rf->has_null || other->has_null;
    if (rf->in_enabled) { // the total filter keeps an IN part only when every partial one has it and the union stays small
        if (!other->in_enabled) {
            rf->in_enabled = false;
        } else {
            rf->in_values.insert(other->in_values.begin(), other->in_values.end());
            if ((int64_t)rf->in_values.size() > SR_RF_IN_FILTER_ROW_LIMIT) rf->in_enabled = false;
        }
        if (!rf->in_enabled) rf->in_values.clear();
    }
    return SR_OK;
}
extern "C" int32_t orc_rf_evaluate(const orc_rf* rf, const sr_chunk_view* in, int32_t slot_id, uint8_t* selection, int32_t merge_and) {
    const sr_col_view* c = find_col(in, slot_id);
    if (!c) return fail(SR_ERR_INVALID_ARGUMENT, "runtime filter slot not in chunk");
    for (int64_t i = 0; i < in->num_rows; i++) {
        uint8_t pass;
        if (c->nulls && c->nulls[i]) {
            pass = rf->has_null ? 1 : 0;
        } else {
            const int64_t v = load_int(c->data, c->type, i);
            pass = v >= rf->min_value && v <= rf->max_value &&
                   (rf->in_enabled && !rf->in_values.empty() ? rf->in_values.count(v) != 0 : orc_rf_test_hash(rf, orc_rf_value_hash(v)) != 0);
        }
        selection[i] = merge_and ? (uint8_t)(selection[i] && pass) : pass;
    }
    return SR_OK;
}
extern "C" int32_t orc_rf_get_info(const orc_rf* rf, sr_rf_info* info) {
    info->min_value = rf->min_value;
    info->max_value = rf->max_value;
    info->num_inserted = rf->num_inserted;
    info->has_null = rf->has_null ? 1 : 0;
    info->log_num_buckets = rf->log_num_buckets;
    info->key_type = rf->key_type;
    info->num_in_values = rf->in_enabled ? (int32_t)rf->in_values.size() : -1;
    return SR_OK;
}
extern "C" const void* orc_rf_directory(const orc_rf* rf, int64_t* bytes) {
    *bytes = (int64_t)rf->directory.size() * 4;
    return rf->directory.data();
}

// ---------------------------------------------------------------------------------------
// segment pages: frame of reference + plain
// ---------------------------------------------------------------------------------------
namespace {
constexpr int FOR_FRAME = 128; // ForEncoder::FRAME_VALUE_NUM (frame_of_reference_coding.h:135)

template <typename T>
int for_bits(T v) { // bits() / bits_less_than_64 (frame_of_reference_coding.h:49-70): sign-extended to 64 bits first
    const uint64_t u = (uint64_t)(int64_t)v;
    return u == 0 ? 0 : 64 - __builtin_clzll(u);
}

// bit_pack (frame_of_reference_coding.cpp:96-118): value i occupies bits [i*bw, (i+1)*bw) of the stream, most significant
// bit first inside the value and inside every byte
template <typename T>
void for_bit_pack(const T* in, int n, int bw, uint8_t* out) {
    typedef typename std::make_unsigned<T>::type U;
    size_t bit = 0;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < bw; k++, bit++) {
            if ((bit & 7) == 0) out[bit >> 3] = 0;
            if (((U)in[i] >> (bw - 1 - k)) & 1) out[bit >> 3] |= (uint8_t)(0x80u >> (bit & 7));
        }
}

template <typename T>
void for_put_le(std::vector<uint8_t>& b, T v) {
    for (size_t i = 0; i < sizeof(T); i++) b.push_back((uint8_t)((typename std::make_unsigned<T>::type)v >> (8 * i)));
}

// bit_packing_one_frame_value (frame_of_reference_coding.cpp:120-211)
template <typename T>
void for_encode_frame(const T* in, int n, std::vector<uint8_t>& buf, std::vector<uint8_t>& formats, std::vector<uint8_t>& widths) {
    T mn = in[0], mx = in[0];
    bool ascending = true, keep = false;
    int bw = 0;
    const T half_max = std::numeric_limits<T>::max() >> 1;
    for (int i = 1; i < n; i++) {
        if (ascending) {
            if (in[i] < in[i - 1]) {
                ascending = false;
            } else if ((T)((in[i] >> 1) - (in[i - 1] >> 1)) > half_max) {
                keep = true;
            } else {
                bw = std::max(bw, for_bits((T)(in[i] - in[i - 1])));
            }
        }
        if (in[i] < mn) {
            mn = in[i];
            continue;
        }
        if (in[i] > mx) mx = in[i];
    }
    if (!ascending && (T)((mx >> 1) - (mn >> 1)) > half_max) keep = true;
    for_put_le(buf, mn);
    T delta[FOR_FRAME];
    const T* packed = delta;
    if (keep) {
        bw = 8 * (int)sizeof(T);
        packed = in;
    } else if (ascending) {
        delta[0] = 0;
        for (int i = 1; i < n; i++) delta[i] = (T)(in[i] - in[i - 1]);
    } else {
        bw = for_bits((T)(mx - mn));
        for (int i = 0; i < n; i++) delta[i] = (T)(in[i] - mn);
    }
    // (the keep-original branch sizes its buffer in bits, frame_of_reference_coding.cpp:172-176: n * bit_width BYTES are
    // appended, of which bit_pack fills the first n * bit_width / 8)
    const size_t len = keep ? (size_t)n * bw : ((size_t)n * bw + 7) / 8;
    const size_t at = buf.size();
    buf.resize(at + len, 0);
    if (bw > 0) for_bit_pack(packed, n, bw, buf.data() + at);
    formats.push_back(keep ? 2 : (ascending ? 1 : 0));
    widths.push_back((uint8_t)bw);
}

template <typename T>
int64_t for_encode(const T* v, int64_t n, uint8_t* out, int64_t cap) {
    std::vector<uint8_t> buf, formats, widths;
    for (int64_t i = 0; i < n; i += FOR_FRAME) for_encode_frame(v + i, (int)std::min<int64_t>(FOR_FRAME, n - i), buf, formats, widths);
    for (size_t f = 0; f < formats.size(); f++) { // flush (frame_of_reference_coding.cpp:213-233)
        buf.push_back(formats[f]);
        buf.push_back(widths[f]);
    }
    buf.push_back((uint8_t)FOR_FRAME);
    for_put_le(buf, (uint32_t)n);
    if ((int64_t)buf.size() > cap) return -(int64_t)buf.size();
    memcpy(out, buf.data(), buf.size());
    return (int64_t)buf.size();
}

// ForDecoder::init + decode_current_frame (frame_of_reference_coding.cpp:246-352)
template <typename T>
int64_t for_decode(const uint8_t* p, int64_t len, T* out, int64_t cap) {
    typedef typename std::make_unsigned<T>::type U;
    if (len < 5) return -1;
    const int frame = p[len - 5];
    uint32_t n;
    memcpy(&n, p + len - 4, 4);
    if (frame == 0) return n == 0 ? 0 : -1;
    const int64_t frames = (n + frame - 1) / frame;
    const int64_t meta = len - 5 - frames * 2;
    if (meta < 0) return -1;
    if ((int64_t)n > cap) return -(int64_t)n;
    int64_t off = 0;
    for (int64_t f = 0; f < frames; f++) {
        const int fmt = p[meta + 2 * f], bw = p[meta + 2 * f + 1];
        const int cnt = (int)std::min<int64_t>(frame, (int64_t)n - f * frame);
        if (off + (int64_t)sizeof(T) > meta) return -1;
        U mn = 0;
        for (size_t i = 0; i < sizeof(T); i++) mn |= (U)p[off + i] << (8 * i);
        const uint8_t* bits = p + off + sizeof(T);
        U prev = mn;
        for (int i = 0; i < cnt; i++) { // bit_unpack (frame_of_reference_coding.cpp:283-300)
            U v = 0;
            const size_t b0 = (size_t)i * bw;
            for (int k = 0; k < bw; k++) {
                const size_t bit = b0 + k;
                v = (U)(v << 1) | ((bits[bit >> 3] >> (7 - (bit & 7))) & 1u);
            }
            if (fmt == 2) {
                out[f * frame + i] = (T)v;
            } else if (fmt == 1) {
                prev = (U)(prev + v);
                out[f * frame + i] = (T)prev;
            } else {
                out[f * frame + i] = (T)(U)(v + mn);
            }
        }
        off += (int64_t)bw * frame / 8 + (int64_t)sizeof(T); // ForDecoder::init's frame offsets: full frames
    }
    return (int64_t)n;
}
} // namespace

extern "C" int64_t orc_for_encode(int32_t elem_size, const void* values, int64_t n, uint8_t* out, int64_t cap) {
    if (elem_size == 4) return for_encode((const int32_t*)values, n, out, cap);
    if (elem_size == 8) return for_encode((const int64_t*)values, n, out, cap);
    return fail(SR_ERR_NOT_SUPPORTED, "frame-of-reference element size");
}
extern "C" int64_t orc_for_decode(int32_t elem_size, const uint8_t* page, int64_t len, void* out, int64_t cap) {
    if (elem_size == 4) return for_decode(page, len, (int32_t*)out, cap);
    if (elem_size == 8) return for_decode(page, len, (int64_t*)out, cap);
    return fail(SR_ERR_NOT_SUPPORTED, "frame-of-reference element size");
}
// PlainPageBuilder::finish / PlainPageDecoder::init (plain_page.h:82-98,146-165): uint32 count, then the values
extern "C" int64_t orc_plain_encode(int32_t elem_size, const void* values, int64_t n, uint8_t* out, int64_t cap) {
    const int64_t need = 4 + n * elem_size;
    if (need > cap) return -need;
    const uint32_t c = (uint32_t)n;
    memcpy(out, &c, 4);
    memcpy(out + 4, values, (size_t)(n * elem_size));
    return need;
}
extern "C" int64_t orc_plain_decode(int32_t elem_size, const uint8_t* page, int64_t len, void* out, int64_t cap) {
    if (len < 4) return -1;
    uint32_t c;
    memcpy(&c, page, 4);
    if (len != 4 + (int64_t)c * elem_size) return -1;
    if ((int64_t)c > cap) return -(int64_t)c;
    memcpy(out, page + 4, (size_t)c * elem_size);
    return (int64_t)c;
}
