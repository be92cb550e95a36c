/*
 * sr_oracle.h -- CPU restatement of the StarRocks BE hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libsr_gpu.so) never links, loads or calls it.
 *
 * Why a restatement: the reference BE cannot be compiled in this image (it needs ~50 third
 * party libraries, generated thrift/protobuf sources, glog/gflags/fmt/brpc headers; see
 * DESIGN.md "Oracle").  Every function below cites the reference file:line it follows, and the
 * restatement is pinned by the reference's own known-answer tests (tests/test_oracle_*.py).
 *
 * Descriptor structs (sr_pred, sr_expr, sr_chunk_view, sr_join_desc, sr_agg_desc ...) are the
 * ones declared in include/sr_gpu_ops.h so the same inputs drive both implementations.
 */
#ifndef SR_ORACLE_H
#define SR_ORACLE_H

#include "../include/sr_gpu_ops.h"

#ifdef __cplusplus
extern "C" {
#endif

/* vector_chunk_size (be/src/common/config.h:915) */
#define ORC_CHUNK_SIZE 4096

/* ---- hash functions ---------------------------------------------------------------- */
/* JoinKeyHash<T,4> / <T,8>  (be/src/exec/join/join_hash_map_helper.h:35-54) */
uint32_t orc_join_key_hash32(uint32_t v, uint32_t num_log_buckets);
uint32_t orc_join_key_hash64(uint64_t v, uint32_t num_log_buckets);
/* JoinKeyHash<Slice>: crc_hash_32 with seed 0x811C9DC5, & (num_buckets-1) (:24-31,56-63) */
uint32_t orc_join_key_hash_slice(const void* data, int32_t size, uint32_t num_buckets);
/* JoinHashMapHelper::calc_bucket_size (:70-77) */
uint32_t orc_calc_bucket_size(uint32_t size);
/* raw CRC32C as computed by SSE4.2 _mm_crc32_u32/u8 (no pre/post inversion) */
uint32_t orc_crc32c(const void* data, int32_t bytes, uint32_t seed);
/* crc_hash_32 proper: CRC32C followed by phmap_mix<4> (hash.h:25-33,126-129) */
uint32_t orc_crc_hash_32(const void* data, int32_t bytes, uint32_t seed);
/* HashUtil::zlib_crc_hash = zlib crc32 */
uint32_t orc_zlib_crc32(const void* data, int32_t bytes, uint32_t seed);
/* HashUtil::fnv_hash (be/src/base/hash/hash_util.hpp:127-134) */
uint32_t orc_fnv_hash(const void* data, int32_t bytes, uint32_t seed);
/* HashUtil::xx_hash3_64 = XXH3_64bits_withSeed of the vendored xxHash (be/src/base/hash/hash_util.cpp:100-102,
 * be/src/base/hash/xxhash.h XXH3_len_1to3 / 4to8 / 9to16_64b) for inputs of 1..16 bytes -- every fixed-width value.
 * Pinned by HashFunctionsTest.test_xx_hash3_64 (be/test/exprs/hash_functions_test.cpp:85-118) and, where the reference tree
 * is present, cross-checked against the real header compiled into oracle/_ref/libxxh3_ref.so. returns 0 for other lengths. */
uint64_t orc_xxh3_64(const void* data, int32_t bytes, uint64_t seed);
/* HashUtil::xorshift32 (:223-228) */
uint32_t orc_xorshift32(uint32_t x);
/* ReduceOp / ModuloOp (:236-244) */
uint32_t orc_reduce_op(uint32_t l, uint32_t r);

/* ---- filter -------------------------------------------------------------------------- */
/* column_filter_range::filter_range<T> (be/src/column/column_filter_range.cpp:39-148):
 * in-place order-preserving compaction of data[from,to) by filter; rows < from are kept.
 * returns the new size. */
int64_t orc_filter_range(const uint8_t* filter, void* data, int32_t elem_size, int64_t from, int64_t to);
/* ColumnPredicate::evaluate + AND-merge (column_operator_predicate.h:41-111,
 * chunk_predicate_evaluator.cpp:84-149): selection[i] = 1 iff every conjunct is true-and-not-null */
int32_t orc_scan_evaluate(const sr_scan_desc* desc, const sr_chunk_view* in, uint8_t* selection);
/* evaluate + Chunk::filter (chunk.cpp:362-372).  out_data[k] / out_nulls[k] are caller
 * allocated buffers (num_rows capacity) for desc->out_slots[k]; out_nulls[k] may be NULL when
 * the input column has no null column.  returns surviving rows or <0. */
int64_t orc_scan_filter(const sr_scan_desc* desc, const sr_chunk_view* in, void** out_data, uint8_t** out_nulls);
/* evaluate one expression over all rows: result in out_i (integer class) or out_d (double
 * class) according to *is_double; out_null[i] = 1 for NULL results. */
int32_t orc_eval_expr(const sr_expr* e, const sr_chunk_view* in, int64_t* out_i, double* out_d, uint8_t* out_null,
                      int32_t* is_double);

/* ---- hash join ----------------------------------------------------------------------- */
/* JoinHashMapMethodType (be/src/exec/join/join_hash_map_method_fwd.h) */
enum orc_join_method {
    ORC_BUCKET_CHAINED = 1,
    ORC_DIRECT_MAPPING = 2,
    ORC_RANGE_DIRECT_MAPPING = 3,
    ORC_RANGE_DIRECT_MAPPING_SET = 4,
    ORC_DENSE_RANGE_DIRECT_MAPPING = 5,
    ORC_LINEAR_CHAINED = 6,
    ORC_LINEAR_CHAINED_SET = 7
};

typedef struct orc_join orc_join;

typedef struct orc_join_options {
    int32_t enable_range_direct_mapping; /* enable_hash_join_range_direct_mapping_opt */
    int32_t enable_linear_chained;       /* enable_hash_join_linear_chained_opt */
    int64_t l2_cache_size;               /* CpuInfo::get_l2_cache_size() */
    int64_t l3_cache_size;               /* CpuInfo L3 (HALF_L3 = l3/2) */
    int32_t force_method;                /* 0 = follow JoinHashMapSelector */
    int32_t chunk_size;                  /* 0 = ORC_CHUNK_SIZE */
} orc_join_options;

typedef struct orc_probe_result {
    int64_t count; /* rows written to probe_index/build_index */
    int32_t has_remain;
    int32_t match_flag; /* 0 NORMAL, 1 ALL_MATCH_ONE, 2 MOST_MATCH_ONE (JoinMatchFlag) */
    int32_t cur_probe_index;
    int32_t cur_row_match_count;
} orc_probe_result;

orc_join* orc_join_create(const sr_join_desc* desc, const orc_join_options* opt);
void orc_join_destroy(orc_join* j);
/* JoinHashTable::append_chunk (join_hash_table.cpp:712-752) */
int32_t orc_join_append_build(orc_join* j, const sr_chunk_view* chunk);
/* JoinHashTable::build (:633-685) = selector (:161-350) + construct_hash_table */
int32_t orc_join_build(orc_join* j);
int32_t orc_join_method(const orc_join* j);
int64_t orc_join_build_rows(const orc_join* j);
int64_t orc_join_bucket_size(const orc_join* j);
int64_t orc_join_min_value(const orc_join* j);
int64_t orc_join_max_value(const orc_join* j);
const uint32_t* orc_join_first(const orc_join* j);
const uint32_t* orc_join_next(const orc_join* j);
/* One reference-style probe step on a chunk of <= chunk_size rows: first call with
 * first_probe = 1 does lookup_init + _probe_from_ht<true>; while has_remain, call again with
 * first_probe = 0 (_probe_from_ht<false>).  probe_index / build_index need chunk_size + 8
 * entries.  (join_hash_map.hpp:381-420, 718-795, 950-1010, 1188-1255) */
int32_t orc_join_probe_chunk(orc_join* j, const sr_chunk_view* probe, int32_t first_probe, uint32_t* probe_index,
                             uint32_t* build_index, orc_probe_result* res);
/* whole-batch probe: loops chunk by chunk, concatenating the index pairs (probe_index is
 * global over the batch).  cap = capacity of the two arrays; returns the match count, or the
 * required capacity as a negative number -(needed) when cap is too small. */
int64_t orc_join_probe_all(orc_join* j, const sr_chunk_view* probe, uint32_t* probe_index, uint32_t* build_index,
                           int64_t cap);
/* _probe_output / _build_output (join_hash_map.hpp:163-269): materialise the join output
 * columns for n index pairs: probe_out_slots then build_out_slots of the desc.
 * out_nulls[k] must be non-NULL for build columns of a LEFT OUTER join. */
/* POST_PROBE phase of RIGHT / FULL joins (join_hash_map.hpp:136-143,420-457) */
int64_t orc_join_probe_remain(orc_join* j, uint32_t* build_index, int64_t cap);
int32_t orc_join_output_remain(orc_join* j, int64_t n, const uint32_t* build_index, const int32_t* probe_types, void** out_data, uint8_t** out_nulls);
int32_t orc_join_output(orc_join* j, const sr_chunk_view* probe, int64_t n, const uint32_t* probe_index,
                        const uint32_t* build_index, void** out_data, uint8_t** out_nulls);

/* ---- hash aggregate ------------------------------------------------------------------- */
typedef struct orc_agg orc_agg;
orc_agg* orc_agg_create(const sr_agg_desc* desc);
void orc_agg_destroy(orc_agg* a);
/* Aggregator::build_hash_map + compute_batch_agg_states, chunk by chunk
 * (aggregator.cpp:907-929,1616-1640; agg_hash_map.h:303-415) */
int32_t orc_agg_push(orc_agg* a, const sr_chunk_view* chunk);
int64_t orc_agg_num_groups(const orc_agg* a);
/* convert_hash_map_to_chunk (aggregator.cpp:1696-1791): groups in state-arena (insertion)
 * order; key columns then result columns. out_data[k] sized num_groups * width. */
int32_t orc_agg_output(orc_agg* a, void** out_data, uint8_t** out_nulls);
/* result type / width of output column k (keys first) */
int32_t orc_agg_out_type(const orc_agg* a, int32_t k);
int32_t orc_agg_num_out_cols(const orc_agg* a);
/* AggregateFunction::merge of another aggregator's states */
int32_t orc_agg_merge(orc_agg* a, const orc_agg* other);
/* pass-through leg of the streaming aggregate: rows -> intermediate rows (Aggregator::output_chunk_by_streaming) */
int32_t orc_agg_convert_to_states(const sr_agg_desc* first_phase_desc, const sr_chunk_view* chunk, void** out_data, uint8_t** out_nulls);

/* ---- exchange partitioning ---------------------------------------------------------- */
/* ExchangeSinkOperator hash + Shuffler::exchange_shuffle + counting sort
 * (exchange_sink_operator.cpp:586-637, shuffler.h:72-89): hash_values / channel_ids
 * (num_rows each), row_indexes (num_rows, rows of channel c at [starts[c], starts[c+1])),
 * channel_starts (num_channels + 1). processed in ORC_CHUNK_SIZE pieces is not needed for
 * hashing (pure per-row), the counting sort is done over the whole batch. */
int32_t orc_hash_partition(const sr_part_desc* desc, const sr_chunk_view* in, uint32_t* hash_values,
                           uint32_t* channel_ids, uint32_t* row_indexes, int64_t* channel_starts);

/* ---- exchange wire format -------------------------------------------------------------- */
/* ChunkPB.data at encode level 0 (ProtobufChunkSerde::serialize_without_meta, be/src/serde/protobuf_serde.cpp:88-140;
 * FixedLengthColumnSerde / NullableColumnSerde, column_array_serde.cpp:214-255,759-782): fixed32 version 1, fixed32 rows, per
 * column [null column: fixed32 n + n bytes] fixed32 byte size + raw values.  rows [row_begin, row_end).  returns the number of
 * bytes written (dst may be NULL to size the buffer), < 0 on error. */
int64_t orc_chunk_serialize(const sr_chunk_view* chunk, int64_t row_begin, int64_t row_end, uint8_t* dst, int64_t cap);

/* ---- whole pipeline (CPU BE stand-in) ------------------------------------------------ */
typedef struct orc_frag_join {
    orc_join* join;
    int32_t probe_key_slot;
    int32_t num_payload;
    int32_t payload_build_slots[SR_MAX_FRAG_PAYLOAD];
} orc_frag_join;

typedef struct orc_fragment_desc {
    sr_scan_desc scan;
    int32_t num_joins;
    int32_t reserved;
    orc_frag_join joins[SR_MAX_FRAG_JOINS];
    sr_agg_desc agg;
} orc_fragment_desc;

/* Runs scan(filter) -> probe x N -> aggregate chunk-at-a-time (4096 rows) exactly like one
 * pipeline driver per thread would (pipeline_driver.cpp:270-500): each operator materialises
 * its output chunk before the next consumes it.  `num_threads` drivers each own a partial
 * aggregator which is merged at the end (two-phase aggregate).  Result lands in *result (a
 * fresh aggregator created from desc->agg by the caller).  rows_passed (optional) receives the
 * number of fact rows that reached the aggregate. */
int32_t orc_fragment_run(const orc_fragment_desc* desc, const sr_chunk_view* fact, int32_t num_threads,
                         orc_agg* result, int64_t* rows_passed);

/* ---- runtime filter: MinMaxRuntimeFilter + SimdBlockFilter (be/src/runtime/runtime_filter.h:58-59,79-240,584-,
 * 1270-1290; runtime_filter.cpp:26-35,114-123).  Pinned by be/test/runtime/runtime_filter_core_test.cpp:49-123. */
typedef struct orc_rf orc_rf;
orc_rf* orc_rf_create(int32_t key_type, int64_t expected_rows, int32_t with_bloom);
void orc_rf_destroy(orc_rf* rf);
void orc_rf_insert_hash(orc_rf* rf, uint64_t hash);    /* SimdBlockFilter::insert_hash */
int32_t orc_rf_test_hash(const orc_rf* rf, uint64_t hash); /* SimdBlockFilter::test_hash */
uint64_t orc_rf_value_hash(int64_t value);             /* phmap_mix<8>(std::hash<T>(value)) for integer-class T */
int32_t orc_rf_insert(orc_rf* rf, const sr_chunk_view* in, int32_t slot_id, int32_t insert_nulls);
int32_t orc_rf_merge(orc_rf* rf, const orc_rf* other); /* SimdBlockFilter::merge + min/max/has_null union */
int32_t orc_rf_evaluate(const orc_rf* rf, const sr_chunk_view* in, int32_t slot_id, uint8_t* selection, int32_t merge_and);
int32_t orc_rf_get_info(const orc_rf* rf, sr_rf_info* info);
const void* orc_rf_directory(const orc_rf* rf, int64_t* bytes);

/* ---- segment pages: frame-of-reference pages (be/src/util/frame_of_reference_coding.{h,cpp}:
 * ForEncoder::put_batch / bit_packing_one_frame_value / flush, ForDecoder::init / decode_current_frame; page wrapper
 * be/src/storage/rowset/frame_of_reference_page.h) and plain pages (plain_page.h:51-158).  elem_size 4 or 8 (signed
 * integers, like the CppType of TYPE_INT / TYPE_BIGINT / TYPE_DATE).  Pinned byte for byte against the reference's own codec
 * compiled into oracle/_ref/libfor_ref.so.  encode: returns the page size, or -needed when cap is too small.
 * decode: returns the number of values, -1 for a corrupt page, -n when cap < n. */
int64_t orc_for_encode(int32_t elem_size, const void* values, int64_t n, uint8_t* out, int64_t cap);
int64_t orc_for_decode(int32_t elem_size, const uint8_t* page, int64_t len, void* out, int64_t cap);
int64_t orc_plain_encode(int32_t elem_size, const void* values, int64_t n, uint8_t* out, int64_t cap);
int64_t orc_plain_decode(int32_t elem_size, const uint8_t* page, int64_t len, void* out, int64_t cap);

/* Aggregator::build_hash_map_with_selection (SELECTIVE_PREAGG): selection[i] = 1 when row i's group is not in the table */
int32_t orc_agg_streaming_selection(const orc_agg* a, const sr_chunk_view* chunk, uint8_t* selection);

const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
