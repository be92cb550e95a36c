// CPU-only test of the host string column and the low-cardinality dictionary path around the GPU operators
// (starrocks_b200/host/column/binary_column.h, host/exec/global_dict.h).  The cases follow the reference's own
//   be/test/column/binary_column_test.cpp (test_get_data :105, test_byte_size :117, test_filter :130,
//   test_append_strings :152, test_compare_at :237, test_filter_range :284, test_resize :314, test_assign :335,
//   test_xor_checksum :586 -- golden 3546653113525744178 --, test_replicate :600, test_append_selective :633)
// plus the dictionary round trip (dictcode_column_iterator.cpp, decoder.cpp, dict_decode_operator.cpp) run through the
// PipelineDriver loop.  Needs no GPU and does not link libsr_gpu.so (sr_type_width is provided below).
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../starrocks_b200/host/exec/global_dict.h"

// the one libsr_gpu symbol chunk.h refers to (make_column); same table as the library
extern "C" int32_t sr_type_width(int32_t type) {
    switch (type) {
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

using namespace starrocks;
using namespace starrocks::pipeline;

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        g_checks++;                                                          \
        if (!(cond)) {                                                       \
            g_failed++;                                                      \
            fprintf(stderr, "%s:%d: CHECK(%s) failed\n", __FILE__, __LINE__, #cond); \
        }                                                                    \
    } while (0)
#define CHECK_EQ(a, b) CHECK((a) == (b))

static void test_get_data() {
    auto column = BinaryColumn::create();
    for (int i = 0; i < 100; i++) column->append_string(std::string("str:") + std::to_string(i));
    const auto& slices = column->immutable_data();
    CHECK_EQ(slices.size(), 100u);
    for (size_t i = 0; i < slices.size(); i++) CHECK_EQ(std::string("str:") + std::to_string(i), slices[i].to_string());
}

static void test_byte_size() {
    auto column = BinaryColumn::create();
    CHECK_EQ(sizeof(BinaryColumn::Offset), column->byte_size());
    std::string s("test_string");
    for (int i = 0; i < 10; i++) column->append_string(s);
    CHECK_EQ(10u, column->size());
    CHECK_EQ(10 * s.size() + 11 * sizeof(BinaryColumn::Offset), column->byte_size());
    CHECK_EQ(s.size() + sizeof(uint32_t), column->byte_size(3));
}

static void test_filter() {
    auto column = BinaryColumn::create();
    for (int i = 0; i < 100; ++i) column->append_string(std::to_string(i));
    Filter filter;
    for (int k = 0; k < 100; ++k) filter.push_back(k % 2);
    column->filter(filter);
    CHECK_EQ(50u, column->size());
    const auto& slices = column->immutable_data();
    for (int i = 0; i < 50; ++i) CHECK_EQ(std::to_string(i * 2 + 1), slices[i].to_string());
    column->check_or_die();
}

static void test_append_strings() {
    std::vector<Slice> values{{"hello"}, {"starrocks"}};
    auto c1 = BinaryColumn::create();
    CHECK(c1->append_strings(values.data(), values.size()));
    CHECK_EQ(values.size(), c1->size());
    std::vector<Slice> values2{{"abcd"}, {"123456"}};
    CHECK(c1->append_strings(values2.data(), values2.size()));
    CHECK_EQ(4u, c1->size());
    for (size_t i = 0; i < 2; i++) CHECK(values[i] == c1->immutable_data()[i]);
    for (size_t i = 0; i < 2; i++) CHECK(values2[i] == c1->immutable_data()[i + 2]);
    CHECK(!c1->append_nulls(1));
    c1->append_default(2);
    CHECK_EQ(6u, c1->size());
    CHECK(c1->get_slice(5).empty());
}

static void test_compare_at() {
    std::vector<Slice> strings{{"bbb"}, {"bbc"}, {"ccc"}};
    auto c1 = BinaryColumn::create();
    auto c2 = BinaryColumn::create();
    c1->append_strings(strings.data(), strings.size());
    c2->append_strings(strings.data(), strings.size());
    for (size_t i = 0; i < 3; i++) CHECK_EQ(0, c1->compare_at(i, i, *c2));
    for (size_t i = 0; i < 3; i++)
        for (size_t j = i + 1; j < 3; j++) {
            CHECK(c1->compare_at(i, j, *c2) < 0);
            CHECK(c2->compare_at(j, i, *c1) > 0);
        }
    CHECK(Slice("ab") < Slice("abc")); // shorter first on a common prefix
    CHECK(Slice("") < Slice("a"));
}

static void test_filter_range() {
    auto column = BinaryColumn::create();
    column->append(Slice("m"));
    for (size_t i = 0; i < 63; i++) column->append(Slice("a"));
    column->append(Slice("bbbbbfffff"));
    column->append(Slice("c"));
    Filter filter;
    filter.push_back(0);
    for (size_t i = 0; i < 65; i++) filter.push_back(1);
    CHECK_EQ(65u, column->filter_range(filter, 0, 66));
    const auto& data = column->immutable_data();
    for (size_t i = 0; i < 63; i++) CHECK(data[i] == Slice("a"));
    CHECK(data[63] == Slice("bbbbbfffff"));
    CHECK(data[64] == Slice("c"));
    column->check_or_die();

    // rows before `from` stay, rows at and after `to` go
    auto c = BinaryColumn::create();
    for (int i = 0; i < 10; i++) c->append_string("v" + std::to_string(i));
    Filter f(10, 0);
    f[4] = f[6] = 1;
    CHECK_EQ(5u, c->filter_range(f, 3, 8)); // v0 v1 v2 | v4 v6
    const char* want[] = {"v0", "v1", "v2", "v4", "v6"};
    for (int i = 0; i < 5; i++) CHECK(c->get_slice(i) == Slice(want[i]));
}

static void test_filter_against_row_at_a_time() {
    std::mt19937 rng(7);
    for (int round = 0; round < 50; round++) {
        const size_t n = rng() % 700;
        auto c = BinaryColumn::create();
        std::vector<std::string> rows(n);
        for (auto& r : rows) {
            r.assign(rng() % 9, (char)('a' + rng() % 26)); // lengths 0..8
            c->append_string(r);
        }
        Filter f(n);
        const int density = rng() % 4; // 0: sparse, 3: dense
        for (auto& b : f) b = (rng() % 4) <= (unsigned)density ? 1 : 0;
        std::vector<std::string> want;
        for (size_t i = 0; i < n; i++)
            if (f[i]) want.push_back(rows[i]);
        CHECK_EQ(want.size(), c->filter(f));
        c->check_or_die();
        bool same = c->size() == want.size();
        for (size_t i = 0; same && i < want.size(); i++) same = c->get_slice(i) == Slice(want[i]);
        CHECK(same);
    }
}

static void test_resize_and_assign() {
    auto c = BinaryColumn::create();
    c->append(Slice("abc"));
    c->append(Slice("def"));
    c->append(Slice("xyz"));
    c->resize(1);
    CHECK_EQ(1u, c->size());
    CHECK(c->get_slice(0) == Slice("abc"));
    c->append(Slice("xxxxxx"));
    CHECK_EQ(2u, c->size());
    CHECK(c->get_slice(1) == Slice("xxxxxx"));
    c->resize(4);
    CHECK_EQ(4u, c->size());
    CHECK(c->get_slice(0) == Slice("abc"));
    CHECK(c->get_slice(1) == Slice("xxxxxx"));
    CHECK(c->get_slice(2) == Slice(""));
    CHECK(c->get_slice(3) == Slice(""));
    c->check_or_die();

    std::vector<Slice> strings{{"bbb"}, {"bbc"}, {"ccc"}};
    auto c2 = BinaryColumn::create();
    c2->append_strings(strings.data(), 3);
    c2->assign(c2->size(), 2);
    for (size_t i = 0; i < 3; i++) CHECK(c2->get_slice(i) == strings[2]);
}

static void test_xor_checksum() {
    auto column = BinaryColumn::create();
    std::string str;
    for (int i = 0; i <= 1000; i++) str.append(std::to_string(i));
    column->append_string(str);
    CHECK_EQ(column->xor_checksum(0, 1), 3546653113525744178LL);
}

static void test_replicate_and_append_selective() {
    auto c1 = BinaryColumn::create();
    c1->append(Slice("abc"));
    c1->append(Slice("def"));
    auto c2 = c1->replicate({0, 3, 5});
    CHECK_EQ(5u, c2->size());
    for (int i = 0; i < 3; i++) CHECK(c2->get_slice(i) == Slice("abc"));
    for (int i = 3; i < 5; i++) CHECK(c2->get_slice(i) == Slice("def"));

    for (uint32_t num_rows : {2048u, 4096u, 40960u}) {
        auto src = BinaryColumn::create();
        for (uint32_t i = 0; i < num_rows; i++) src->append_string(std::string(i % 16 + 8, (char)('a' + (i % 26))));
        std::vector<uint32_t> indexes;
        for (uint32_t i = 0; i < num_rows; i += 16) indexes.push_back(i);
        auto dst = BinaryColumn::create();
        dst->append_selective(*src, indexes.data(), 0, (uint32_t)indexes.size());
        const size_t n0 = dst->size();
        CHECK_EQ(indexes.size(), n0);
        bool same = true;
        for (uint32_t i = 0; i < n0; i++) same = same && src->get_slice(indexes[i]) == dst->get_slice(i);
        dst->append_selective(*src, indexes.data(), 10, (uint32_t)indexes.size() - 10);
        CHECK_EQ(n0 + indexes.size() - 10, dst->size());
        for (uint32_t i = 10; i < indexes.size(); i++) same = same && src->get_slice(indexes[i]) == dst->get_slice(n0 + i - 10);
        CHECK(same);
        dst->check_or_die();
    }
    // append of a row range and of one value many times
    auto a = BinaryColumn::create();
    a->append(Slice("x"));
    auto b = BinaryColumn::create();
    for (int i = 0; i < 6; i++) b->append_string("row" + std::to_string(i));
    a->append(*b, 2, 3);
    CHECK_EQ(4u, a->size());
    CHECK(a->get_slice(1) == Slice("row2") && a->get_slice(3) == Slice("row4"));
    a->append_value_multiple_times(*a, 1, 2);
    CHECK(a->size() == 6 && a->get_slice(5) == Slice("row2"));
    a->check_or_die();
}

static void test_wire_format() {
    // BinaryColumnSerde at encode level 0 (column_array_serde.cpp:279-313): ["ab", "", "cde"]
    auto c = BinaryColumn::create();
    c->append(Slice("ab"));
    c->append(Slice(""));
    c->append(Slice("cde"));
    const uint8_t want[] = {5, 0, 0, 0, 'a', 'b', 'c', 'd', 'e', 16, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 2, 0, 0, 0, 5, 0, 0, 0};
    CHECK_EQ((size_t)c->max_serialized_size(), sizeof(want));
    std::vector<uint8_t> buf(sizeof(want));
    CHECK(c->serialize(buf.data()) == buf.data() + buf.size());
    CHECK(memcmp(buf.data(), want, sizeof(want)) == 0);
    auto d = BinaryColumn::create();
    CHECK(d->deserialize(buf.data(), buf.data() + buf.size()) == buf.data() + buf.size());
    CHECK(d->size() == 3 && d->get_slice(0) == Slice("ab") && d->get_slice(1).empty() && d->get_slice(2) == Slice("cde"));
    CHECK(d->deserialize(buf.data(), buf.data() + buf.size() - 1) == nullptr); // truncated
    buf[17] = 9;                                                                // offsets no longer sorted / consistent
    CHECK(d->deserialize(buf.data(), buf.data() + buf.size()) == nullptr);
    CHECK(d->size() == 3); // a failed read leaves the column alone
}

static void test_chunk_slice_with_strings() {
    auto strs = BinaryColumn::create();
    auto ints = std::make_shared<Int32Column>(SR_TYPE_INT);
    auto nulls = std::make_shared<NullColumn>(SR_TYPE_BOOLEAN);
    for (int i = 0; i < 10; i++) {
        strs->append_string(std::string(i, 'z'));
        ints->append(i * 3);
        nulls->append(i % 4 == 0);
    }
    Chunk chunk;
    chunk.append_column(std::make_shared<NullableColumn>(strs, nulls), 7);
    chunk.append_column(ints, 9);
    auto part = chunk.slice(4, 3);
    CHECK_EQ(3u, part->num_rows());
    const auto& col = part->get_column_by_slot_id(7);
    CHECK(col->is_nullable() && col->is_binary() && col->logical_type() == TYPE_VARCHAR);
    auto* data = static_cast<const BinaryColumn*>(static_cast<const NullableColumn&>(*col).data_column().get());
    CHECK(data->get_slice(0) == Slice("zzzz") && data->get_slice(2) == Slice("zzzzzz"));
    CHECK(col->null_data()[0] == 1 && col->null_data()[1] == 0);
    CHECK_EQ(reinterpret_cast<const int32_t*>(part->get_column_by_slot_id(9)->raw_data())[1], 15);
    CHECK_EQ(sr_type_width(TYPE_VARCHAR), 0); // the C-ABI has no string type: such a column cannot be handed to the GPU library
}

// ---- the dictionary path ------------------------------------------------------------------------------------------------
class VectorSource final : public SourceOperator {
public:
    VectorSource(std::vector<ChunkPtr> chunks) : SourceOperator(nullptr, 0, "vector_source", 0, false, 0), _chunks(std::move(chunks)) {}
    bool has_output() const override { return _next < _chunks.size(); }
    bool is_finished() const override { return _next >= _chunks.size(); }
    StatusOr<ChunkPtr> pull_chunk(RuntimeState*) override { return _chunks[_next++]; }

private:
    std::vector<ChunkPtr> _chunks;
    size_t _next = 0;
};
class CollectSink final : public Operator {
public:
    CollectSink() : Operator(nullptr, 2, "collect_sink", 2, false, 0) {}
    bool has_output() const override { return false; }
    bool need_input() const override { return !_finishing; }
    bool is_finished() const override { return _finishing; }
    Status set_finishing(RuntimeState*) override {
        _finishing = true;
        return Status::OK();
    }
    StatusOr<ChunkPtr> pull_chunk(RuntimeState*) override { return Status::InternalError("sink"); }
    Status push_chunk(RuntimeState*, const ChunkPtr& c) override {
        chunks.push_back(c);
        return Status::OK();
    }
    std::vector<ChunkPtr> chunks;

private:
    bool _finishing = false;
};

static void test_low_cardinality_dictionary_path() {
    // a "c_region"-like column with NULLs; the FE's dictionary of it; two segments with different local dictionaries
    const std::vector<std::string> regions{"AMERICA", "ASIA", "AFRICA", "EUROPE", "MIDDLE EAST"};
    GlobalDictLists lists{make_sorted_global_dict(/*columnId=*/5, regions)};
    CHECK(lists[0].strings[0] == "AFRICA" && lists[0].ids[0] == 1 && lists[0].strings[4] == "MIDDLE EAST" && lists[0].ids[4] == 5);
    FragmentDictState dict_state;
    CHECK(dict_state.init_query_global_dict(lists).ok());
    const auto& maps = dict_state.query_global_dicts().at(5);
    CHECK_EQ(maps.first.at(Slice("ASIA")), 3);
    CHECK(maps.second.at(4) == Slice("EUROPE"));

    std::mt19937 rng(3);
    std::vector<ChunkPtr> id_chunks;
    std::vector<std::string> truth;
    std::vector<uint8_t> truth_null;
    for (int segment = 0; segment < 2; segment++) {
        const size_t n = 3000 + 500 * segment;
        auto words = BinaryColumn::create();
        std::vector<uint8_t> nulls(n);
        for (size_t i = 0; i < n; i++) {
            const auto& w = regions[(rng() % (segment ? 5 : 3)) + 0]; // segment 0 never sees the last two words
            nulls[i] = rng() % 11 == 0;
            words->append_string(nulls[i] ? "" : w);
            truth.push_back(nulls[i] ? "" : w);
            truth_null.push_back(nulls[i]);
        }
        // what the segment stores: a local dictionary (first-seen order) + local codes, -1 in NULL rows
        BinaryColumn local_words;
        std::vector<int32_t> codes;
        encode_local_dict(*words, &local_words, &codes);
        for (size_t i = 0; i < n; i++)
            if (nulls[i]) codes[i] = -1;
        GlobalDictCodeConverter conv;
        CHECK(conv.build_code_convert_map(local_words, maps.first).ok()); // the empty word of NULL rows is tolerated
        auto ids = conv.decode_string_dict_codes(codes.data(), nulls.data(), n);
        bool ids_ok = true;
        for (size_t i = 0; i < n; i++) ids_ok = ids_ok && ids->get_data()[i] == (nulls[i] ? 0 : maps.first.at(words->get_slice(i)));
        CHECK(ids_ok);
        auto nullcol = std::make_shared<NullColumn>(SR_TYPE_BOOLEAN, nulls);
        auto revenue = std::make_shared<Int64Column>(SR_TYPE_BIGINT);
        for (size_t i = 0; i < n; i++) revenue->append((int64_t)i);
        auto chunk = std::make_shared<Chunk>();
        chunk->append_column(revenue, 1);
        chunk->append_column(std::make_shared<NullableColumn>(ids, nullcol), 5); // what the GPU operators would see: TYPE_INT
        CHECK_EQ(chunk->get_column_by_slot_id(5)->logical_type(), (int32_t)SR_TYPE_INT);
        id_chunks.push_back(chunk);
    }
    // an unknown word fails the scan-side translation with the reference's message
    {
        BinaryColumn local_words;
        local_words.append(Slice("ATLANTIS"));
        GlobalDictCodeConverter conv;
        Status st = conv.build_code_convert_map(local_words, maps.first);
        CHECK(!st.ok() && st.message() == "not found slice:ATLANTIS in global dict");
    }

    // source -> dict_decode -> sink through the driver loop; slot 5 (ids) comes out as slot 15 (strings), order kept
    DictDecodeOperatorFactory factory(1, 1, {5}, {15}, &dict_state);
    RuntimeState state(4096);
    CHECK(factory.prepare(&state).ok());
    auto sink = std::make_shared<CollectSink>();
    PipelineDriver driver({std::make_shared<VectorSource>(id_chunks), factory.create(1, 0), sink});
    CHECK(driver.prepare(&state).ok());
    auto st = driver.process(&state);
    CHECK(st.ok() && st.value() == PipelineDriver::FINISH);
    CHECK_EQ(sink->chunks.size(), 2u);
    size_t row = 0;
    bool same = true;
    for (auto& c : sink->chunks) {
        CHECK(c->num_columns() == 2 && c->slot_of_index(0) == 1 && c->slot_of_index(1) == 15 && !c->is_slot_exist(5));
        const auto& col = c->get_column_by_slot_id(15);
        CHECK(col->is_nullable() && col->is_binary());
        auto* data = static_cast<const BinaryColumn*>(static_cast<const NullableColumn&>(*col).data_column().get());
        for (size_t i = 0; i < c->num_rows(); i++, row++) {
            same = same && col->null_data()[i] == truth_null[row];
            same = same && data->get_slice(i) == Slice(truth[row]);
        }
    }
    CHECK(same && row == truth.size());

    // an id outside the dictionary, and a plan that names a column without a dictionary
    auto bad = std::make_shared<Int32Column>(SR_TYPE_INT);
    bad->append(42);
    auto r = create_global_dict_decoder(maps.second)->decode_string(*bad);
    CHECK(!r.ok() && r.status().message() == "Dict Decode failed, Dict can't take cover all key :42");
    DictDecodeOperatorFactory no_dict(1, 1, {6}, {16}, &dict_state);
    Status st2 = no_dict.prepare(&state);
    CHECK(!st2.ok() && st2.message() == "Not found dict for cid:6");
}

int main() {
    test_get_data();
    test_byte_size();
    test_filter();
    test_append_strings();
    test_compare_at();
    test_filter_range();
    test_filter_against_row_at_a_time();
    test_resize_and_assign();
    test_xor_checksum();
    test_replicate_and_append_selective();
    test_wire_format();
    test_chunk_slice_with_strings();
    test_low_cardinality_dictionary_path();
    if (g_failed) {
        fprintf(stderr, "BINARY_COLUMN_TEST_FAILED %d of %d checks\n", g_failed, g_checks);
        return 1;
    }
    printf("BINARY_COLUMN_TEST_OK %d checks\n", g_checks);
    return 0;
}
