// BinaryColumn for the host side of the GPU operators: the VARCHAR / CHAR column of the reference
//   be/src/column/binary_column.h:30-467, binary_column.cpp (offsets + bytes, Slice accessors, filter_range,
//   append_selective, xor_checksum), be/src/base/string/slice.h (Slice),
//   be/src/serde/column_array_serde.cpp:256-360 (BinaryColumnSerde: the column's bytes inside ChunkPB.data).
// Strings never cross the C-ABI: the GPU operators group, join and filter on the dictionary codes the low-cardinality
// rewrite puts in their place (exec/global_dict.h, the reference's runtime/global_dict); a BinaryColumn handed to
// make_chunk_view carries TYPE_VARCHAR, which libsr_gpu refuses with SR_ERR_NOT_SUPPORTED.  This class exists so the
// plan around the GPU operators (dictionary build at the scan, DictDecodeOperator at the top) has the reference's
// string column to work with.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "chunk.h"

namespace starrocks {

// be/src/types/logical_type.h:43.  Not an sr_type: sr_type_width(TYPE_VARCHAR) == 0.
constexpr int32_t TYPE_VARCHAR = 17;

struct Slice {
    const char* data = "";
    size_t size = 0;
    Slice() = default;
    Slice(const char* d, size_t n) : data(d), size(n) {}
    Slice(const uint8_t* d, size_t n) : data(reinterpret_cast<const char*>(d)), size(n) {}
    Slice(const std::string& s) : data(s.data()), size(s.size()) {}
    Slice(const char* s) : data(s), size(strlen(s)) {}
    bool empty() const { return size == 0; }
    std::string to_string() const { return std::string(data, size); }
    // byte order, shorter first on a common prefix (Slice::compare)
    int compare(const Slice& b) const {
        const size_t n = std::min(size, b.size);
        const int r = n == 0 ? 0 : memcmp(data, b.data, n);
        if (r != 0) return r;
        return size < b.size ? -1 : (size > b.size ? 1 : 0);
    }
};
inline bool operator==(const Slice& a, const Slice& b) { return a.size == b.size && (a.size == 0 || memcmp(a.data, b.data, a.size) == 0); }
inline bool operator!=(const Slice& a, const Slice& b) { return !(a == b); }
inline bool operator<(const Slice& a, const Slice& b) { return a.compare(b) < 0; }

using Filter = std::vector<uint8_t>;

class BinaryColumn final : public Column {
public:
    using Offset = uint32_t;
    using Offsets = std::vector<Offset>;
    using Bytes = std::vector<uint8_t>;
    using Container = std::vector<Slice>;

    BinaryColumn() : _offsets(1, 0) {}
    BinaryColumn(Bytes bytes, Offsets offsets) : _bytes(std::move(bytes)), _offsets(std::move(offsets)) {
        if (_offsets.empty()) _offsets.push_back(0);
        check_or_die();
    }
    static std::shared_ptr<BinaryColumn> create() { return std::make_shared<BinaryColumn>(); }

    // ---- Column ----
    size_t size() const override { return _offsets.size() - 1; }
    int32_t logical_type() const override { return TYPE_VARCHAR; }
    bool is_binary() const override { return true; }
    const uint8_t* raw_data() const override { return _bytes.data(); }
    uint8_t* mutable_raw_data() override { return _bytes.data(); }
    size_t type_size() const override { return sizeof(Slice); }
    // shrink: drop the tail rows and their bytes; grow: empty strings
    void resize(size_t n) override {
        if (n < size()) _bytes.resize(_offsets[n]);
        _offsets.resize(n + 1, (Offset)_bytes.size());
        _slices_valid = false;
    }
    ColumnPtr cut(size_t offset, size_t n) const override {
        auto out = create();
        out->append(*this, offset, n);
        return out;
    }

    // ---- sizes ----
    size_t byte_size() const { return _bytes.size() + _offsets.size() * sizeof(Offset); }
    size_t byte_size(size_t idx) const { return _offsets[idx + 1] - _offsets[idx] + sizeof(uint32_t); }
    size_t byte_size(size_t from, size_t n) const { return _offsets[from + n] - _offsets[from] + n * sizeof(Offset); }
    void reserve(size_t rows, size_t bytes) {
        _offsets.reserve(rows + 1);
        _bytes.reserve(bytes);
    }

    // ---- element access ----
    Slice get_slice(size_t i) const { return Slice(_bytes.data() + _offsets[i], _offsets[i + 1] - _offsets[i]); }
    Bytes& get_bytes() { return _bytes; }
    const Bytes& get_bytes() const { return _bytes; }
    Offsets& get_offset() { return _offsets; }
    const Offsets& get_offset() const { return _offsets; }
    // one Slice per row, rebuilt lazily after a mutation (BinaryColumn::immutable_data / build_slices)
    const Container& immutable_data() const {
        if (!_slices_valid) {
            _slices.resize(size());
            for (size_t i = 0; i < size(); i++) _slices[i] = get_slice(i);
            _slices_valid = true;
        }
        return _slices;
    }
    std::string debug_item(size_t i) const { return "'" + get_slice(i).to_string() + "'"; }

    // ---- appends ----
    void append(const Slice& s) {
        check_capacity(s.size);
        _bytes.insert(_bytes.end(), s.data, s.data + s.size);
        _offsets.push_back((Offset)_bytes.size());
        _slices_valid = false;
    }
    void append_string(const std::string& s) { append(Slice(s)); }
    bool append_strings(const Slice* data, size_t n) {
        size_t total = 0;
        for (size_t i = 0; i < n; i++) total += data[i].size;
        check_capacity(total);
        _bytes.reserve(_bytes.size() + total);
        _offsets.reserve(_offsets.size() + n);
        for (size_t i = 0; i < n; i++) {
            _bytes.insert(_bytes.end(), data[i].data, data[i].data + data[i].size);
            _offsets.push_back((Offset)_bytes.size());
        }
        _slices_valid = false;
        return true;
    }
    bool append_nulls(size_t) { return false; } // a BinaryColumn holds no NULLs; NullableColumn does
    void append_default() { append_default(1); }
    void append_default(size_t count) {
        _offsets.insert(_offsets.end(), count, (Offset)_bytes.size());
        _slices_valid = false;
    }
    void append(const BinaryColumn& src, size_t offset, size_t count) {
        const Offset b0 = src._offsets[offset], b1 = src._offsets[offset + count];
        check_capacity(b1 - b0);
        const Offset shift = (Offset)_bytes.size() - b0; // modulo 2^32 on purpose
        _bytes.insert(_bytes.end(), src._bytes.begin() + b0, src._bytes.begin() + b1);
        _offsets.reserve(_offsets.size() + count);
        for (size_t i = 1; i <= count; i++) _offsets.push_back(src._offsets[offset + i] + shift);
        _slices_valid = false;
    }
    void append_selective(const BinaryColumn& src, const uint32_t* indexes, uint32_t from, uint32_t n) {
        size_t total = 0;
        for (uint32_t i = 0; i < n; i++) total += src._offsets[indexes[from + i] + 1] - src._offsets[indexes[from + i]];
        check_capacity(total);
        size_t pos = _bytes.size();
        _bytes.resize(pos + total);
        _offsets.reserve(_offsets.size() + n);
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t r = indexes[from + i];
            const size_t len = src._offsets[r + 1] - src._offsets[r];
            if (len) memcpy(_bytes.data() + pos, src._bytes.data() + src._offsets[r], len);
            pos += len;
            _offsets.push_back((Offset)pos);
        }
        _slices_valid = false;
    }
    void append_value_multiple_times(const BinaryColumn& src, uint32_t index, uint32_t n) {
        const std::string v = src.get_slice(index).to_string(); // src may be *this
        for (uint32_t i = 0; i < n; i++) append(Slice(v));
    }
    // n copies of row idx, replacing the content (Column::assign)
    void assign(size_t n, size_t idx) {
        const std::string v = get_slice(idx).to_string();
        _bytes.clear();
        _offsets.assign(1, 0);
        for (size_t i = 0; i < n; i++) append(Slice(v));
    }
    // row i repeated offsets[i + 1] - offsets[i] times (Column::replicate)
    std::shared_ptr<BinaryColumn> replicate(const std::vector<uint32_t>& rep_offsets) const {
        auto out = create();
        for (size_t i = 0; i + 1 < rep_offsets.size() && i < size(); i++)
            for (uint32_t k = rep_offsets[i]; k < rep_offsets[i + 1]; k++) out->append(get_slice(i));
        return out;
    }

    // ---- filter ----
    // Keeps rows [0, from) and, of [from, to), those whose filter byte is non-zero; rows at and after `to` are dropped
    // (Column::filter_range contract).  Runs of kept rows move with one memmove each.  Returns the new row count.
    size_t filter_range(const Filter& filter, size_t from, size_t to) {
        size_t out_rows = from;
        Offset out_bytes = _offsets[from];
        size_t i = from;
        while (i < to) {
            if (!filter[i]) {
                i++;
                continue;
            }
            size_t j = i + 1;
            while (j < to && filter[j]) j++;
            const Offset b0 = _offsets[i], len = _offsets[j] - b0;
            if (b0 != out_bytes && len) memmove(_bytes.data() + out_bytes, _bytes.data() + b0, len);
            const Offset shift = out_bytes - b0;
            for (size_t r = i; r < j; r++) _offsets[out_rows + (r - i) + 1] = _offsets[r + 1] + shift;
            out_rows += j - i;
            out_bytes += len;
            i = j;
        }
        _offsets.resize(out_rows + 1);
        _bytes.resize(out_bytes);
        _slices_valid = false;
        return out_rows;
    }
    size_t filter(const Filter& filter) { return filter_range(filter, 0, std::min(filter.size(), size())); }

    // ---- copies ----
    std::shared_ptr<BinaryColumn> clone_empty() const { return create(); }
    std::shared_ptr<BinaryColumn> clone() const { return std::make_shared<BinaryColumn>(_bytes, _offsets); }
    void swap_column(BinaryColumn& rhs) {
        _bytes.swap(rhs._bytes);
        _offsets.swap(rhs._offsets);
        _slices_valid = rhs._slices_valid = false;
    }
    void reset_column() {
        _bytes.clear();
        _offsets.assign(1, 0);
        _slices_valid = false;
    }

    // ---- comparisons / checksums ----
    int compare_at(size_t left, size_t right, const BinaryColumn& rhs) const { return get_slice(left).compare(rhs.get_slice(right)); }
    // XOR of every row's bytes taken as little-endian 64-bit words, the last <8 bytes one by one
    // (binary_column.cpp:846-886; the AVX2 lanes of the reference fold to the same value because XOR commutes)
    int64_t xor_checksum(uint32_t from, uint32_t to) const {
        int64_t sum = 0;
        for (size_t i = from; i < to; i++) {
            const uint8_t* p = _bytes.data() + _offsets[i];
            size_t n = _offsets[i + 1] - _offsets[i];
            // the reference consumes 32-byte blocks first, then 8-byte words: both are 8-byte words at the same positions
            for (; n >= 8; n -= 8, p += 8) {
                int64_t w;
                memcpy(&w, p, 8);
                sum ^= w;
            }
            for (size_t k = 0; k < n; k++) sum ^= p[k];
        }
        return sum;
    }

    // ---- wire format (encode level 0): [u32 byte count][bytes][u32 size of the offset array in bytes][offsets] ----
    int64_t max_serialized_size() const { return (int64_t)(sizeof(uint32_t) * 2 + _bytes.size() + _offsets.size() * sizeof(Offset)); }
    uint8_t* serialize(uint8_t* buff) const {
        const uint32_t nb = (uint32_t)_bytes.size(), no = (uint32_t)(_offsets.size() * sizeof(Offset));
        memcpy(buff, &nb, 4);
        buff += 4;
        if (nb) memcpy(buff, _bytes.data(), nb);
        buff += nb;
        memcpy(buff, &no, 4);
        buff += 4;
        memcpy(buff, _offsets.data(), no);
        return buff + no;
    }
    // returns the position behind the column, nullptr when the buffer is too short or the offsets are not a valid
    // non-decreasing sequence ending at the byte count
    const uint8_t* deserialize(const uint8_t* buff, const uint8_t* end) {
        uint32_t nb = 0, no = 0;
        if (end - buff < 4) return nullptr;
        memcpy(&nb, buff, 4);
        buff += 4;
        if ((size_t)(end - buff) < nb) return nullptr;
        Bytes bytes(buff, buff + nb);
        buff += nb;
        if (end - buff < 4) return nullptr;
        memcpy(&no, buff, 4);
        buff += 4;
        if ((size_t)(end - buff) < no || no % sizeof(Offset) != 0 || no == 0) return nullptr;
        Offsets offsets(no / sizeof(Offset));
        memcpy(offsets.data(), buff, no);
        buff += no;
        if (offsets.front() != 0 || offsets.back() != nb || !std::is_sorted(offsets.begin(), offsets.end())) return nullptr;
        _bytes.swap(bytes);
        _offsets.swap(offsets);
        _slices_valid = false;
        return buff;
    }

    void check_or_die() const {
        if (_offsets.empty() || _offsets.front() != 0 || _offsets.back() != _bytes.size() || !std::is_sorted(_offsets.begin(), _offsets.end()))
            throw std::runtime_error("BinaryColumn: offsets do not describe the byte array");
    }

private:
    // Column::capacity_limit_reached: a BinaryColumn addresses its bytes with 32-bit offsets
    void check_capacity(size_t more) const {
        if (_bytes.size() + more > 0xFFFFFFFFull) throw std::length_error("BinaryColumn: more than 4 GiB of string bytes in one column");
    }

    Bytes _bytes;
    Offsets _offsets; // size() + 1 entries, _offsets[0] == 0
    mutable Container _slices;
    mutable bool _slices_valid = false;
};

} // namespace starrocks
