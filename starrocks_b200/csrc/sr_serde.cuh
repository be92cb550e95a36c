// sr_serde.cuh -- ChunkPB.data wire format of the exchange, encode level 0:
//   ProtobufChunkSerde::serialize_without_meta      be/src/serde/protobuf_serde.cpp:88-140   (version, num_rows, columns)
//   FixedLengthColumnSerde::serialize / deserialize be/src/serde/column_array_serde.cpp:214-255 (fixed32 byte size + raw values)
//   NullableColumnSerde                             be/src/serde/column_array_serde.cpp:759-782 (null column, then data column)
// Column payloads are moved with one byte-exact copy each (cudaMemcpyAsync: the offsets inside the payload are only 4-byte
// aligned at best); the small header words are written by one kernel (device destination) or by the host.
#pragma once

struct sr_serde {
    sr_ctx* ctx = nullptr;
    std::vector<DevBuf> bufs; // 2 per column: data, nulls
    DevBuf staging;           // host payloads are copied to the device once, then split
};

namespace srd {
struct SerdeWord {
    unsigned long long offset;
    uint32_t value;
    uint32_t pad;
};
struct SerdeWords {
    SerdeWord w[2 + 2 * SR_MAX_OUT_COLS];
    int32_t n;
};
// little-endian fixed32 words at arbitrary byte offsets of the payload
__global__ void k_serde_header(uint8_t* __restrict__ dst, const __grid_constant__ SerdeWords words) {
    const int i = threadIdx.x;
    if (i < words.n) {
        uint8_t* p = dst + words.w[i].offset;
        const uint32_t v = words.w[i].value;
        p[0] = (uint8_t)v;
        p[1] = (uint8_t)(v >> 8);
        p[2] = (uint8_t)(v >> 16);
        p[3] = (uint8_t)(v >> 24);
    }
}
} // namespace srd

static int64_t serde_size(const sr_chunk_view* c, int64_t rows) {
    int64_t b = 8;
    for (int k = 0; k < c->num_cols; k++) {
        const int w = srd::type_width(c->cols[k].type);
        if (c->cols[k].nulls) b += 4 + rows;
        b += 4 + rows * w;
    }
    return b;
}

static int32_t serde_serialize(sr_ctx* ctx, const sr_chunk_view* c, int64_t r0, int64_t r1, void* dst, int64_t cap, int32_t dst_mem, sr_chunk_pb_meta* meta) {
    if (c->num_cols < 0 || c->num_cols > SR_MAX_OUT_COLS) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "%d columns", c->num_cols);
    if (r0 < 0 || r1 < r0 || r1 > c->num_rows) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "rows [%lld, %lld) of %lld", (long long)r0, (long long)r1, (long long)c->num_rows);
    const int64_t rows = r1 - r0;
    for (int k = 0; k < c->num_cols; k++) {
        const int w = srd::type_width(c->cols[k].type);
        if (w == 0) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "unknown column type %d (slot %d)", c->cols[k].type, c->cols[k].slot_id);
        if (rows * w >= (1ll << 32)) return sr_fail(ctx, SR_ERR_NOT_SUPPORTED, "a column of %lld bytes: the format stores sizes as uint32", (long long)(rows * w));
    }
    const int64_t total = serde_size(c, rows);
    if (total > cap) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "destination of %lld bytes, %lld needed", (long long)cap, (long long)total);
    if (dst_mem != SR_MEM_HOST && dst_mem != SR_MEM_DEVICE && dst_mem != SR_MEM_HOST_PINNED) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "dst_mem %d", dst_mem);
    const bool dst_dev = dst_mem == SR_MEM_DEVICE, src_dev = c->mem == SR_MEM_DEVICE;
    const cudaMemcpyKind kind = src_dev ? (dst_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost) : (dst_dev ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost);
    uint8_t* out = (uint8_t*)dst;
    srd::SerdeWords words;
    words.n = 0;
    auto word = [&](int64_t off, uint32_t v) {
        words.w[words.n].offset = (unsigned long long)off;
        words.w[words.n].value = v;
        words.n++;
    };
    word(0, 1u);             // version
    word(4, (uint32_t)rows); // num_rows
    int64_t off = 8;
    for (int k = 0; k < c->num_cols; k++) {
        const sr_col_view& col = c->cols[k];
        const int w = srd::type_width(col.type);
        if (col.nulls) {
            word(off, (uint32_t)rows);
            off += 4;
            if (rows) SR_CUDA(ctx, cudaMemcpyAsync(out + off, col.nulls + r0, (size_t)rows, kind, ctx->stream));
            off += rows;
        }
        word(off, (uint32_t)(rows * w));
        off += 4;
        if (rows) SR_CUDA(ctx, cudaMemcpyAsync(out + off, (const uint8_t*)col.data + r0 * w, (size_t)(rows * w), kind, ctx->stream));
        off += rows * w;
        if (meta) {
            meta->slot_ids[k] = col.slot_id;
            meta->types[k] = col.type;
            meta->is_nulls[k] = col.nulls ? 1 : 0;
            meta->is_consts[k] = 0;
        }
    }
    if (dst_dev) {
        srd::k_serde_header<<<1, 2 + 2 * SR_MAX_OUT_COLS, 0, ctx->stream>>>(out, words);
        SR_LAUNCH_CHECK(ctx);
    } else {
        SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the payload copies have landed; the host writes the header words
        for (int i = 0; i < words.n; i++) {
            const uint32_t v = words.w[i].value;
            uint8_t* p = out + words.w[i].offset;
            p[0] = (uint8_t)v, p[1] = (uint8_t)(v >> 8), p[2] = (uint8_t)(v >> 16), p[3] = (uint8_t)(v >> 24);
        }
    }
    if (meta) {
        meta->serialized_size = total;
        meta->num_rows = rows;
        meta->num_cols = c->num_cols;
        meta->reserved = 0;
    }
    return SR_OK;
}

static int32_t serde_deserialize(sr_serde* h, const void* src, int64_t bytes, int32_t src_mem, const sr_chunk_pb_meta* meta, sr_chunk_out* out) {
    sr_ctx* ctx = h->ctx;
    if (meta->num_cols < 0 || meta->num_cols > SR_MAX_OUT_COLS) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "%d columns", meta->num_cols);
    if (bytes < 8) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "payload of %lld bytes has no header", (long long)bytes);
    // the header words and the per-column sizes steer the parse: read them on the host (device payload: the first pass
    // copies the whole payload's size words one by one -- at most 2 per column)
    const bool src_dev = src_mem == SR_MEM_DEVICE;
    const uint8_t* in = (const uint8_t*)src;
    auto read32 = [&](int64_t off, uint32_t* v) -> int32_t {
        if (off + 4 > bytes) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "payload truncated at byte %lld", (long long)off);
        uint8_t b[4];
        if (src_dev) {
            SR_CUDA(ctx, cudaMemcpyAsync(b, in + off, 4, cudaMemcpyDeviceToHost, ctx->stream));
            SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            memcpy(b, in + off, 4);
        }
        *v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
        return SR_OK;
    };
    uint32_t version = 0, rows = 0;
    SR_TRY(read32(0, &version));
    SR_TRY(read32(4, &rows));
    if (version != 1) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "ChunkPB.data version %u (expected 1)", version);
    if (h->bufs.size() < 2 * (size_t)meta->num_cols) {
        std::vector<DevBuf> nb(2 * (size_t)meta->num_cols);
        for (size_t i = 0; i < h->bufs.size(); i++) std::swap(nb[i], h->bufs[i]);
        h->bufs.swap(nb);
    }
    const cudaMemcpyKind kind = src_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int64_t off = 8;
    out->num_cols = meta->num_cols;
    out->mem = SR_MEM_DEVICE;
    out->num_rows = rows;
    for (int k = 0; k < meta->num_cols; k++) {
        const int w = srd::type_width(meta->types[k]);
        if (w == 0) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "unknown column type %d", meta->types[k]);
        out->cols[k].type = meta->types[k];
        out->cols[k].slot_id = meta->slot_ids[k];
        out->cols[k].nulls = nullptr;
        uint32_t sz = 0;
        if (meta->is_nulls[k]) {
            SR_TRY(read32(off, &sz));
            off += 4;
            if (sz != rows || off + sz > bytes) return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "column %d: null column of %u bytes for %u rows", k, sz, rows);
            SR_TRY(h->bufs[2 * k + 1].reserve(ctx, (size_t)std::max<uint32_t>(sz, 1)));
            if (sz) SR_CUDA(ctx, cudaMemcpyAsync(h->bufs[2 * k + 1].p, in + off, sz, kind, ctx->stream));
            out->cols[k].nulls = (uint8_t*)h->bufs[2 * k + 1].p;
            off += sz;
        }
        SR_TRY(read32(off, &sz));
        off += 4;
        if ((int64_t)sz != (int64_t)rows * w || off + sz > bytes)
            return sr_fail(ctx, SR_ERR_INVALID_ARGUMENT, "column %d: %u bytes for %u rows of width %d", k, sz, rows, w);
        SR_TRY(h->bufs[2 * k].reserve(ctx, (size_t)std::max<uint32_t>(sz, 1)));
        if (sz) SR_CUDA(ctx, cudaMemcpyAsync(h->bufs[2 * k].p, in + off, sz, kind, ctx->stream));
        out->cols[k].data = h->bufs[2 * k].p;
        off += sz;
    }
    if (!src_dev) SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the caller's host buffer may go away
    return SR_OK;
}
