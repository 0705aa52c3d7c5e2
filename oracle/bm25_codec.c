/*
 * bm25_codec.c — CPU oracle for the reference's posting-block codec.
 *
 * TEST INFRASTRUCTURE ONLY (see bm25_oracle.h): the product decodes blocks on the GPU
 * (vectorchord-bm25_b200/csrc/bm25x_blocks.cu) and never links this file.
 *
 * Restates, in scalar C, what the reference does with SIMD/macro-generated code:
 *   crates/bm25/src/compression.rs:36-136      block level: metadata byte + payload
 *   crates/simd/src/bitpacking.rs:14-98        the compress!/decompress! macros (4-lane vertical layout)
 *   crates/simd/src/bitpacking_u32_ordered.rs  delta (D1) + bit packing of 128 ascending doc ids
 *   crates/simd/src/bitpacking_u32_unordered.rs  bit packing of 128 term frequencies (no delta)
 *   crates/simd/src/bytepacking_u32_ordered.rs / _unordered.rs   short (<128) blocks: 1..4 little-endian bytes each
 *
 * Parity pins: the reference's own tests for these files are random round trips only
 * (bitpacking_u32_ordered.rs:239-259 and siblings) — there are no golden byte vectors.  tests/test_codec.py
 * repeats those round trips and adds byte layouts derived by hand from the macro (bitpacking.rs:27-50).
 */
#include <stdint.h>
#include <string.h>

#include "bm25_oracle.h"

static uint8_t bits_of(uint32_t reduce_or) { /* 1 + ilog2, or 0 (bitpacking_u32_ordered.rs:24-28) */
    uint8_t n = 0;
    while (reduce_or) {
        n++;
        reduce_or >>= 1;
    }
    return n;
}

/* bitpacking_u32_ordered.rs:14-31 / bytepacking_u32_ordered.rs:14-30 (before the div_ceil) */
static uint8_t delta_bits(uint32_t min, const uint32_t *in, uint32_t n) {
    uint32_t last = min, acc = 0;
    for (uint32_t i = 0; i < n; i++) {
        acc |= in[i] - last;
        last = in[i];
    }
    return bits_of(acc);
}

static uint8_t raw_bits(const uint32_t *in, uint32_t n) { /* bitpacking_u32_unordered.rs:16-27 */
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; i++) acc |= in[i];
    return bits_of(acc);
}

static void put_u32(uint8_t *p, uint32_t v) { /* zerocopy::Unalign<u32> on a little-endian target */
    p[0] = (uint8_t)v;
    p[1] = (uint8_t)(v >> 8);
    p[2] = (uint8_t)(v >> 16);
    p[3] = (uint8_t)(v >> 24);
}
static uint32_t get_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* The compress! macro (bitpacking.rs:14-52) with T = 4 x u32: input vector `it` holds values 4*it..4*it+3; lane l of
 * every vector forms an independent bit stream in which value `it` sits at bit it*bw; output vector j holds bits
 * 32j..32j+31 of the four streams. */
static void pack128(uint8_t bw, const uint32_t *vals, uint8_t *out) {
    if (bw == 0) return;
    if (bw == 32) { /* bitpacking_u32_ordered.rs:119-121: the raw values, no delta */
        for (int i = 0; i < 128; i++) put_u32(out + 4 * i, vals[i]);
        return;
    }
    memset(out, 0, (size_t)bw * 16);
    for (uint32_t it = 0; it < 32; it++)
        for (uint32_t l = 0; l < 4; l++) {
            const uint32_t v = vals[4 * it + l];
            const uint32_t bit = it * bw, j = bit >> 5, cur = bit & 31u;
            uint8_t *w0 = out + 4 * (4 * j + l);
            put_u32(w0, get_u32(w0) | (v << cur));
            if (cur + bw > 32) { /* bitpacking.rs:45-49: the carry opens the next word */
                uint8_t *w1 = out + 4 * (4 * (j + 1) + l);
                put_u32(w1, get_u32(w1) | (v >> (32 - cur)));
            }
        }
}

/* The decompress! macro (bitpacking.rs:56-98). */
static void unpack128(uint8_t bw, const uint8_t *in, uint32_t *vals) {
    if (bw == 32) {
        for (int i = 0; i < 128; i++) vals[i] = get_u32(in + 4 * i);
        return;
    }
    const uint32_t mask = bw ? (0xFFFFFFFFu >> (32 - bw)) : 0u;
    for (uint32_t it = 0; it < 32; it++)
        for (uint32_t l = 0; l < 4; l++) {
            if (bw == 0) {
                vals[4 * it + l] = 0;
                continue;
            }
            const uint32_t bit = it * bw, j = bit >> 5, cur = bit & 31u;
            uint32_t v = (get_u32(in + 4 * (4 * j + l)) >> cur) & mask;
            if (cur + bw > 32) v |= (get_u32(in + 4 * (4 * (j + 1) + l)) << (32 - cur)) & mask;
            vals[4 * it + l] = v;
        }
}

/* compression.rs:36-63.  Returns the payload size; *meta = flag << 7 | width. */
uint32_t orc_compress_document_ids(uint32_t min_doc, const uint32_t *docs, uint32_t n, uint8_t *meta, uint8_t *out) {
    if (n == 128) {
        const uint8_t bw = delta_bits(min_doc, docs, 128);
        uint32_t d[128];
        if (bw == 32) {
            memcpy(d, docs, sizeof d);
        } else { /* delta(): v0 - state, v1 - v0, ...  (bitpacking_u32_ordered.rs:82-91) */
            uint32_t last = min_doc;
            for (int i = 0; i < 128; i++) {
                d[i] = docs[i] - last;
                last = docs[i];
            }
        }
        pack128(bw, d, out);
        *meta = bw;
        return (uint32_t)bw * 16u;
    }
    uint8_t by = (uint8_t)((delta_bits(min_doc, docs, n) + 7) / 8);
    if (by == 0) by = 1; /* bytepacking_u32_ordered.rs:29 div_ceil(8).max(1) */
    uint32_t last = min_doc;
    for (uint32_t i = 0; i < n; i++) { /* bytepacking_u32_ordered.rs:37-60: low `by` bytes, little endian */
        /* byte width 4 stores the ids themselves, no delta (bytepacking_u32_ordered.rs:195: `4 => copy_from_slice`) */
        const uint32_t dl = by == 4 ? docs[i] : docs[i] - last;
        for (uint8_t k = 0; k < by; k++) out[(size_t)i * by + k] = (uint8_t)(dl >> (8 * k));
        last = docs[i];
    }
    *meta = (uint8_t)(0x80u | by);
    return n * by;
}

/* compression.rs:65-94.  Returns the number of values, or UINT32_MAX on malformed input. */
uint32_t orc_decompress_document_ids(uint32_t min_doc, uint8_t meta, const uint8_t *in, uint32_t n_bytes,
                                     uint32_t *docs) {
    const uint8_t width = meta & 0x7Fu;
    if ((meta >> 7) == 0) {
        if (width > 32 || n_bytes != (uint32_t)width * 16u) return UINT32_MAX;
        unpack128(width, in, docs);
        if (width != 32) { /* bitpacking_u32_ordered.rs decompress delta: running sum seeded with min */
            uint32_t state = min_doc;
            for (int i = 0; i < 128; i++) {
                state += docs[i];
                docs[i] = state;
            }
        }
        return 128;
    }
    if (width < 1 || width > 4 || n_bytes % width) return UINT32_MAX;
    const uint32_t n = n_bytes / width;
    if (n > 128) return UINT32_MAX;
    uint32_t state = min_doc;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t dl = 0;
        for (uint8_t k = 0; k < width; k++) dl |= (uint32_t)in[(size_t)i * width + k] << (8 * k);
        /* width 4: raw ids, the running sum is not applied (bytepacking_u32_ordered.rs:211) */
        state = width == 4 ? dl : state + dl;
        docs[i] = state;
    }
    return n;
}

/* compression.rs:96-111 */
uint32_t orc_compress_term_frequencies(const uint32_t *tfs, uint32_t n, uint8_t *meta, uint8_t *out) {
    if (n == 128) {
        const uint8_t bw = raw_bits(tfs, 128);
        pack128(bw, tfs, out);
        *meta = bw;
        return (uint32_t)bw * 16u;
    }
    uint8_t by = (uint8_t)((raw_bits(tfs, n) + 7) / 8);
    if (by == 0) by = 1;
    for (uint32_t i = 0; i < n; i++)
        for (uint8_t k = 0; k < by; k++) out[(size_t)i * by + k] = (uint8_t)(tfs[i] >> (8 * k));
    *meta = (uint8_t)(0x80u | by);
    return n * by;
}

/* compression.rs:113-136 */
uint32_t orc_decompress_term_frequencies(uint8_t meta, const uint8_t *in, uint32_t n_bytes, uint32_t *tfs) {
    const uint8_t width = meta & 0x7Fu;
    if ((meta >> 7) == 0) {
        if (width > 32 || n_bytes != (uint32_t)width * 16u) return UINT32_MAX;
        unpack128(width, in, tfs);
        return 128;
    }
    if (width < 1 || width > 4 || n_bytes % width) return UINT32_MAX;
    const uint32_t n = n_bytes / width;
    if (n > 128) return UINT32_MAX;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t v = 0;
        for (uint8_t k = 0; k < width; k++) v |= (uint32_t)in[(size_t)i * width + k] << (8 * k);
        tfs[i] = v;
    }
    return n;
}

/* Block the postings of a whole CSR corpus as flush.rs:78-120 does: per term, blocks of 128 (the last one shorter),
 * min_document_id = the block's first doc.  Two passes: sizes (bytes == NULL) then fill.
 * Per block b: blk_min[b], blk_n[b], meta_doc[b], meta_tf[b], doc_off[b], tf_off[b] (byte offsets into `bytes`).
 * term_blk_off[t] = first block of term t.  Returns the total payload size. */
uint64_t orc_encode_blocks(uint32_t n_terms, const uint64_t *post_off, const uint32_t *post_doc,
                           const uint32_t *post_tf, uint64_t *term_blk_off, uint32_t *blk_min, uint32_t *blk_n,
                           uint8_t *meta_doc, uint8_t *meta_tf, uint64_t *doc_off, uint64_t *tf_off,
                           uint8_t *bytes) {
    uint64_t nb = 0, nbytes = 0;
    uint8_t scratch[512], m;
    for (uint32_t t = 0; t < n_terms; t++) {
        if (term_blk_off) term_blk_off[t] = nb;
        for (uint64_t p = post_off[t]; p < post_off[t + 1]; p += 128) {
            const uint32_t n = (uint32_t)(post_off[t + 1] - p < 128 ? post_off[t + 1] - p : 128);
            uint32_t sz = orc_compress_document_ids(post_doc[p], post_doc + p, n, &m, bytes ? bytes + nbytes : scratch);
            if (bytes) {
                blk_min[nb] = post_doc[p];
                blk_n[nb] = n;
                meta_doc[nb] = m;
                doc_off[nb] = nbytes;
            }
            nbytes += sz;
            sz = orc_compress_term_frequencies(post_tf + p, n, &m, bytes ? bytes + nbytes : scratch);
            if (bytes) {
                meta_tf[nb] = m;
                tf_off[nb] = nbytes;
            }
            nbytes += sz;
            nb++;
        }
    }
    if (term_blk_off) term_blk_off[n_terms] = nb;
    return nbytes;
}
