"""Oracle block codec (oracle/bm25_codec.c) against the reference's own test strategy and hand-derived layouts.

The reference tests its codec with random round trips only (crates/simd/src/bitpacking_u32_ordered.rs:239-259,
bitpacking_u32_unordered.rs, bytepacking_u32_*.rs `fn test`) — repeated here — and has no golden byte vectors; the
layout vectors below are derived by hand from the compress! macro (crates/simd/src/bitpacking.rs:14-52)."""
import numpy as np
import pytest

from oracle import oracle as orc


def test_roundtrip_ordered_all_bitwidths():
    # bitpacking_u32_ordered.rs:239-259
    rng = np.random.default_rng(1)
    for i in range(33):
        for rep in range(4):
            hi = (1 << i) if i < 32 else (1 << 32)
            data = np.sort(rng.integers(0, hi, 128, dtype=np.uint64).astype(np.uint32))
            mn = int(data[0])
            meta, payload = orc.compress_document_ids(mn, data)
            assert meta >> 7 == 0 and (meta & 0x7F) <= i
            assert len(payload) == (meta & 0x7F) * 16
            out = orc.decompress_document_ids(mn, meta, payload)
            assert np.array_equal(out, data)


def test_roundtrip_unordered_all_bitwidths():
    rng = np.random.default_rng(2)
    for i in range(33):
        hi = (1 << i) if i < 32 else (1 << 32)
        data = rng.integers(0, hi, 128, dtype=np.uint64).astype(np.uint32)
        meta, payload = orc.compress_term_frequencies(data)
        assert meta >> 7 == 0 and (meta & 0x7F) <= i and len(payload) == (meta & 0x7F) * 16
        assert np.array_equal(orc.decompress_term_frequencies(meta, payload), data)


def test_roundtrip_short_blocks_bytepacking():
    # bytepacking_u32_ordered.rs / _unordered.rs `fn test`: lengths below 128 take 1..4 bytes per value
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 31, 64, 127):
        for bits in (0, 1, 7, 8, 9, 16, 17, 24, 25, 32):
            hi = (1 << bits) if bits < 32 else (1 << 32)
            docs = np.sort(rng.integers(0, max(hi, 1), n, dtype=np.uint64).astype(np.uint32))
            meta, payload = orc.compress_document_ids(int(docs[0]), docs)
            assert meta >> 7 == 1 and 1 <= (meta & 0x7F) <= 4 and len(payload) == n * (meta & 0x7F)
            assert np.array_equal(orc.decompress_document_ids(int(docs[0]), meta, payload), docs)
            tfs = rng.integers(0, max(hi, 1), n, dtype=np.uint64).astype(np.uint32)
            meta, payload = orc.compress_term_frequencies(tfs)
            assert meta >> 7 == 1 and len(payload) == n * (meta & 0x7F)
            assert (meta & 0x7F) == max(1, (int(tfs.max()).bit_length() + 7) // 8)
            assert np.array_equal(orc.decompress_term_frequencies(meta, payload), tfs)


def test_layout_bitwidth1_vertical_lanes():
    # bitwidth 1, unordered: value 4*it+l is bit `it` of 32-bit word l (one output vector of four words).
    tfs = np.zeros(128, dtype=np.uint32)
    tfs[4 * 5 + 2] = 1      # it=5, lane 2
    tfs[4 * 31 + 0] = 1     # it=31, lane 0
    meta, payload = orc.compress_term_frequencies(tfs)
    assert meta == 1 and len(payload) == 16
    words = payload.view("<u4")
    assert list(words) == [1 << 31, 0, 1 << 5, 0]


def test_layout_bitwidth3_carry():
    # bitwidth 3: value `it` of a lane sits at bit 3*it of that lane's stream; it=10 straddles words 0 and 1
    # (cursor 30: two bits in word 0, one carried into word 1 — bitpacking.rs:45-49).
    tfs = np.zeros(128, dtype=np.uint32)
    tfs[4 * 10 + 1] = 0b111
    tfs[0] = 0b101
    meta, payload = orc.compress_term_frequencies(tfs)
    assert meta == 3 and len(payload) == 48
    w = payload.view("<u4").reshape(3, 4)   # [output vector j][lane]
    assert w[0, 0] == 0b101
    assert w[0, 1] == 0b11 << 30 and w[1, 1] == 0b1
    assert w.sum() == 0b101 + (0b11 << 30) + 1


def test_layout_ordered_delta_and_raw32():
    # ordered: deltas against the previous value, the first against min (bitpacking_u32_ordered.rs:82-91)
    docs = np.arange(128, dtype=np.uint32) * 2 + 1000      # deltas: 0, 2, 2, ...
    meta, payload = orc.compress_document_ids(1000, docs)
    assert meta == 2
    w = payload.view("<u4").reshape(2, 4)
    # lane 0 holds deltas of values 0,4,8,...: first is 0 then 2s (0b10 at bits 2*it)
    assert w[0, 0] == int("10" * 15 + "00", 2) and w[0, 1] == int("10" * 16, 2)
    # bitwidth 32 stores the raw values, not the deltas (bitpacking_u32_ordered.rs:119-121)
    docs = np.sort(np.concatenate([[0], np.full(127, 0x80000000, dtype=np.uint64) + np.arange(127)]).astype(np.uint32))
    meta, payload = orc.compress_document_ids(0, docs)
    assert meta == 32 and np.array_equal(payload.view("<u4"), docs)
    assert np.array_equal(orc.decompress_document_ids(0, meta, payload), docs)


def test_layout_bytepacking_little_endian():
    docs = np.array([10, 10 + 0x1234, 10 + 0x1234 + 1], dtype=np.uint32)
    meta, payload = orc.compress_document_ids(10, docs)
    assert meta == 0x82 and list(payload) == [0, 0, 0x34, 0x12, 1, 0]
    meta, payload = orc.compress_term_frequencies(np.array([1, 2, 3], dtype=np.uint32))
    assert meta == 0x81 and list(payload) == [1, 2, 3]
    # all-zero deltas still take one byte each (div_ceil(8).max(1), bytepacking_u32_ordered.rs:29)
    meta, payload = orc.compress_document_ids(7, np.array([7], dtype=np.uint32))
    assert meta == 0x81 and list(payload) == [0]


def test_malformed_blocks_rejected():
    assert orc.decompress_document_ids(0, 33, np.zeros(33 * 16, dtype=np.uint8)) is None      # bitwidth out of bound
    assert orc.decompress_document_ids(0, 5, np.zeros(5 * 16 - 1, dtype=np.uint8)) is None    # unexpected input len
    assert orc.decompress_term_frequencies(0x85, np.zeros(10, dtype=np.uint8)) is None       # bytewidth 5
    assert orc.decompress_term_frequencies(0x82, np.zeros(3, dtype=np.uint8)) is None        # ragged payload


def test_encode_blocks_matches_corpus():
    # flush.rs:78-120: blocks of 128 per term, the last one shorter; decoding gives the CSR back
    c = orc.Corpus.synth(0xB25C0DE0 + 77, 3000, 40, 20, 60, zipf_s=1.0)
    eb = orc.EncodedBlocks(c)
    df = (c.post_off[1:] - c.post_off[:-1]).astype(np.int64)
    assert eb.n_blocks == int(((df + 127) // 128).sum()) and (df > 128).any() and (df % 128 != 0).any()
    for t in range(c.n_terms):
        docs, tfs = [], []
        for b in range(int(eb.term_blk_off[t]), int(eb.term_blk_off[t + 1])):
            md, mt = int(eb.meta_doc[b]), int(eb.meta_tf[b])
            n = int(eb.blk_n[b])
            nd = (md & 0x7F) * 16 if md >> 7 == 0 else (md & 0x7F) * n
            nt = (mt & 0x7F) * 16 if mt >> 7 == 0 else (mt & 0x7F) * n
            assert int(eb.tf_off[b]) == int(eb.doc_off[b]) + nd
            d = orc.decompress_document_ids(int(eb.blk_min[b]), md, eb.bytes[int(eb.doc_off[b]):int(eb.doc_off[b]) + nd])
            f = orc.decompress_term_frequencies(mt, eb.bytes[int(eb.tf_off[b]):int(eb.tf_off[b]) + nt])
            assert len(d) == n == len(f) and (n == 128 or b == int(eb.term_blk_off[t + 1]) - 1)
            docs.append(d)
            tfs.append(f)
        lo, hi = int(c.post_off[t]), int(c.post_off[t + 1])
        if hi > lo:
            assert np.array_equal(np.concatenate(docs), c.post_doc[lo:hi])
            assert np.array_equal(np.concatenate(tfs), c.post_tf[lo:hi])


def test_roundtrip_property_based():
    # hypothesis version of the reference's random round trips: arbitrary lengths, gaps and magnitudes
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 128), st.integers(0, 32), st.integers(0, 2 ** 32 - 1), st.randoms(use_true_random=False))
    def check(n, bits, start, rnd):
        hi = (1 << bits) - 1 if bits else 0
        gaps = [rnd.randint(0, hi) for _ in range(n)]
        docs, cur = [], start
        for g in gaps:                      # ascending ids that stay inside u32
            cur = min(cur + g, 2 ** 32 - 1)
            docs.append(cur)
        docs = np.array(docs, dtype=np.uint32)
        meta, payload = orc.compress_document_ids(int(docs[0]), docs)
        assert (meta >> 7 == 0) == (n == 128)
        assert np.array_equal(orc.decompress_document_ids(int(docs[0]), meta, payload), docs)
        tfs = np.array([rnd.randint(0, hi) for _ in range(n)], dtype=np.uint32)
        meta, payload = orc.compress_term_frequencies(tfs)
        assert np.array_equal(orc.decompress_term_frequencies(meta, payload), tfs)

    check()


def test_layout_bytewidth4_is_raw_not_delta():
    """A token's last block whose largest gap needs 4 bytes stores the ids THEMSELVES, little endian, no delta
    (crates/simd/src/bytepacking_u32_ordered.rs:195 `4 => output.copy_from_slice(as_bytes(input))`, :211 the raw copy
    back; widths 1..3 store deltas, :37-60).  Byte string derived by hand from those two lines."""
    docs = np.array([7, 7 + (1 << 24) + 5, 0x03020100, 0xFFFFFFF0], dtype=np.uint32)
    meta, payload = orc.compress_document_ids(7, docs)
    assert meta == (0x80 | 4) and len(payload) == 16
    want = bytes([7, 0, 0, 0,   12, 0, 0, 1,   0x00, 0x01, 0x02, 0x03,   0xF0, 0xFF, 0xFF, 0xFF])
    assert payload.tobytes() == want
    assert np.array_equal(orc.decompress_document_ids(7, meta, payload), docs)
    # `min` is ignored by the raw form: any seed decodes to the same ids
    assert np.array_equal(orc.decompress_document_ids(123456, meta, payload), docs)
    # one byte narrower: deltas against the previous id, the first against min
    docs3 = np.array([7, 7 + (1 << 24) - 1, 7 + (1 << 24) + 2], dtype=np.uint32)
    meta, payload = orc.compress_document_ids(7, docs3)
    assert meta == (0x80 | 3)
    assert payload.tobytes() == bytes([0, 0, 0,   0xFF, 0xFF, 0xFF,   3, 0, 0])
    assert np.array_equal(orc.decompress_document_ids(7, meta, payload), docs3)
