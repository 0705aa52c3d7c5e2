"""ctypes wrapper around oracle/libbm25_oracle.so.

TEST INFRASTRUCTURE ONLY (see oracle/bm25_oracle.h): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbm25_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("bm25_oracle.c", "bm25_codec.c", "bm25_synth.c", "bm25_oracle.h")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


class WandStats(C.Structure):
    _fields_ = [("docs_scored", C.c_uint64), ("blocks_decoded", C.c_uint64),
                ("postings_touched", C.c_uint64), ("pivots", C.c_uint64)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    u8p, u32p, u64p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                             C.POINTER(C.c_uint64), C.POINTER(C.c_double))
    L.orc_fieldnorm_to_length.restype = C.c_uint32
    L.orc_fieldnorm_to_length.argtypes = [C.c_uint8]
    L.orc_length_to_fieldnorm.restype = C.c_uint8
    L.orc_length_to_fieldnorm.argtypes = [C.c_uint32]
    L.orc_idf.restype = C.c_double
    L.orc_idf.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_tf.restype = C.c_double
    L.orc_tf.argtypes = [C.c_uint8, C.c_uint32, C.c_double, C.c_double, C.c_double]
    L.orc_cache_new.restype = None
    L.orc_cache_new.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double, f64p, f64p]
    L.orc_cache_evaluate.restype = C.c_double
    L.orc_cache_evaluate.argtypes = [C.c_double, f64p, C.c_uint8, C.c_uint32]
    L.orc_score_from_f64.restype = C.c_int64
    L.orc_score_from_f64.argtypes = [C.c_double]
    L.orc_score_to_f64.restype = C.c_double
    L.orc_score_to_f64.argtypes = [C.c_int64]
    L.orc_index_build.restype = C.c_void_p
    L.orc_index_build.argtypes = [C.c_uint32, u32p, C.c_uint32, u64p, u32p, u32p, C.c_double, C.c_double]
    L.orc_index_free.restype = None
    L.orc_index_free.argtypes = [C.c_void_p]
    L.orc_index_n_docs.restype = C.c_uint32
    L.orc_index_n_docs.argtypes = [C.c_void_p]
    L.orc_index_avgdl.restype = C.c_double
    L.orc_index_avgdl.argtypes = [C.c_void_p]
    L.orc_index_df.restype = C.c_uint32
    L.orc_index_df.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_index_fieldnorm.restype = C.c_uint8
    L.orc_index_fieldnorm.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_search_exhaustive.restype = C.c_int
    L.orc_search_exhaustive.argtypes = [C.c_void_p, u32p, C.c_int, C.c_int, u8p, u32p, f64p, u32p]
    L.orc_search_wand.restype = C.c_int
    L.orc_search_wand.argtypes = [C.c_void_p, u32p, C.c_int, C.c_int, u8p, u32p, f64p, C.POINTER(WandStats)]
    L.orc_search_wand_batch.restype = None
    L.orc_search_wand_batch.argtypes = [C.c_void_p, C.c_int, u32p, u32p, C.c_int, C.c_int, u32p, f64p, u32p,
                                        C.POINTER(WandStats)]
    L.orc_search_exhaustive_batch.restype = None
    L.orc_search_exhaustive_batch.argtypes = [C.c_void_p, C.c_int, u32p, u32p, C.c_int, C.c_int, u32p, f64p, u32p]
    L.orc_evaluate.restype = C.c_double
    L.orc_evaluate.argtypes = [C.c_void_p, u32p, u32p, C.c_int, u32p, C.c_int]
    L.orc_splitmix64.restype = C.c_uint64
    L.orc_splitmix64.argtypes = [C.c_uint64]
    L.orc_zipf_thresholds.restype = None
    L.orc_zipf_thresholds.argtypes = [C.c_uint32, C.c_double, u64p]
    L.orc_draw.restype = C.c_uint64
    L.orc_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    L.orc_draw_term.restype = C.c_uint32
    L.orc_draw_term.argtypes = [C.c_uint64, C.c_uint32, u64p]
    L.orc_index_block_wand.restype = None
    L.orc_index_block_wand.argtypes = [C.c_void_p, u8p, u32p]
    L.orc_index_n_blocks.restype = C.c_uint64
    L.orc_index_n_blocks.argtypes = [C.c_void_p]
    L.orc_synth_corpus.restype = C.c_int
    L.orc_synth_corpus.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, u32p,
                                   u64p, u32p, u32p, C.c_uint64, u64p]
    L.orc_synth_queries.restype = C.c_int
    L.orc_synth_queries.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, u64p, u32p,
                                    u32p]
    L.orc_synth_doc.restype = C.c_int
    L.orc_synth_doc.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, u32p, u32p]
    L.orc_search_growing.restype = C.c_int
    L.orc_search_growing.argtypes = [C.c_void_p, C.c_uint32, u8p, u8p, u64p, u32p, u32p, u32p, C.c_int, C.c_int, u8p,
                                     u32p, f64p]
    L.orc_compress_document_ids.restype = C.c_uint32
    L.orc_compress_document_ids.argtypes = [C.c_uint32, u32p, C.c_uint32, u8p, u8p]
    L.orc_decompress_document_ids.restype = C.c_uint32
    L.orc_decompress_document_ids.argtypes = [C.c_uint32, C.c_uint8, u8p, C.c_uint32, u32p]
    L.orc_compress_term_frequencies.restype = C.c_uint32
    L.orc_compress_term_frequencies.argtypes = [u32p, C.c_uint32, u8p, u8p]
    L.orc_decompress_term_frequencies.restype = C.c_uint32
    L.orc_decompress_term_frequencies.argtypes = [C.c_uint8, u8p, C.c_uint32, u32p]
    L.orc_encode_blocks.restype = C.c_uint64
    L.orc_encode_blocks.argtypes = [C.c_uint32, u64p, u32p, u32p, u64p, u32p, u32p, u8p, u8p, u64p, u64p, u8p]
    _lib = L
    return L


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Corpus:
    """Term-major CSR corpus on the host (the reference's `Mapping(term, doc, tf)` order,
    crates/bm25/src/segment.rs:23-45) + exact doc lengths."""

    def __init__(self, n_docs, doc_len, n_terms, post_off, post_doc, post_tf, k1=1.2, b=0.75):
        self.n_docs = int(n_docs)
        self.n_terms = int(n_terms)
        self.doc_len = _u32(doc_len)
        self.post_off = np.ascontiguousarray(post_off, dtype=np.uint64)
        self.post_doc = _u32(post_doc)
        self.post_tf = _u32(post_tf)
        self.k1 = float(k1)
        self.b = float(b)

    @staticmethod
    def from_docs(docs, n_terms=None, k1=1.2, b=0.75):
        """docs: list of dict{term: tf} (doc id = position). Mirrors cast_tsvector_to_document
        (src/datatype/tsvector.rs:84-94): sorted keys, length = Σ tf."""
        n_docs = len(docs)
        if n_terms is None:
            n_terms = 1 + max((max(d) for d in docs if d), default=0)
        lists = [[] for _ in range(n_terms)]
        doc_len = np.zeros(n_docs, dtype=np.uint32)
        for i, d in enumerate(docs):
            doc_len[i] = min(sum(d.values()), 0xFFFFFFFF)
            for t, tf in d.items():
                lists[t].append((i, tf))
        off = np.zeros(n_terms + 1, dtype=np.uint64)
        pd, pt = [], []
        for t, l in enumerate(lists):
            off[t + 1] = off[t] + len(l)
            for (i, tf) in l:
                pd.append(i)
                pt.append(tf)
        return Corpus(n_docs, doc_len, n_terms, off, np.array(pd, dtype=np.uint32), np.array(pt, dtype=np.uint32),
                      k1, b)

    @staticmethod
    def synth_bulk(seed, n_docs, vocab, len_min, len_max=None, zipf_s=0.0, nthreads=0, k1=1.2, b=0.75):
        """The same corpus as synth(), generated in C for full-size configs (bm25_synth.c, OpenMP)."""
        L = lib()
        len_max = len_min if len_max is None else len_max
        nthreads = nthreads or len(os.sched_getaffinity(0))
        cap = int(n_docs) * int(max(len_max, 1))
        doc_len = np.zeros(n_docs, dtype=np.uint32)
        off = np.zeros(vocab + 1, dtype=np.uint64)
        pd = np.empty(max(cap, 1), dtype=np.uint32)
        pt = np.empty(max(cap, 1), dtype=np.uint32)
        n = C.c_uint64(0)
        rc = L.orc_synth_corpus(seed, n_docs, vocab, len_min, len_max, zipf_s, nthreads, _p(doc_len, C.c_uint32),
                                _p(off, C.c_uint64), _p(pd, C.c_uint32), _p(pt, C.c_uint32), cap, C.byref(n))
        if rc != 0:
            raise RuntimeError(f"orc_synth_corpus failed ({rc})")
        return Corpus(n_docs, doc_len, vocab, off, pd[:n.value], pt[:n.value], k1, b)

    @staticmethod
    def synth(seed, n_docs, vocab, len_min, len_max=None, zipf_s=0.0, k1=1.2, b=0.75):
        """Oracle-side generator (independent restatement of the synthetic spec)."""
        L = lib()
        len_max = len_min if len_max is None else len_max
        thr = None
        if zipf_s > 0:
            thr = np.zeros(vocab, dtype=np.uint64)
            L.orc_zipf_thresholds(vocab, zipf_s, _p(thr, C.c_uint64))
        terms = np.zeros(len_max, dtype=np.uint32)
        tfs = np.zeros(len_max, dtype=np.uint32)
        ln = C.c_uint32(0)
        doc_len = np.zeros(n_docs, dtype=np.uint32)
        tl, dl, fl = [], [], []
        for d in range(n_docs):
            n = L.orc_synth_doc(seed, d, vocab, len_min, len_max, _p(thr, C.c_uint64) if thr is not None else None,
                                _p(terms, C.c_uint32), _p(tfs, C.c_uint32), C.byref(ln))
            doc_len[d] = ln.value
            tl.append(terms[:n].copy())
            fl.append(tfs[:n].copy())
            dl.append(np.full(n, d, dtype=np.uint32))
        t = np.concatenate(tl) if tl else np.zeros(0, np.uint32)
        f = np.concatenate(fl) if fl else np.zeros(0, np.uint32)
        dd = np.concatenate(dl) if dl else np.zeros(0, np.uint32)
        order = np.argsort(t.astype(np.uint64) * np.uint64(1 << 32) + dd.astype(np.uint64), kind="stable")
        t, f, dd = t[order], f[order], dd[order]
        off = np.zeros(vocab + 1, dtype=np.uint64)
        np.cumsum(np.bincount(t, minlength=vocab), out=off[1:])
        return Corpus(n_docs, doc_len, vocab, off, dd, f, k1, b)


class OracleIndex:
    def __init__(self, corpus: Corpus):
        self.c = corpus  # keeps the borrowed arrays alive
        L = lib()
        self.h = L.orc_index_build(corpus.n_docs, _p(corpus.doc_len, C.c_uint32), corpus.n_terms,
                                   _p(corpus.post_off, C.c_uint64), _p(corpus.post_doc, C.c_uint32),
                                   _p(corpus.post_tf, C.c_uint32), corpus.k1, corpus.b)
        if not self.h:
            raise ValueError("orc_index_build: invalid corpus")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h)
            self.h = None

    def block_wand(self):
        """(wand_fieldnorm u8[], wand_term_frequency u32[]) of every block, (token, block) order (flush.rs:101-120)."""
        n = int(lib().orc_index_n_blocks(self.h))
        fn = np.zeros(max(n, 1), dtype=np.uint8)
        tf = np.zeros(max(n, 1), dtype=np.uint32)
        lib().orc_index_block_wand(self.h, _p(fn, C.c_uint8), _p(tf, C.c_uint32))
        return fn[:n], tf[:n]

    @property
    def avgdl(self):
        return lib().orc_index_avgdl(self.h)

    def df(self, t):
        return lib().orc_index_df(self.h, t)

    def search_exhaustive(self, terms, k, allow=None):
        terms = _u32(terms)
        od = np.zeros(k, dtype=np.uint32)
        os_ = np.zeros(k, dtype=np.float64)
        tg = C.c_uint32(0)
        al = np.ascontiguousarray(allow, dtype=np.uint8) if allow is not None else None
        n = lib().orc_search_exhaustive(self.h, _p(terms, C.c_uint32), len(terms), k, _p(al, C.c_uint8),
                                        _p(od, C.c_uint32), _p(os_, C.c_double), C.byref(tg))
        return od[:n], os_[:n], tg.value

    def search_wand(self, terms, k, allow=None, stats=None):
        terms = _u32(terms)
        od = np.zeros(k, dtype=np.uint32)
        os_ = np.zeros(k, dtype=np.float64)
        al = np.ascontiguousarray(allow, dtype=np.uint8) if allow is not None else None
        n = lib().orc_search_wand(self.h, _p(terms, C.c_uint32), len(terms), k, _p(al, C.c_uint8),
                                  _p(od, C.c_uint32), _p(os_, C.c_double),
                                  C.byref(stats) if stats is not None else None)
        return od[:n], os_[:n]

    def search_batch(self, q_off, q_terms, k, nthreads=1, wand=True):
        q_off = _u32(q_off)
        q_terms = _u32(q_terms)
        nq = len(q_off) - 1
        od = np.zeros((nq, k), dtype=np.uint32)
        os_ = np.zeros((nq, k), dtype=np.float64)
        on = np.zeros(nq, dtype=np.uint32)
        st = WandStats()
        if wand:
            lib().orc_search_wand_batch(self.h, nq, _p(q_off, C.c_uint32), _p(q_terms, C.c_uint32), k, nthreads,
                                        _p(od, C.c_uint32), _p(os_, C.c_double), _p(on, C.c_uint32), C.byref(st))
        else:
            lib().orc_search_exhaustive_batch(self.h, nq, _p(q_off, C.c_uint32), _p(q_terms, C.c_uint32), k,
                                              nthreads, _p(od, C.c_uint32), _p(os_, C.c_double), _p(on, C.c_uint32))
        return od, os_, on, st

    def search_growing(self, g: "GrowingDocs", terms, k, allow=None):
        """The growing-segment scan of bm25::search (search.rs:83-135) against this sealed index: exhaustive,
        canonical order; ids are growing ordinals."""
        terms = _u32(terms)
        out_d = np.zeros(max(k, 1), dtype=np.uint32)
        out_s = np.zeros(max(k, 1), dtype=np.float64)
        al = np.ascontiguousarray(allow, dtype=np.uint8) if allow is not None else None
        n = lib().orc_search_growing(self.h, g.n_docs, _p(g.fieldnorm, C.c_uint8), _p(g.deleted, C.c_uint8),
                                     _p(g.elem_off, C.c_uint64), _p(g.elem_term, C.c_uint32),
                                     _p(g.elem_tf, C.c_uint32), _p(terms, C.c_uint32), len(terms), int(k),
                                     _p(al, C.c_uint8), _p(out_d, C.c_uint32), _p(out_s, C.c_double))
        return out_d[:n].copy(), out_s[:n].copy()

    def evaluate(self, doc_terms, doc_tfs, query_terms):
        dt, df_, qt = _u32(doc_terms), _u32(doc_tfs), _u32(query_terms)
        return lib().orc_evaluate(self.h, _p(dt, C.c_uint32), _p(df_, C.c_uint32), len(dt), _p(qt, C.c_uint32),
                                  len(qt))


def gen_queries(seed, nq, vocab, nterms_min, nterms_max, df_of, zipf_s=0.0):
    """Queries per SURVEY §8d: draw nterms distinct terms with df > 0 from the corpus distribution."""
    L = lib()
    thr = None
    if zipf_s > 0:
        thr = np.zeros(vocab, dtype=np.uint64)
        L.orc_zipf_thresholds(vocab, zipf_s, _p(thr, C.c_uint64))
    tp = _p(thr, C.c_uint64) if thr is not None else None
    off = [0]
    out = []
    for i in range(nq):
        m = nterms_min
        if nterms_max > nterms_min:
            u = L.orc_draw(seed, i, 0xFFFFFFFF)
            m = nterms_min + (((u >> 32) * (nterms_max - nterms_min + 1)) >> 32)
        got = []
        j = 0
        while len(got) < m and j < 64 * m + 64:
            t = L.orc_draw_term(L.orc_draw(seed, i, j), vocab, tp)
            j += 1
            if t in got or df_of(t) == 0:
                continue
            got.append(t)
        got.sort()
        out.extend(got)
        off.append(len(out))
    return np.array(off, dtype=np.uint32), np.array(out, dtype=np.uint32)


def gen_queries_bulk(seed, nq, vocab, nterms_min, nterms_max, post_off, zipf_s=0.0):
    """gen_queries() in C (bm25_synth.c) for full-size batches."""
    q_off = np.zeros(nq + 1, dtype=np.uint32)
    q_terms = np.zeros(max(nq * nterms_max, 1), dtype=np.uint32)
    post_off = np.ascontiguousarray(post_off, dtype=np.uint64)
    rc = lib().orc_synth_queries(seed, nq, vocab, nterms_min, nterms_max, zipf_s, _p(post_off, C.c_uint64),
                                 _p(q_off, C.c_uint32), _p(q_terms, C.c_uint32))
    if rc != 0:
        raise RuntimeError(f"orc_synth_queries failed ({rc})")
    return q_off, q_terms[:int(q_off[-1])].copy()


# ---- posting-block codec (bm25_codec.c; compression.rs:36-136) ----

def compress_document_ids(min_doc, docs):
    """-> (metadata byte, payload bytes)   (compression.rs:36-63)"""
    docs = np.ascontiguousarray(docs, dtype=np.uint32)
    out = np.zeros(512, dtype=np.uint8)
    meta = C.c_uint8(0)
    n = lib().orc_compress_document_ids(int(min_doc), _p(docs, C.c_uint32), len(docs), C.byref(meta), _p(out, C.c_uint8))
    return meta.value, out[:n].copy()


def decompress_document_ids(min_doc, meta, payload):
    """-> doc ids, or None when the block is malformed   (compression.rs:65-94)"""
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    out = np.zeros(128, dtype=np.uint32)
    n = lib().orc_decompress_document_ids(int(min_doc), int(meta), _p(payload, C.c_uint8), len(payload), _p(out, C.c_uint32))
    return None if n == 0xFFFFFFFF else out[:n].copy()


def compress_term_frequencies(tfs):
    tfs = np.ascontiguousarray(tfs, dtype=np.uint32)
    out = np.zeros(512, dtype=np.uint8)
    meta = C.c_uint8(0)
    n = lib().orc_compress_term_frequencies(_p(tfs, C.c_uint32), len(tfs), C.byref(meta), _p(out, C.c_uint8))
    return meta.value, out[:n].copy()


def decompress_term_frequencies(meta, payload):
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    out = np.zeros(128, dtype=np.uint32)
    n = lib().orc_decompress_term_frequencies(int(meta), _p(payload, C.c_uint8), len(payload), _p(out, C.c_uint32))
    return None if n == 0xFFFFFFFF else out[:n].copy()


class EncodedBlocks:
    """A corpus' postings as the reference's sealed segment stores them: blocks of 128 in the block codec
    (flush.rs:78-120).  Arrays are what bm25x_index_create_from_blocks takes."""

    def __init__(self, corpus: "Corpus"):
        L = lib()
        T = corpus.n_terms
        off = corpus.post_off
        df = (off[1:] - off[:-1]).astype(np.uint64)
        nb = int(((df + 127) // 128).sum())
        self.corpus = corpus
        self.term_blk_off = np.zeros(T + 1, dtype=np.uint64)
        args = (T, _p(off, C.c_uint64), _p(corpus.post_doc, C.c_uint32), _p(corpus.post_tf, C.c_uint32))
        n_bytes = L.orc_encode_blocks(*args, _p(self.term_blk_off, C.c_uint64), None, None, None, None, None, None, None)
        self.blk_min = np.zeros(max(nb, 1), dtype=np.uint32)
        self.blk_n = np.zeros(max(nb, 1), dtype=np.uint32)
        self.meta_doc = np.zeros(max(nb, 1), dtype=np.uint8)
        self.meta_tf = np.zeros(max(nb, 1), dtype=np.uint8)
        self.doc_off = np.zeros(max(nb, 1), dtype=np.uint64)
        self.tf_off = np.zeros(max(nb, 1), dtype=np.uint64)
        self.bytes = np.zeros(max(int(n_bytes), 1), dtype=np.uint8)
        got = L.orc_encode_blocks(*args, _p(self.term_blk_off, C.c_uint64), _p(self.blk_min, C.c_uint32), _p(self.blk_n, C.c_uint32),
                                  _p(self.meta_doc, C.c_uint8), _p(self.meta_tf, C.c_uint8),
                                  _p(self.doc_off, C.c_uint64), _p(self.tf_off, C.c_uint64), _p(self.bytes, C.c_uint8))
        assert got == n_bytes and int(self.term_blk_off[T]) == nb
        self.n_blocks = nb
        self.n_bytes = int(n_bytes)


class GrowingDocs:
    """Documents inserted since the last seal (the VectorTuple chain of search.rs:83-135), doc-major: document g holds
    (term ordinal, tf) elements elem_off[g]..elem_off[g+1], ascending terms; norms quantised like flush.rs:58."""

    def __init__(self, elem_off, elem_term, elem_tf, doc_len, deleted=None):
        self.elem_off = np.ascontiguousarray(elem_off, dtype=np.uint64)
        self.elem_term = _u32(elem_term)
        self.elem_tf = _u32(elem_tf)
        self.doc_len = _u32(doc_len)
        self.n_docs = len(self.elem_off) - 1
        self.fieldnorm = np.array([lib().orc_length_to_fieldnorm(int(x)) for x in self.doc_len], dtype=np.uint8)
        self.deleted = np.ascontiguousarray(deleted, dtype=np.uint8) if deleted is not None else None

    @staticmethod
    def from_corpus(c: "Corpus", deleted=None):
        """Transpose a term-major corpus into its documents."""
        df = (c.post_off[1:] - c.post_off[:-1]).astype(np.int64)
        term = np.repeat(np.arange(c.n_terms, dtype=np.uint32), df)
        order = np.lexsort((term, c.post_doc))
        cnt = np.bincount(c.post_doc, minlength=c.n_docs).astype(np.uint64)
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
        return GrowingDocs(off, term[order], c.post_tf[order], c.doc_len, deleted)
