"""N>1 host logic on CPU: two gloo ranks shard a query batch, each answers its shard, rank 0 gathers — the result must
equal the unsharded answer.  The per-rank searcher here is the CPU oracle (no GPU in this test); on the GPU box the
same search_sharded() wraps Index.search_batch (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import _pkg
    from oracle import oracle
    m = _pkg.load()
    from vectorchord_bm25_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    c = oracle.Corpus.synth(77, 3000, 200, 8, 40, 0.8)
    ix = oracle.OracleIndex(c)
    q_off, q_terms = oracle.gen_queries(78, 37, 200, 1, 5, ix.df, 0.8)   # 37: uneven split

    def search_fn(o, t, k):
        d, s, n, _ = ix.search_batch(o, t, k, nthreads=1, wand=False)
        return {"doc": d, "score": s, "n": n}

    got = shard.search_sharded(search_fn, q_off, q_terms, 10, rank, world)
    # the tensor path of the GPU bench (bench.py strong_leg): this rank's rows as byte tensors, ONE dist.gather per
    # array with the shards padded to a common size (37 = 19 + 18 rows), trimmed on rank 0
    import torch
    sub_off, sub_terms, lo, hi = shard.shard_queries(q_off, q_terms, rank, world)
    part = search_fn(sub_off, sub_terms, 10)
    as_bytes = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(hi - lo, -1))
    rows = shard.gather_rows({k: as_bytes(part[k]) for k in ("doc", "score", "n")}, 37, rank, world)
    if rank == 0:
        want = search_fn(q_off, q_terms, 10)
        ok = all(np.array_equal(got[k], want[k]) for k in ("doc", "score", "n")) and got["doc"].shape == (37, 10)
        for k in ("doc", "score", "n"):
            ok = ok and rows[k].numpy().tobytes() == np.ascontiguousarray(want[k]).tobytes()
        open(out_path, "w").write("ok" if ok else "mismatch")
    else:
        assert rows is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    import _pkg
    _pkg.load()
    from vectorchord_bm25_b200 import shard
    for nq in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(nq, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nq
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gloo_sharded_search(tmp_path):
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
