"""Serving micro-batcher (host logic, CPU): concurrent single-row requests ride in shared calls, every caller gets its
own row back, grouping by (k, normalize), deadline flush, error propagation, drain on close."""
import threading
import time

import numpy as np
import pytest

import sse_serve


class Recorder(object):
    def __init__(self, delay=0.0, fail_on=None):
        self.batches, self.delay, self.fail_on = [], delay, fail_on

    def __call__(self, tokens, k, normalize):
        self.batches.append((tokens.shape[0], k, normalize))
        if self.delay:
            time.sleep(self.delay)
        if self.fail_on is not None and (tokens[:, 0] == self.fail_on).any():
            raise RuntimeError("boom")
        base = tokens[:, :1].astype(np.float32)                       # a result that identifies the row it came from
        scores = base + np.arange(k, dtype=np.float32)[None, :] * (0.5 if normalize else 1.0)
        ids = (tokens[:, :1] * 100 + np.arange(k)[None, :]).astype(np.int32)
        return scores, ids


def test_concurrent_requests_share_calls_and_get_their_own_rows():
    rec = Recorder(delay=0.01)
    with sse_serve.MicroBatcher(rec, max_batch=16, max_wait_ms=20.0) as mb:
        out = {}

        def client(i):
            s, ids = mb.query(np.full(8, i, np.int32), k=3)
            out[i] = (s, ids)
        th = [threading.Thread(target=client, args=(i,)) for i in range(40)]
        for t in th: t.start()
        for t in th: t.join()
    assert sorted(out) == list(range(40))
    for i, (s, ids) in out.items():
        assert s.tolist() == [i, i + 1, i + 2] and ids.tolist() == [100 * i, 100 * i + 1, 100 * i + 2]
    assert sum(b for b, _k, _n in rec.batches) == 40 and max(b for b, _k, _n in rec.batches) <= 16
    assert len(rec.batches) < 40 and mb.calls == len(rec.batches) and mb.rows == 40      # batching happened


def test_groups_by_call_arguments_and_flushes_on_deadline():
    rec = Recorder()
    with sse_serve.MicroBatcher(rec, max_batch=64, max_wait_ms=30.0) as mb:
        t0 = time.monotonic()
        f1 = mb.submit(np.full(4, 1, np.int32), k=2, normalize=False)
        f2 = mb.submit(np.full(4, 2, np.int32), k=2, normalize=True)
        f3 = mb.submit(np.full(4, 3, np.int32), k=5, normalize=False)
        r1, r2, r3 = f1.result(2), f2.result(2), f3.result(2)
        waited = time.monotonic() - t0
    assert 0.02 < waited < 1.0                                         # a lone request waits max_wait, not forever
    assert sorted(rec.batches) == [(1, 2, False), (1, 2, True), (1, 5, False)]
    assert r1[0].tolist() == [1, 2] and r2[0].tolist() == [2, 2.5] and len(r3[0]) == 5


def test_full_batch_does_not_wait_and_overflow_stays_queued():
    rec = Recorder()
    with sse_serve.MicroBatcher(rec, max_batch=4, max_wait_ms=5000.0) as mb:
        futs = [mb.submit(np.full(4, i, np.int32), k=1) for i in range(9)]
        first = [f.result(2)[1][0] for f in futs[:8]]                  # two full batches are served immediately
        assert first == [100 * i for i in range(8)]
        assert not futs[8].done()                                      # the ninth waits for company (or for close)
    assert futs[8].result(0)[1][0] == 800                              # close() drained it
    assert [b for b, _k, _n in rec.batches] == [4, 4, 1]


def test_failure_reaches_every_waiter_of_that_batch_only():
    rec = Recorder(fail_on=7)
    with sse_serve.MicroBatcher(rec, max_batch=2, max_wait_ms=5.0) as mb:
        a, b = mb.submit(np.full(4, 7, np.int32), k=1), mb.submit(np.full(4, 8, np.int32), k=1)
        with pytest.raises(RuntimeError):
            a.result(2)
        with pytest.raises(RuntimeError):
            b.result(2)
        assert mb.query(np.full(4, 9, np.int32), k=1, timeout=2)[1][0] == 900     # the batcher keeps serving
        c, d = mb.submit(np.full(4, 1, np.int32), k=1), mb.submit(np.full(5, 1, np.int32), k=1)
        with pytest.raises(ValueError):
            c.result(2)                                                # rows of different lengths cannot be stacked
    with pytest.raises(RuntimeError):
        mb.submit(np.zeros(4, np.int32), k=1)                          # closed
