"""Vectorised train-batch sampler (data.Data.get_train_batch_arrays) against the rules of the reference loop
(data.py:95-115, restated in data.Data.get_train_batch): layout, window quirk, positives verified, negatives never
verified, uniform choices."""
import time

import numpy as np

import data as data_mod


def make_data(P=300, N=50, T=6, seed=0):
    rng = np.random.RandomState(seed)
    d = data_mod.Data.__new__(data_mod.Data)
    d.rng = np.random.RandomState(seed + 1)
    d.fullSetTargetIds = ["t%d" % j for j in range(N)]
    d.rawnegSetLen = N
    d.encodedFullTargetSpace = {t: [j] * T for j, t in enumerate(d.fullSetTargetIds)}          # row content = target number
    d.rawTrainPosCorpus = []
    for i in range(P):
        ver = ["t%d" % j for j in sorted(rng.choice(N, size=int(rng.randint(1, 5)), replace=False))]
        d.rawTrainPosCorpus.append(([i] * T, ver))                                             # row content = positive number
    return d


def test_layout_window_and_membership_rules():
    d = make_data()
    B = 32
    seen_pos, seen_neg = set(), set()
    for _ in range(300):
        src, tgt, lab = d.get_train_batch_arrays(B)
        n = lab.shape[0] // 2
        assert src.dtype == np.int32 and tgt.dtype == np.int32 and lab.dtype == np.float32
        assert src.shape == (2 * n, 6) and tgt.shape == (2 * n, 6) and n <= B
        assert lab.tolist() == [1.0, 0.0] * n
        rows = src[0::2, 0]
        assert np.array_equal(src[0::2], src[1::2])                        # the same source row feeds the pair
        assert rows[0] >= B and np.array_equal(rows, np.arange(rows[0], rows[0] + n))      # contiguous window, never in the first B
        for r, p, q in zip(rows.tolist(), tgt[0::2, 0].tolist(), tgt[1::2, 0].tolist()):
            ver = {int(t[1:]) for t in d.rawTrainPosCorpus[r][1]}
            assert p in ver and q not in ver
            seen_pos.add((r, p)); seen_neg.add(q)
    assert len(seen_neg) == 50                                             # every target shows up as a negative
    # every verified target of frequently drawn positives is eventually chosen
    r = 150
    assert {p for (rr, p) in seen_pos if rr == r} == {int(t[1:]) for t in d.rawTrainPosCorpus[r][1]}


def test_negative_choice_is_uniform_over_non_verified_targets():
    d = make_data(P=40, N=10, seed=3)
    counts = np.zeros(10)
    r = 25
    ver = {int(t[1:]) for t in d.rawTrainPosCorpus[r][1]}
    for _ in range(4000):
        src, tgt, _ = d.get_train_batch_arrays(8)
        hit = np.nonzero(src[0::2, 0] == r)[0]
        for h in hit:
            counts[tgt[1::2, 0][h]] += 1
    allowed = [j for j in range(10) if j not in ver]
    assert counts[list(ver)].sum() == 0
    freq = counts[allowed] / counts[allowed].sum()
    assert np.abs(freq - 1.0 / len(allowed)).max() < 0.05


def test_same_rules_as_the_loop_sampler_and_faster():
    d = make_data(P=5000, N=2000, T=50, seed=5)
    t = time.perf_counter()
    for _ in range(20): d.get_train_batch_arrays(512)
    fast = time.perf_counter() - t
    t = time.perf_counter()
    for _ in range(20):
        s, g, l = d.get_train_batch(512)
        np.array(s, np.int32); np.array(g, np.int32)                       # what get_train_feed_dict does with the lists
    slow = time.perf_counter() - t
    print("sampler: arrays %.2f ms/batch, loop %.2f ms/batch" % (fast / 20 * 1e3, slow / 20 * 1e3))
    assert fast < slow
