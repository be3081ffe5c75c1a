"""The sequential half of closest-L on the CPU: mcq_closest_replay is a host function of the product library (the
reference's tie shuffle, closest_seqs.c:57-99, replayed over the selection groups the device hands over).  Here the
groups are built with numpy from the checker's own Hamming distances - per query the references closer than the
distance T of sorted element n, sorted by (distance, index), and those at distance T in index order - and the
replayed lists, their order and the number of libc random() draws are compared with the reference's
closest_sequences (oracle/_ref when present, else the C oracle).  No CUDA call is made."""
import numpy as np
import pytest

import cpu_libs
from mcqueen_b200 import api, build, synth

MAGIC = 0x4D435147


def checker():
    return cpu_libs.Checker("ref" if cpu_libs.have_ref() else "oracle")


def groups_block(dists, n):
    """The flat block of mcq_closest_groups (mcq_closest.cu, kGroupsMagic) from rows of distances."""
    flat = [MAGIC, len(dists), n]
    for d in dists:
        order = np.lexsort((np.arange(d.size), d))          # by (distance, index)
        T = d[order[n]]                                     # closest_seqs.c:85: the element behind the list
        lt = order[d[order] < T]
        eq = np.nonzero(d == T)[0]
        flat += [lt.size, eq.size] + lt.tolist() + eq.tolist()
    return np.array(flat, np.int32)


@pytest.mark.parametrize("L,skip", [(20, False), (20, True), (1, False), (120, False), (64, True)])
def test_replay_of_numpy_groups_matches_closest_sequences(L, skip):
    build.build()
    d = synth.make_dataset(seed=5, S=40, N=130, Q=24, H=4, ref_unknown_frac=0.1, query_gap=True)
    refs, qs = d["ref_seqs"], d["query_seqs"]
    chk = checker()
    n = L + 1 if skip else L
    dists = [chk.hamming(refs, q) for q in qs]
    cpu_libs.srandom(11)
    want = np.stack([chk.closest(h, n)[1 if skip else 0:] for h in dists])      # dmodel.c:362, 380-386
    after = cpu_libs.libc_random()
    cpu_libs.srandom(11)
    got = api.closest_replay(groups_block(dists, n), L, skip_first=skip)
    assert np.array_equal(got, want)
    assert cpu_libs.libc_random() == after                  # the same number of draws


def test_replay_with_heavy_ties_and_blocks_in_rank_order():
    """Near-identical references (almost everything ties at the boundary), and the blocks of three 'ranks' replayed
    in rank order from one seed: the lists of the single pass."""
    build.build()
    rng = np.random.default_rng(2)
    base = rng.integers(0, 4, 60)
    refs = np.frombuffer(b"TCAG", np.uint8)[np.repeat(base[None], 400, 0)].copy()
    refs[::7, 3] = ord("N")
    refs[::11, 9] = ord("A")
    qs = refs[:12].copy()
    qs[:, 5] = ord("-")
    chk = checker()
    dists = [chk.hamming(refs, q) for q in qs]
    n = 25
    cpu_libs.srandom(3)
    want = np.stack([chk.closest(h, n) for h in dists])
    cpu_libs.srandom(3)
    got = np.concatenate([api.closest_replay(groups_block(dists[a:b], n), n) for a, b in ((0, 5), (5, 6), (6, 12))])
    assert np.array_equal(got, want)


def test_replay_rejects_malformed_blocks():
    build.build()
    good = groups_block([np.arange(10, dtype=np.int32)], 3)
    with pytest.raises(api._lib.McqError, match="another closest_n"):
        api.closest_replay(good, 4)
    bad = good.copy()
    bad[0] = 7
    with pytest.raises(api._lib.McqError, match="not a block"):
        api.closest_replay(bad, 3)
    with pytest.raises(api._lib.McqError, match="truncated|inconsistent"):
        api.closest_replay(good[:-2], 3)
