"""Closest-L selection on the GPU against the CPU oracle: Hamming distances and the selected index
lists are bit-exact, including their order and the number of libc random() draws consumed."""
import numpy as np
import pytest

import cpu_libs
from mcqueen_b200 import synth

pytestmark = pytest.mark.gpu


def checker():
    return cpu_libs.Checker("ref" if cpu_libs.have_ref() else "oracle")


@pytest.fixture(scope="module")
def api():
    from mcqueen_b200 import api as m
    return m


def _data(S=100, N=300, Q=40, seed=5, **kw):
    return synth.make_dataset(seed=seed, S=S, N=N, Q=Q, H=4, ref_unknown_frac=0.1, query_gap=True, **kw)


@pytest.mark.parametrize("S,N,Q", [(100, 300, 40), (7, 130, 33), (43, 1000, 5), (500, 257, 70),
                                   (7000, 40, 9)])      # 21 000 bases: the 8-queries-per-block variant
def test_hamming_batch_bit_exact(api, S, N, Q):
    d = _data(S, N, Q)
    refs, qs = d["ref_seqs"].copy(), d["query_seqs"].copy()
    refs[3, :5] = np.frombuffer(b"acgtn", np.uint8)       # lower case + unknown
    qs[1, :4] = np.frombuffer(b"tg-a", np.uint8)
    chk = checker()
    got = api.hamming_batch(refs, qs)
    for q in range(Q):
        assert np.array_equal(got[q], chk.hamming(refs, qs[q])), q


@pytest.mark.parametrize("L,skip", [(20, False), (20, True), (1, False), (298, False), (64, True)])
def test_closest_batch_bit_exact(api, L, skip):
    d = _data()
    refs, qs = d["ref_seqs"], d["query_seqs"]
    chk = checker()
    n = L + 1 if skip else L
    cpu_libs.srandom(99)
    want = np.zeros((qs.shape[0], L), np.int32)
    for q in range(qs.shape[0]):
        c = chk.closest(chk.hamming(refs, qs[q]), n)
        want[q] = c[1:] if skip else c
    after_cpu = cpu_libs.libc_random()
    cpu_libs.srandom(99)
    got = api.closest_batch(refs, qs, L, skip_first=skip)
    after_gpu = cpu_libs.libc_random()
    assert np.array_equal(got, want)
    assert after_cpu == after_gpu          # same number of draws consumed


def test_engineered_ties(api):
    """All references identical except a few: the tie window spans almost the whole panel."""
    rng = np.random.default_rng(0)
    base = rng.integers(0, 4, 60)
    refs = np.frombuffer(b"TCAG", np.uint8)[np.repeat(base[None], 500, 0)].copy()
    for r in range(0, 500, 50):
        refs[r, :3] = ord("N")
    qs = refs[:7].copy()
    qs[:, 10] = ord("-")
    chk = checker()
    for L in (5, 10, 11, 490):
        cpu_libs.srandom(7)
        want = np.stack([chk.closest(chk.hamming(refs, qs[q]), L) for q in range(7)])
        cpu_libs.srandom(7)
        got = api.closest_batch(refs, qs, L)
        assert np.array_equal(got, want), L


def test_dropin_symbols(api):
    d = _data(S=50, N=400, Q=6)
    refs, qs = d["ref_seqs"], d["query_seqs"]
    chk = checker()
    for q in range(6):
        ham = np.zeros(400, np.int32)
        api.hamming_distance_considering_unknown(refs, qs[q], 400, 150, ham)
        assert np.array_equal(ham, chk.hamming(refs, qs[q]))
        for n in (1, 25, 399):
            cpu_libs.srandom(q + n)
            want = chk.closest(ham, n)
            cpu_libs.srandom(q + n)
            got = np.zeros(n, np.int32)
            api.closest_sequences(400, ham, n, got)
            assert np.array_equal(got, want)
    # arbitrary int distances (negative, large) through the drop-in: same rule
    ham = np.array([5, -3, 70000, 5, 5, -3, 9, 5, 100, 5], np.int32)
    for n in (1, 2, 3, 6, 9):
        cpu_libs.srandom(n)
        want = chk.closest(ham, n)
        cpu_libs.srandom(n)
        got = np.zeros(n, np.int32)
        api.closest_sequences(10, ham, n, got)
        assert np.array_equal(got, want)


def test_consensus_batch_matches_host_rule(api):
    """per-query consensus of the closest references (amino_acids.c:118-141) incl. unknown bases and ties"""
    from mcqueen_b200 import genetics as gen, prep
    d = _data(S=40, N=200, Q=25)
    refs = d["ref_seqs"].copy()
    refs[::7, 3] = ord("-"); refs[5, :] = ord("N"); refs[9, 10:20] = np.frombuffer(b"acgtacgtac", np.uint8)
    rng = np.random.default_rng(1)
    for L in (1, 2, 17, 199):
        closest = np.stack([rng.choice(200, L, replace=False) for _ in range(25)]).astype(np.int32)
        got = api.consensus_batch(refs, closest)
        codes = gen.base_codes(refs)
        want = np.stack([prep.majority_codes(codes[closest[q]]) for q in range(25)])
        assert np.array_equal(got, want), L


@pytest.mark.parametrize("S,N,Q,L", [(40, 200, 25, 17), (40, 200, 25, 199), (171, 700, 9, 100), (500, 300, 12, 1),
                                     (43, 2100, 70, 500)])
def test_consensus_batch_matches_reference_generate_consensus_subset(api, S, N, Q, L):
    """k_consensus against the reference's own generate_consensus_subset (amino_acids.c:118-141, called per
    query at dmodel.c:664-681): unknown bases, lower case, ties, positions where no reference has a base."""
    if not cpu_libs.have_ref():
        pytest.skip("oracle/_ref/libmcq_ref.so not present")
    d = _data(S=S, N=N, Q=Q)
    refs = d["ref_seqs"].copy()
    refs[::7, 3] = ord("-"); refs[5, :] = ord("N"); refs[9, 10:20] = np.frombuffer(b"acgtacgtac", np.uint8)
    refs[:, 7] = ord("N")                                   # nobody has a base here: 'T' (code 0)
    refs[: N // 2, 11] = ord("C"); refs[N // 2:, 11] = ord("A")   # engineered near-ties
    rng = np.random.default_rng(L)
    closest = np.stack([rng.choice(N, L, replace=False) for _ in range(Q)]).astype(np.int32)
    want = cpu_libs.reference_consensus_subset(refs, closest)
    got = api.consensus_batch(refs, closest)
    assert np.array_equal(got, want)


def test_sharded_groups_replayed_in_rank_order_equal_the_single_call(api):
    """Queries sharded over ranks: every rank's block of selection groups, replayed in rank order from one seed,
    gives the lists (and the number of draws) of the single-process call - and of the CPU reference."""
    d = _data(S=100, N=300, Q=41)
    refs, qs = d["ref_seqs"], d["query_seqs"]
    for L, skip in ((20, False), (33, True)):
        cpu_libs.srandom(5)
        want = api.closest_batch(refs, qs, L, skip_first=skip)
        after = cpu_libs.libc_random()
        blocks = [api.closest_groups(refs, np.ascontiguousarray(qs[a:b]), L, skip_first=skip)
                  for a, b in ((0, 13), (13, 14), (14, 41))]
        cpu_libs.srandom(5)
        got = np.concatenate([api.closest_replay(b, L, skip_first=skip) for b in blocks])
        assert np.array_equal(got, want) and cpu_libs.libc_random() == after
    with pytest.raises(api._lib.McqError, match="another closest_n"):
        api.closest_replay(blocks[0], 20)


def test_many_boundary_ties_overflow_the_pool_and_are_redone(api):
    """A panel of near-identical references: the '== T' groups hold almost the whole panel for every query, far
    more than the pool of a chunk is sized for - the chunk is redone with a pool of the exact size."""
    rng = np.random.default_rng(1)
    base = rng.integers(0, 4, 90)
    N, Q = 40000, 96
    refs = np.frombuffer(b"TCAG", np.uint8)[np.repeat(base[None], N, 0)].copy()
    refs[::1000, :2] = ord("N")
    qs = refs[:Q].copy()
    qs[:, 7] = ord("-")
    chk = checker()
    cpu_libs.srandom(3)
    want = np.stack([chk.closest(chk.hamming(refs, qs[q]), 50) for q in range(0, Q, 16)])
    cpu_libs.srandom(3)
    got = api.closest_batch(refs, np.ascontiguousarray(qs[::16]), 50)
    assert np.array_equal(got, want)
    got_all = api.closest_batch(refs, qs, 50)            # 96 x ~40 000 ties > the pool of 2^20 entries
    assert got_all.shape == (Q, 50) and all(len(set(r)) == 50 for r in got_all)
