"""The substitution-class tables are DERIVED from the genetic code (mcqueen_b200/csrc/mcq_tables.cpp for the
library, mcqueen_b200/genetics.py for the host mirror); the reference stores them as literals
(constants.c:405-688: beta_S, beta_V, alpha_S, alpha_V, S_TS, S_TV, NS_TS, NS_TV and :690 two_step).  Both
derivations are checked here, entry by entry, against the reference's own compiled constants.o
(oracle/_ref/libmcq_ref.so) and against a committed digest of those tables for boxes without oracle/_ref.
CPU only: mcq_codon_class is a host function of the library."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import cpu_libs
from mcqueen_b200 import _lib, build, genetics as gen

# sha256 over the reference's NS_TS, NS_TV, S_TS, S_TV (int32, row-major, in that order), taken from
# oracle/_ref/libmcq_ref.so (constants.c:424-688 compiled unmodified)
REFERENCE_DIGEST = "a7e80947b8d436a6e46960fc34d9d40ee9dfb53f72fcadd489c1cc72f0068e53"


def library_classes():
    build.build()
    lib = _lib.load()
    return np.array([[lib.mcq_codon_class(a, b) for b in range(64)] for a in range(64)], np.int32)


def from_classes(cls):
    return [(cls == k).astype(np.int32) for k in (1, 2, 3, 4)]      # NS_TS, NS_TV, S_TS, S_TV


def reference_tables():
    if not cpu_libs.have_ref():
        pytest.skip("oracle/_ref/libmcq_ref.so not present")
    lib = C.CDLL(cpu_libs.REF_SO)
    mats = {n: np.ctypeslib.as_array((C.c_int * 4096).in_dll(lib, n)).reshape(64, 64).copy()
            for n in ("NS_TS", "NS_TV", "S_TS", "S_TV", "two_step")}
    sums = {n: np.ctypeslib.as_array((C.c_int * 64).in_dll(lib, n)).copy()
            for n in ("alpha_S", "alpha_V", "beta_S", "beta_V")}
    return mats, sums


def test_library_and_host_derivations_agree():
    cls = library_classes()
    ns_ts, ns_tv, s_ts, s_tv = from_classes(cls)
    assert np.array_equal(ns_ts, gen.NS_TS) and np.array_equal(ns_tv, gen.NS_TV)
    assert np.array_equal(s_ts, gen.S_TS) and np.array_equal(s_tv, gen.S_TV)
    # every codon has nine one-step neighbours, each in exactly one class
    assert np.array_equal((cls > 0).sum(1), np.full(64, 9))
    assert np.array_equal(cls > 0, cls.T > 0)


def test_tables_match_reference_constants():
    mats, sums = reference_tables()
    ns_ts, ns_tv, s_ts, s_tv = from_classes(library_classes())
    assert np.array_equal(ns_ts, mats["NS_TS"]) and np.array_equal(ns_tv, mats["NS_TV"])
    assert np.array_equal(s_ts, mats["S_TS"]) and np.array_equal(s_tv, mats["S_TV"])
    assert np.array_equal(gen.TWO_STEP, mats["two_step"])
    # the row sums the emission kernels use as d_cnt (constants.c:405-422)
    assert np.array_equal(ns_ts.sum(1), sums["alpha_S"]) and np.array_equal(ns_tv.sum(1), sums["alpha_V"])
    assert np.array_equal(s_ts.sum(1), sums["beta_S"]) and np.array_equal(s_tv.sum(1), sums["beta_V"])
    assert np.array_equal(gen.ALPHA_S, sums["alpha_S"]) and np.array_equal(gen.BETA_V, sums["beta_V"])


def digest_of(mats):
    h = hashlib.sha256()
    for m in mats:
        h.update(np.ascontiguousarray(m, np.int32).tobytes())
    return h.hexdigest()


def test_derived_tables_match_the_committed_digest():
    """Runs everywhere (no oracle/_ref needed): the digest was taken from the reference's compiled tables."""
    assert digest_of(from_classes(library_classes())) == REFERENCE_DIGEST


def test_reference_tables_match_the_committed_digest():
    mats, _ = reference_tables()
    assert digest_of([mats[n] for n in ("NS_TS", "NS_TV", "S_TS", "S_TV")]) == REFERENCE_DIGEST
