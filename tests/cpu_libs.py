"""ctypes access, for TESTS ONLY, to the two CPU checkers:

* ``oracle/liboracle.so``      - our C restatement (oracle/mcq_oracle.c), always buildable;
* ``oracle/_ref/libmcq_ref.so`` - the reference's own translation units compiled unmodified
  (oracle/Makefile), present in the build container and shipped prebuilt to the GPU box.

Both are wrapped behind the same small Python API (``Checker``) so a test can run the
same inputs through either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libmcq_ref.so")
REF_DMODEL = os.path.join(ORACLE_DIR, "_ref", "dmodel_ref")

_libc = C.CDLL("libc.so.6")
_libc.srandom.argtypes = [C.c_uint]
_libc.random.restype = C.c_long


def srandom(seed):
    _libc.srandom(seed)


def libc_random():
    return _libc.random()


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return _p(a, C.c_double)


def ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return _p(a, C.c_int)


def cp(a):
    assert a.dtype in (np.int8, np.uint8) and a.flags.c_contiguous
    return _p(a, C.c_char)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "mcq_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])
    return ORACLE_SO


def have_ref():
    return os.path.exists(REF_SO)


class Checker:
    """Common face of the C oracle (kind='oracle') and the compiled reference (kind='ref')."""

    def __init__(self, kind):
        self.kind = kind
        if kind == "oracle":
            self.lib = C.CDLL(build_oracle())
            self.lib.orc_site_llk.restype = C.c_double
        elif kind == "ref":
            self.lib = C.CDLL(REF_SO)
        else:
            raise ValueError(kind)

    # ---- forward / backward (reference layouts: F[L][S], c2c[64][S]) -------------------
    def forward(self, p, q, c2c_q, R, F_in, rn_in, rn_out, F_out, start, end):
        S, L, N = p.S, p.L, p.N
        cl = np.ascontiguousarray(p.closest_refs[q])
        cons = np.ascontiguousarray(p.cons_query_nucls[q])
        if self.kind == "oracle":
            self.lib.orc_forward(S, L, N, dp(c2c_q), cp(p.ref_triplets), dp(R), dp(F_in), ip(rn_in),
                                 ip(rn_out), dp(F_out), start, end, ip(cl), cp(cons))
        else:
            self.lib.update_F_numerical_recipes(
                S, L, N, dp(c2c_q), cp(p.ref_codons), cp(p.ref_triplets), dp(R), dp(F_in), ip(rn_in),
                ip(rn_out), dp(F_out), start, end, ip(cl), 3 * S, cp(cons))

    def init_forward(self, p, q, c2c_q, R, F, rn):
        """initialise_F_numerical_recipes(0..S-1) (forward_backward.c:11-76)."""
        S, L, N = p.S, p.L, p.N
        if self.kind == "oracle":
            self.forward(p, q, c2c_q, R, F, rn, rn, F, 0, S - 1)
        else:
            cl = np.ascontiguousarray(p.closest_refs[q])
            cons = np.ascontiguousarray(p.cons_query_nucls[q])
            self.lib.initialise_F_numerical_recipes(
                S, L, N, dp(c2c_q), cp(p.ref_codons), cp(p.ref_triplets), dp(R), dp(F), ip(rn),
                0, S - 1, ip(cl), 3 * S, cp(cons))

    def backward(self, p, q, c2c_q, R, B, bn, update):
        S, L, N = p.S, p.L, p.N
        cl = np.ascontiguousarray(p.closest_refs[q])
        cons = np.ascontiguousarray(p.cons_query_nucls[q])
        if self.kind == "oracle":
            self.lib.orc_backward(S, L, N, dp(c2c_q), cp(p.ref_triplets), dp(R), dp(B), ip(bn), update,
                                  ip(cl), cp(cons))
        else:
            self.lib.update_B_numerical_recipes(
                S, L, N, dp(c2c_q), cp(p.ref_codons), cp(p.ref_triplets), dp(R), dp(B), ip(bn), update,
                ip(cl), 3 * S, cp(cons))

    def site_llk(self, p, F, Fn, B, Bn, site):
        if self.kind == "oracle":
            return self.lib.orc_site_llk(p.S, p.L, dp(F), ip(Fn), dp(B) if B is not None else None,
                                         ip(Bn) if Bn is not None else None, site)
        # the reference has no function for this; restate dmodel.c:899-901 / mcmc_moves.c:220-223
        import math
        dot = 0.0
        if B is None:
            for r in range(p.L):
                dot += float(F[r, site])
            return math.log(dot) - int(Fn[site]) * math.log(1e6)
        for r in range(p.L):
            dot += float(F[r, site]) * float(B[r, site])
        return math.log(dot) - (int(Fn[site]) + int(Bn[site])) * math.log(1e6)

    # ---- emissions ------------------------------------------------------------------------
    def emission(self, p, c2c, A, start, end, which_hla, mode, omega=None, esc=None, rev=None,
                 mu=None, rate_out=None):
        omega = p.omega if omega is None else omega
        esc = p.esc if esc is None else esc
        rev = p.rev if rev is None else rev
        mu = p.mu if mu is None else mu
        rate_out = p.rate_out_of_codons(omega) if rate_out is None else rate_out
        S, N, Q, H = p.S, p.N, p.Q, p.H
        if self.kind == "oracle":
            self.lib.orc_emission(S, N, Q, H, ip(p.n_site_codons), cp(p.site_codons), dp(c2c), dp(rate_out),
                                  start, end, which_hla, C.c_double(mu), dp(esc), dp(rev), cp(p.hla),
                                  cp(p.consensus_codons), dp(omega), cp(p.query_codons), dp(A), mode,
                                  C.c_double(p.kappa), C.c_double(p.phi), dp(p.prior))
            return
        lib = self.lib
        C.c_double.in_dll(lib, "kappa").value = p.kappa
        C.c_double.in_dll(lib, "phi").value = p.phi
        prior_c1 = (C.c_double * 64).in_dll(lib, "prior_C1")
        for i in range(64):
            prior_c1[i] = p.prior[i]
        # hla_types is char*[Q] of NUL-terminated '0'/'1' strings (load_data.c:125-157)
        rows = [C.create_string_buffer(bytes(p.hla[q].tobytes())) for q in range(Q)]
        arr = (C.c_char_p * Q)(*[C.cast(r, C.c_char_p) for r in rows])
        lib.create_codon_to_codon_HLA(
            S, N, Q, H, ip(p.n_site_codons), cp(p.site_codons), dp(c2c), dp(rate_out), start, end,
            which_hla, C.c_double(mu), dp(esc), dp(rev), arr, cp(p.consensus_codons), dp(omega),
            ip(p.ns_ts_sum), ip(p.ns_tv_sum), cp(p.type_of_change), cp(p.query_codons), dp(A), mode)

    # ---- closest-L --------------------------------------------------------------------------
    def hamming(self, ref_seqs, query):
        N, nb = ref_seqs.shape
        out = np.zeros(N, np.int32)
        # char *rows[N]: the row addresses as one numpy array (a Python-level cast per row costs 0.1 s at N = 1e5)
        assert ref_seqs.flags.c_contiguous
        ptrs = (ref_seqs.ctypes.data + np.arange(N, dtype=np.uint64) * np.uint64(ref_seqs.strides[0])).astype(np.uint64)
        rows = ptrs.ctypes.data_as(C.POINTER(C.c_char_p))
        q = np.ascontiguousarray(query)
        fn = self.lib.orc_hamming if self.kind == "oracle" else self.lib.hamming_distance_considering_unknown
        fn(rows, cp(q), N, nb, ip(out))
        return out

    def closest(self, hamming, n):
        out = np.zeros(n, np.int32)
        fn = self.lib.orc_closest if self.kind == "oracle" else self.lib.closest_sequences
        fn(len(hamming), ip(np.ascontiguousarray(hamming, np.int32)), n, ip(out))
        return out


def full_state(chk, p, c2c=None, A=None):
    """Emission table + F, B, rnorms and per-query llk for every query (dmodel.c:873-902)."""
    Q, S, L = p.Q, p.S, p.L
    if c2c is None:
        c2c = np.zeros((Q, 64, S))
        A = np.zeros((Q, 64, S))
        chk.emission(p, c2c, A, 0, S - 1, -1, 2)
    F = np.zeros((Q, L, S)); B = np.zeros((Q, L, S))
    Fn = np.zeros((Q, S), np.int32); Bn = np.zeros((Q, S), np.int32)
    llk = np.zeros(Q)
    for q in range(Q):
        chk.init_forward(p, q, c2c[q], p.R, F[q], Fn[q])
        chk.backward(p, q, c2c[q], p.R, B[q], Bn[q], S - 1)
        llk[q] = chk.site_llk(p, F[q], Fn[q], None, None, S - 1)
    return dict(c2c=c2c, A=A, F=F, B=B, Fn=Fn, Bn=Bn, llk=llk)


def reference_consensus_subset(refs, closest):
    """generate_consensus_subset (amino_acids.c:118-141) of the compiled reference, as base codes."""
    from mcqueen_b200 import genetics as gen
    lib = C.CDLL(REF_SO)
    N, nb = refs.shape
    rows = [C.create_string_buffer(bytes(refs[r]), nb + 1) for r in range(N)]
    seqs = (C.c_char_p * N)(*[C.cast(b, C.c_char_p) for b in rows])
    out = np.zeros((closest.shape[0], nb), np.int8)
    buf = C.create_string_buffer(nb + 1)
    for q in range(closest.shape[0]):
        cl = np.ascontiguousarray(closest[q], np.int32)
        lib.generate_consensus_subset(seqs, C.c_int(cl.size), C.c_int(nb), buf, cl.ctypes.data_as(C.c_void_p))
        out[q] = gen.base_codes(np.frombuffer(buf.raw[:nb], np.uint8))
    return out
