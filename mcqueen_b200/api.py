"""Host-side mirror of the reference's interface for the likelihood path, on top of the C ABI.

Two layers, matching include/mcq_b200.h:

* the five functions of forward_backward.h / closest_seqs.h with the reference's names,
  argument order and in-place semantics (numpy arrays stand in for the C arrays);
* ``Context`` - the batched, device-resident API the MCMC loops collapse onto.

Nothing here computes: every call goes to libmcq_b200.so and raises if it is unavailable.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import McqDims, McqParams, McqProposal, check

PROP_REC, PROP_SEL, PROP_ESC, PROP_REV, PROP_MU = 0, 1, 2, 3, 4
DO_HLA_ESC, DO_HLA_REV, DO_HLA_BOTH, DO_ALL_HLAS = 0, 1, 2, -1   # mcmc_moves.h:6-9
BUF_F, BUF_B, BUF_TMP_F, BUF_F_KEPT, BUF_B_KEPT = 0, 1, 2, 3, 4
TAB_C2C, TAB_A, TAB_TMP_C2C, TAB_TMP_A = 0, 1, 2, 3
OPT_SPARSE_ESC = 1
OPT_LAZY_REFILL = 2
OPT_SHARED_PANEL = 3
OPT_ASYNC_UPLOAD = 4
OPT_UPLOAD_CHUNK_QUERIES = 5
OPT_EXCHANGE = 6
OPT_WARPS_PER_QUERY = 7
OPT_PIPELINE_MAX_QUERIES = 8
OPT_CHECKPOINT_STRIDE = 9
K_GATHER, K_EMISSION, K_FORWARD, K_BACKWARD, K_COMMIT, K_SUM = range(6)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _req(a, dtype, name):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous:
        raise TypeError(f"{name} must be a C-contiguous numpy array of {np.dtype(dtype).name}")
    return a


def _chars(a, name):
    if not isinstance(a, np.ndarray) or a.dtype not in (np.int8, np.uint8) or not a.flags.c_contiguous:
        raise TypeError(f"{name} must be a C-contiguous int8/uint8 numpy array")
    return a


# --------------------------------------------------------------------------------------------
# the reference's five functions (same names, same argument order, results written in place)
# --------------------------------------------------------------------------------------------
def initialise_F_numerical_recipes(num_sites, closest_n, total_num_refs, codon_to_codon_HLA, ref_codons,
                                   ref_triplets, R, F, F_rnorm, start_site, end_site, closest_refs,
                                   num_bases, cons_query_nucls):
    """forward_backward.h:6-16.  F is [closest_n][num_sites], filled in place."""
    lib = _lib.load()
    lib.initialise_F_numerical_recipes(
        num_sites, closest_n, total_num_refs, _vp(_req(codon_to_codon_HLA, np.float64, "codon_to_codon_HLA")),
        _vp(_chars(ref_codons, "ref_codons")), _vp(_chars(ref_triplets, "ref_triplets")),
        _vp(_req(R, np.float64, "R")), _vp(_req(F, np.float64, "F")), _vp(_req(F_rnorm, np.int32, "F_rnorm")),
        start_site, end_site, _vp(_req(closest_refs, np.int32, "closest_refs")), num_bases,
        _vp(_chars(cons_query_nucls, "cons_query_nucls")))


def update_F_numerical_recipes(num_sites, closest_n, total_num_refs, codon_to_codon_HLA, ref_codons,
                               ref_triplets, R, F_in, F_rnorm_in, F_rnorm_out, F_out, start_site, end_site,
                               closest_refs, num_bases, cons_query_nucls):
    """forward_backward.h:18-30."""
    lib = _lib.load()
    lib.update_F_numerical_recipes(
        num_sites, closest_n, total_num_refs, _vp(_req(codon_to_codon_HLA, np.float64, "codon_to_codon_HLA")),
        _vp(_chars(ref_codons, "ref_codons")), _vp(_chars(ref_triplets, "ref_triplets")),
        _vp(_req(R, np.float64, "R")), _vp(_req(F_in, np.float64, "F_in")),
        _vp(_req(F_rnorm_in, np.int32, "F_rnorm_in")), _vp(_req(F_rnorm_out, np.int32, "F_rnorm_out")),
        _vp(_req(F_out, np.float64, "F_out")), start_site, end_site,
        _vp(_req(closest_refs, np.int32, "closest_refs")), num_bases,
        _vp(_chars(cons_query_nucls, "cons_query_nucls")))


def update_B_numerical_recipes(num_sites, closest_n, total_num_refs, codon_to_codon_HLA, ref_codons,
                               ref_triplets, R, B, B_rnorm, update, closest_refs, num_bases,
                               cons_query_nucls):
    """forward_backward.h:32-42."""
    lib = _lib.load()
    lib.update_B_numerical_recipes(
        num_sites, closest_n, total_num_refs, _vp(_req(codon_to_codon_HLA, np.float64, "codon_to_codon_HLA")),
        _vp(_chars(ref_codons, "ref_codons")), _vp(_chars(ref_triplets, "ref_triplets")),
        _vp(_req(R, np.float64, "R")), _vp(_req(B, np.float64, "B")), _vp(_req(B_rnorm, np.int32, "B_rnorm")),
        update, _vp(_req(closest_refs, np.int32, "closest_refs")), num_bases,
        _vp(_chars(cons_query_nucls, "cons_query_nucls")))


def hamming_distance_considering_unknown(ref_seqs, query_seqs, num_refs, num_bases, hamming):
    """closest_seqs.h:4-6.  ref_seqs: (num_refs, num_bases) u8 ASCII; hamming: int32[num_refs] out."""
    lib = _lib.load()
    ref_seqs = _chars(ref_seqs, "ref_seqs")
    rows = (C.c_char_p * num_refs)(*[C.cast(ref_seqs[r].ctypes.data, C.c_char_p) for r in range(num_refs)])
    lib.hamming_distance_considering_unknown(C.cast(rows, C.c_void_p), _vp(_chars(query_seqs, "query_seqs")),
                                             num_refs, num_bases, _vp(_req(hamming, np.int32, "hamming")))


def closest_sequences(num_refs, hamming, closest_n, closest_refs):
    """closest_seqs.h:8-9.  Draws from libc random() as the reference does."""
    lib = _lib.load()
    lib.closest_sequences(num_refs, _vp(_req(hamming, np.int32, "hamming")), closest_n,
                          _vp(_req(closest_refs, np.int32, "closest_refs")))


# --------------------------------------------------------------------------------------------
# batched closest-L
# --------------------------------------------------------------------------------------------
def hamming_batch(ref_seqs, query_seqs, device=0):
    lib = _lib.load()
    ref_seqs, query_seqs = _chars(ref_seqs, "ref_seqs"), _chars(query_seqs, "query_seqs")
    N, nb = ref_seqs.shape
    Q = query_seqs.shape[0]
    out = np.zeros((Q, N), np.int32)
    check(lib.mcq_hamming_batch(device, _vp(ref_seqs), nb, N, _vp(query_seqs), nb, Q, nb, _vp(out)))
    return out


def closest_batch(ref_seqs, query_seqs, closest_n, skip_first=False, device=0):
    """fill_closest_n (dmodel.c:356-387) for all queries; consumes libc random() like the reference."""
    lib = _lib.load()
    ref_seqs, query_seqs = _chars(ref_seqs, "ref_seqs"), _chars(query_seqs, "query_seqs")
    N, nb = ref_seqs.shape
    Q = query_seqs.shape[0]
    out = np.zeros((Q, closest_n), np.int32)
    check(lib.mcq_closest_batch(device, _vp(ref_seqs), nb, N, _vp(query_seqs), nb, Q, nb, closest_n,
                                1 if skip_first else 0, _vp(out)))
    return out


def closest_groups(ref_seqs, query_seqs, closest_n, skip_first=False, device=0):
    """The device part of closest_batch for the given queries: one flat int32 block of selection groups (to be
    exchanged between ranks and replayed in rank order with closest_replay)."""
    lib = _lib.load()
    ref_seqs, query_seqs = _chars(ref_seqs, "ref_seqs"), _chars(query_seqs, "query_seqs")
    N, nb = ref_seqs.shape
    Q = query_seqs.shape[0]
    ptr, cnt = C.POINTER(C.c_int)(), C.c_size_t()
    check(lib.mcq_closest_groups(device, _vp(ref_seqs), nb, N, _vp(query_seqs), nb, Q, nb, closest_n,
                                 1 if skip_first else 0, C.byref(ptr), C.byref(cnt)))
    try:
        return np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy()
    finally:
        lib.mcq_free(ptr)


def closest_replay(groups, closest_n, skip_first=False):
    """The sequential part: the reference's tie shuffles with libc random() for the queries of one block."""
    lib = _lib.load()
    groups = _req(np.ascontiguousarray(groups, np.int32), np.int32, "groups")
    out = np.zeros((int(groups[1]), closest_n), np.int32)
    nq = C.c_int()
    check(lib.mcq_closest_replay(_vp(groups), groups.size, closest_n, 1 if skip_first else 0, _vp(out), C.byref(nq)))
    return out


def consensus_batch(ref_seqs, closest_refs, device=0):
    """cons_query_nucls (Q, num_bases) int8 codes 0..3: per-query majority base over its closest references
    (dmodel.c:664-681)."""
    lib = _lib.load()
    ref_seqs = _chars(ref_seqs, "ref_seqs")
    closest_refs = _req(np.ascontiguousarray(closest_refs, np.int32), np.int32, "closest_refs")
    N, nb = ref_seqs.shape
    Q, L = closest_refs.shape
    out = np.zeros((Q, nb), np.int8)
    check(lib.mcq_consensus_batch(device, _vp(ref_seqs), nb, N, _vp(closest_refs), Q, L, nb, _vp(out)))
    return out


def consensus_last_timing():
    """device ms of the consensus kernel of the last consensus_batch call"""
    a = C.c_float()
    check(_lib.load().mcq_consensus_last_timing(C.byref(a)))
    return a.value


def closest_last_timing():
    """(distance kernel ms, selection kernels ms) of the last closest_batch / hamming_batch call."""
    a, b = C.c_float(), C.c_float()
    check(_lib.load().mcq_closest_last_timing(C.byref(a), C.byref(b)))
    return a.value, b.value


# --------------------------------------------------------------------------------------------
# batched context
# --------------------------------------------------------------------------------------------
class Context:
    """One shard of queries resident on one GPU (mcq_ctx_* of include/mcq_b200.h)."""

    def __init__(self, num_sites, closest_n, total_num_refs, num_query, num_hla_types, device=0):
        self.lib = _lib.load()
        self.S, self.L, self.N, self.Q, self.H = num_sites, closest_n, total_num_refs, num_query, num_hla_types
        self._h = C.c_void_p()
        dims = McqDims(num_sites, closest_n, total_num_refs, num_query, num_hla_types, device)
        check(self.lib.mcq_ctx_create(C.byref(self._h), C.byref(dims)))

    @classmethod
    def from_problem(cls, p, device=0, q0=0, q1=None):
        """Context for queries q0..q1-1 of a prep.Problem, static data and parameters uploaded."""
        q1 = p.Q if q1 is None else q1
        ctx = cls(p.S, p.L, p.N, q1 - q0, p.H, device)
        ctx.upload_static(p.ref_triplets, p.closest_refs[q0:q1], p.cons_query_nucls[q0:q1],
                          p.query_codons[q0:q1], p.hla[q0:q1], p.consensus_codons, p.n_site_codons,
                          p.site_codons)
        ctx.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
                       rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
        return ctx

    def close(self):
        if self._h:
            self.lib.mcq_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_static(self, ref_triplets, closest_refs, cons_query_nucls, query_codons, hla_types,
                      consensus_codons, num_ref_site_codons, ref_site_codons):
        a = [np.ascontiguousarray(x) for x in (ref_triplets, closest_refs, cons_query_nucls, query_codons,
                                               hla_types, consensus_codons, num_ref_site_codons,
                                               ref_site_codons)]
        _chars(a[0], "ref_triplets"); _req(a[1], np.int32, "closest_refs"); _chars(a[2], "cons_query_nucls")
        _chars(a[3], "query_codons"); _chars(a[4], "hla_types"); _chars(a[5], "consensus_codons")
        _req(a[6], np.int32, "num_ref_site_codons"); _chars(a[7], "ref_site_codons")
        assert a[0].shape == (self.N, self.S) and a[1].shape == (self.Q, self.L)
        assert a[2].shape == (self.Q, 3 * self.S) and a[3].shape == (self.Q, self.S)
        assert a[4].shape == (self.Q, self.H) and a[7].shape == (self.S, 64)
        check(self.lib.mcq_ctx_upload_static(self._h, *[_vp(x) for x in a]))

    def set_params(self, mu, kappa, phi, prior=None, R=None, omega=None, rate_out=None, esc=None, rev=None):
        keep = []

        def ptr(x):
            if x is None:
                return None
            x = np.ascontiguousarray(x, np.float64)
            keep.append(x)
            return x.ctypes.data_as(_lib.c_double_p)

        prm = McqParams(mu, kappa, phi, ptr(prior), ptr(R), ptr(omega), ptr(rate_out), ptr(esc), ptr(rev))
        check(self.lib.mcq_ctx_set_params(self._h, C.byref(prm)))

    def emission(self, start, end, which_hla=DO_ALL_HLAS, mode=DO_HLA_BOTH):
        check(self.lib.mcq_ctx_emission(self._h, start, end, which_hla, mode))

    def init_state(self):
        v = C.c_double()
        check(self.lib.mcq_ctx_init_state(self._h, C.byref(v)))
        return v.value

    def full_llk(self, with_emission=True):
        v = C.c_double()
        check(self.lib.mcq_ctx_full_llk(self._h, 1 if with_emission else 0, C.byref(v)))
        return v.value

    def llk_only(self, with_emission=True):
        v = C.c_double()
        check(self.lib.mcq_ctx_llk_only(self._h, 1 if with_emission else 0, C.byref(v)))
        return v.value

    def tune(self):
        """Times the full sweeps at every residency; returns the blocks per SM chosen for
        (forward stored, forward likelihood-only, backward)."""
        k = (C.c_int * 3)()
        check(self.lib.mcq_ctx_tune(self._h, k))
        return tuple(k)

    def backward_fill(self):
        check(self.lib.mcq_ctx_backward_fill(self._h))

    def propose(self, kind, start, end=None, value0=0.0, value1=0.0, hla=0, split=None):
        end = start if end is None else end
        split = end + 1 if split is None else split
        prop = McqProposal(kind, start, end, hla, split, value0, value1)
        v = C.c_double()
        check(self.lib.mcq_ctx_propose(self._h, C.byref(prop), C.byref(v)))
        return v.value

    def accept(self):
        check(self.lib.mcq_ctx_accept(self._h))

    def reject(self):
        check(self.lib.mcq_ctx_reject(self._h))

    def flush(self):
        check(self.lib.mcq_ctx_flush(self._h))

    def set_option(self, option, value):
        check(self.lib.mcq_ctx_set_option(self._h, option, value))

    def enable_timing(self, level=1):
        """0 off, 1 CUDA events around each API call, 2 also around each kernel launch."""
        check(self.lib.mcq_ctx_enable_timing(self._h, int(level)))

    def kernel_time(self, kclass, reset=False):
        """(total device ms, launches) of one kernel class since the last reset."""
        ms, n = C.c_float(), C.c_int()
        check(self.lib.mcq_ctx_kernel_time(self._h, kclass, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value

    def timer_start(self):
        check(self.lib.mcq_ctx_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.lib.mcq_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def synchronize(self):
        check(self.lib.mcq_ctx_synchronize(self._h))

    def last_timing(self):
        ms, n = C.c_float(), C.c_int()
        check(self.lib.mcq_ctx_last_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def device_bytes(self):
        return self.lib.mcq_ctx_device_bytes(self._h)

    def matrix(self, which, q):
        """(F|B|tmp_F)[q] as the reference lays it out, [L][S], plus its rnorm[S]."""
        out = np.zeros((self.L, self.S))
        rn = np.zeros(self.S, np.int32)
        check(self.lib.mcq_ctx_get_matrix(self._h, which, q, _vp(out), _vp(rn)))
        return out, rn

    def table(self, which, q):
        """codon_to_codon_HLA[q] / A[q] as the reference lays them out, [64][S]."""
        out = np.zeros((64, self.S))
        check(self.lib.mcq_ctx_get_table(self._h, which, q, _vp(out)))
        return out

    def query_llk(self, proposed=False):
        out = np.zeros(self.Q)
        check(self.lib.mcq_ctx_get_query_llk(self._h, 1 if proposed else 0, _vp(out)))
        return out

    def params(self):
        mu = C.c_double()
        R = np.zeros(self.S); om = np.zeros(self.S); ro = np.zeros((64, self.S))
        esc = np.zeros((self.H, self.S)); rev = np.zeros(self.S)
        check(self.lib.mcq_ctx_get_params(self._h, C.byref(mu), _vp(R), _vp(om), _vp(ro), _vp(esc), _vp(rev)))
        return dict(mu=mu.value, R=R, omega=om, rate_out=ro, esc=esc, rev=rev)

    def peer_handle(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        check(self.lib.mcq_ctx_peer_handle(self._h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def attach_peers(self, nranks, rank, handles):
        """handles: list of nranks 64-byte IPC handles (one per rank, from peer_handle())."""
        flat = (C.c_ubyte * (64 * nranks)).from_buffer_copy(b"".join(handles))
        check(self.lib.mcq_ctx_attach_peers(self._h, nranks, rank, C.cast(flat, C.c_void_p)))

    def exchange_sum(self, value):
        """Sum of one double per rank over the attached ranks (rank order; identity without peers)."""
        v = C.c_double(value)
        check(self.lib.mcq_ctx_exchange_sum(self._h, C.byref(v)))
        return v.value

    def attach_comm(self, nranks, rank, unique_id: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        check(self.lib.mcq_ctx_attach_comm(self._h, nranks, rank, C.cast(buf, C.c_void_p)))


def host_register(arr):
    """Page-lock a numpy array's buffer (uploads from it then run at full PCIe rate)."""
    check(_lib.load().mcq_host_register(_vp(arr), arr.nbytes))


def host_unregister(arr):
    check(_lib.load().mcq_host_unregister(_vp(arr)))


def comm_unique_id() -> bytes:
    lib = _lib.load()
    buf = (C.c_ubyte * 128)()
    check(lib.mcq_comm_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)
