"""Codon encodings and model constants of the McQueen path (host side, numpy).

Encodings follow amino_acids.c:10-17 / amino_acids.h:24-30 of the reference: bases
``T,C,A,G -> 0..3`` (anything else 4), a codon is the base-4 number of its three
bases, a *triplet* the base-5 number (digit 4 = unknown).  The substitution-class
tables (constants.c:405-690: NS_TS, NS_TV, S_TS, S_TV and their row sums alpha_S,
alpha_V, beta_S, beta_V) are *derived* here from the standard genetic code rather
than stored; tests/test_tables.py checks the derivation against the reference's
compiled tables wherever oracle/_ref exists.
"""
from __future__ import annotations

import numpy as np

AMINO = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
STOP_CODONS = (10, 11, 14)

# model defaults: constants.c:85 (kappa), :88 (phi), :13-14 (BIG, BIGI), :17,:20 (init mu, R)
KAPPA = 7.595744
PHI = 0.9482171
BIG = 1000000.0
BIGI = 100.0
INIT_MU = 20.0
INIT_R = 0.01

# codon usage "prior_C1" (data, constants.c:93-109); dmodel.c:867-869 replaces the last
# entry by 1 - sum(others) before first use - see normalised_prior().
PRIOR_C1 = np.array([
    0.01660719, 0.005966229, 0.04147626, 0.02127574, 0.0005901346, 0.0005592375, 0.0110179,
    0.000117409, 0.01344333, 0.0002224591, 1.235884e-05, 4.943536e-05, 0.008641918,
    0.0005159815, 3.08971e-06, 0.02271246, 0.004325594, 0.0001143193, 0.003095889, 0.005067124,
    0.0264232, 0.01026711, 0.04408398, 0.000692095, 0.01278522, 0.004915728, 0.04811605,
    0.01067495, 6.179419e-06, 1.235884e-05, 0.0003027916, 0.00414639, 0.01751247, 0.009772752,
    0.03740403, 0.05353231, 0.0158224, 0.02728214, 0.02607715, 0.0002286385, 0.03277255,
    0.01464831, 0.03445644, 0.01345569, 0.01164512, 0.01513031, 0.03834948, 0.008382382,
    0.004328683, 0.009099195, 0.03378907, 0.01560921, 0.02094514, 0.01382336, 0.04780708,
    0.005224699, 0.02168976, 0.01999351, 0.03876968, 0.01819221, 0.002045388, 0.006686132,
    0.04171726, 0.02556735], dtype=np.float64)


def normalised_prior(prior=None):
    """dmodel.c:867-869: last entry := 1 - (sequential sum of the first 63)."""
    p = np.array(PRIOR_C1 if prior is None else prior, dtype=np.float64)
    acc = 0.0
    for i in range(63):
        acc += float(p[i])
    p[63] = 1 - acc
    return p


def _build_tables():
    t = np.zeros((5, 64, 64), np.int32)
    for c1 in range(64):
        for c2 in range(64):
            d1 = ((c1 >> 4) & 3, (c1 >> 2) & 3, c1 & 3)
            d2 = ((c2 >> 4) & 3, (c2 >> 2) & 3, c2 & 3)
            diff = [i for i in range(3) if d1[i] != d2[i]]
            if len(diff) >= 2:
                t[4, c1, c2] = 1
            elif len(diff) == 1:
                a, b = d1[diff[0]], d2[diff[0]]
                ts = (a ^ b) == 1
                syn = AMINO[c1] == AMINO[c2]
                t[(0 if ts else 1) if not syn else (2 if ts else 3), c1, c2] = 1
    return t


_T = _build_tables()
NS_TS, NS_TV, S_TS, S_TV, TWO_STEP = _T[0], _T[1], _T[2], _T[3], _T[4]
ALPHA_S, ALPHA_V = NS_TS.sum(1), NS_TV.sum(1)
BETA_S, BETA_V = S_TS.sum(1), S_TV.sum(1)


def one_step_neighbours(c):
    return [d for d in range(64) if (NS_TS[c, d] | NS_TV[c, d] | S_TS[c, d] | S_TV[c, d])]


_CODE = np.full(256, 4, np.int8)
for _ch, _v in zip("TCAG", range(4)):
    _CODE[ord(_ch)] = _v
    _CODE[ord(_ch.lower())] = _v
_ASCII = np.frombuffer(b"TCAG", dtype=np.uint8)


def base_codes(seqs):
    """ASCII (.., n) u8 -> codes 0..4 (amino_acids.c:10-17)."""
    return _CODE[np.asarray(seqs, dtype=np.uint8)]


def codons_to_ascii(codons):
    c = np.asarray(codons).astype(np.int64)
    out = np.empty(c.shape + (3,), np.uint8)
    out[..., 0] = _ASCII[(c >> 4) & 3]
    out[..., 1] = _ASCII[(c >> 2) & 3]
    out[..., 2] = _ASCII[c & 3]
    return out.reshape(c.shape[:-1] + (c.shape[-1] * 3,))


def to_codons(codes):
    """codes (.., 3S) all in 0..3 -> base-4 codon numbers (amino_acids.h:24-26)."""
    c = np.asarray(codes).astype(np.int16).reshape(codes.shape[:-1] + (-1, 3))
    return (c[..., 0] * 16 + c[..., 1] * 4 + c[..., 2]).astype(np.int8)


def to_triplets(codes):
    """codes (.., 3S) in 0..4 -> base-5 triplet numbers (amino_acids.h:28-30)."""
    c = np.asarray(codes).astype(np.int16).reshape(codes.shape[:-1] + (-1, 3))
    return (c[..., 0] * 25 + c[..., 1] * 5 + c[..., 2]).astype(np.int8)
