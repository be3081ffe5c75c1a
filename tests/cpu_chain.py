"""TEST code: the bodies of the reference's move loops restated on top of a CPU checker
(the C oracle or the compiled reference), in the reference's array shapes.  It plays the role
of `dmodel`'s state (c2c, A, F, B, rnorms, parameters) so that the CUDA context can be compared
with it proposal by proposal.

mcmc_moves.c:213-224 (REC), :345-395 (SEL), :789-819 (ESC/REV windows), :1082-1134 (MU) for the
proposal; :261-279, :445-463, :955-981, :1158-1181 for the accept path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import cpu_libs
from mcqueen_b200 import genetics as gen

REC, SEL, ESC, REV, MU = 0, 1, 2, 3, 4


class CpuState:
    def __init__(self, chk, p):
        self.chk, self.p = chk, p
        self.orc = cpu_libs.Checker("oracle")
        self.R = p.R.copy(); self.omega = p.omega.copy(); self.mu = float(p.mu)
        self.esc = p.esc.copy(); self.rev = p.rev.copy()
        self.rate_out = p.rate_out_of_codons(self.omega)
        st = cpu_libs.full_state(chk, p)
        self.c2c, self.A, self.F, self.B = st["c2c"], st["A"], st["F"], st["B"]
        self.Fn, self.Bn = st["Fn"], st["Bn"]
        self.llk_q = st["llk"].copy()
        self.tF = np.zeros_like(self.F); self.tFn = np.zeros_like(self.Fn)
        self.pending = None

    @property
    def llk(self):
        return self._seq_sum(self.llk_q)

    @staticmethod
    def _seq_sum(v):
        acc = 0.0
        for x in v:
            acc += float(x)
        return acc

    def _eval(self, c2c, R, start, end, dot=True):
        p = self.p
        out = np.zeros(p.Q)
        for q in range(p.Q):
            self.chk.forward(p, q, c2c[q], R, self.F[q], self.Fn[q], self.tFn[q], self.tF[q], start, end)
            if dot:
                out[q] = self.chk.site_llk(p, self.tF[q], self.tFn[q], self.B[q], self.Bn[q], end)
            else:
                out[q] = self.chk.site_llk(p, self.tF[q], self.tFn[q], None, None, end)
        return out

    def propose(self, kind, start, end=None, value0=0.0, value1=0.0, hla=0, split=None):
        p = self.p
        end = start if end is None else end
        split = end + 1 if split is None else split
        new = dict(kind=kind, start=start, end=end, hla=hla)
        if kind == REC:
            R = self.R.copy(); R[start] = value0
            new.update(R=R)
            llk_q = self._eval(self.c2c, R, start, end)
        elif kind == SEL:
            omega = self.omega.copy(); omega[start] = value0
            rate = self.rate_out.copy()
            for k in range(p.n_site_codons[start]):
                c1 = int(p.site_codons[start, k])
                rate[c1, start] = p.kappa * gen.BETA_S[c1] + gen.BETA_V[c1] + \
                    value0 * (p.kappa * gen.ALPHA_S[c1] + gen.ALPHA_V[c1])
            c2c, A = self.c2c.copy(), self.A.copy()
            self.chk.emission(p, c2c, A, start, end, -1, 2, omega=omega, esc=self.esc, rev=self.rev,
                              mu=self.mu, rate_out=rate)
            new.update(omega=omega, rate_out=rate, c2c=c2c, A=A)
            llk_q = self._eval(c2c, self.R, start, end)
        elif kind in (ESC, REV):
            esc, rev = self.esc.copy(), self.rev.copy()
            tgt = esc[hla] if kind == ESC else rev[0]
            tgt[start:split] = value0
            tgt[split:end + 1] = value1
            c2c, A = self.c2c.copy(), self.A.copy()
            self.chk.emission(p, c2c, A, start, end, hla, 0 if kind == ESC else 1, omega=self.omega,
                              esc=esc, rev=rev, mu=self.mu, rate_out=self.rate_out)
            new.update(esc=esc, rev=rev, c2c=c2c, A=A)
            llk_q = self._eval(c2c, self.R, start, end)
        elif kind == MU:
            c2c, A = np.zeros_like(self.c2c), np.zeros_like(self.A)
            self.orc.lib.orc_mu_rescale(p.S, p.N, p.Q, cpu_libs.ip(p.n_site_codons), cpu_libs.cp(p.site_codons),
                                        cpu_libs.dp(self.c2c), cpu_libs.dp(self.A), cpu_libs.dp(c2c),
                                        cpu_libs.dp(A), cpu_libs.cp(p.query_codons), C.c_double(self.mu),
                                        C.c_double(value0), C.c_double(p.phi), cpu_libs.dp(p.prior))
            new.update(mu=value0, c2c=c2c, A=A, start=0, end=p.S - 1)
            llk_q = self._eval(c2c, self.R, 0, p.S - 1, dot=False)
        else:
            raise ValueError(kind)
        new["llk_q"] = llk_q
        self.pending = new
        return self._seq_sum(llk_q)

    def accept(self):
        p, n = self.p, self.pending
        for k in ("R", "omega", "rate_out", "esc", "rev", "mu", "c2c", "A"):
            if k in n:
                setattr(self, k, n[k])
        s0, s1 = n["start"], n["end"]
        if n["kind"] == MU:
            self.F, self.tF = self.tF, self.F
            self.Fn, self.tFn = self.tFn, self.Fn
            for q in range(p.Q):
                self.chk.backward(p, q, self.c2c[q], self.R, self.B[q], self.Bn[q], p.S - 1)
        else:
            for q in range(p.Q):
                self.F[q][:, s0:s1 + 1] = self.tF[q][:, s0:s1 + 1]
                self.Fn[q][s0:s1 + 1] = self.tFn[q][s0:s1 + 1]
                if s1 + 1 < p.S:
                    self.chk.forward(p, q, self.c2c[q], self.R, self.F[q], self.Fn[q], self.Fn[q], self.F[q],
                                     s1 + 1, p.S - 1)
                self.chk.backward(p, q, self.c2c[q], self.R, self.B[q], self.Bn[q], s1)
        self.llk_q = n["llk_q"]
        self.pending = None

    def reject(self):
        self.pending = None
