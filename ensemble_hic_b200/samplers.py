"""HMC and Gibbs samplers with the constructor / ``sample()`` surface the reference uses
(binf.samplers.hmc.HMCSampler, binf.samplers.gibbs.GibbsSampler: setup_functions.py:239-264,
run_simulation.py:98-99).

binf and csb are absent from the reference tree; the algorithm restated here -- csb's
FastLeapFrog (half kick, (L-1) x (drift, kick), drift, half kick), unit mass matrix,
Metropolis accept on H = -log p + |p|^2/2, multiplicative timestep adaptation while
n_samples < timestep_adaption_limit -- is "parity unpinned": the reference has no test for it
(SURVEY.md 8c).  Momenta and the acceptance uniform come from numpy's global RNG exactly as in
the reference stack, so a seeded run reproduces accept/reject decisions of the CPU restatement
in oracle/oracle.py.

For a :class:`~ensemble_hic_b200.posteriors.GPUPosterior` the whole trajectory is ONE C-ABI
call (ehic_hmc_trajectory): positions and momenta stay on the device for all L+1 gradient
evaluations.
"""
import numpy as np

from ._binf import BinfState


class HMCSampler(object):
    def __init__(self, pdf, state, timestep, nsteps, timestep_adaption_limit=0,
                 adaption_uprate=1.05, adaption_downrate=0.95, variable_name="structures"):
        self.pdf = pdf
        self.state = np.array(state, dtype=np.float64)
        self.timestep = float(timestep)
        self.nsteps = int(nsteps)
        self.variable_name = variable_name
        self._adaption_limit = int(timestep_adaption_limit)
        self._uprate, self._downrate = float(adaption_uprate), float(adaption_downrate)
        self.n_samples = 0
        self.n_accepted = 0
        self.last_move_accepted = False
        self.last_H = (np.nan, np.nan)

    @property
    def acceptance_rate(self):
        return self.n_accepted / float(self.n_samples) if self.n_samples else 0.0

    # ---- trajectory ---------------------------------------------------------------------------
    def _energy(self, x):
        return -self.pdf.log_prob(**{self.variable_name: x})

    def _gradient(self, x):
        return self.pdf.gradient(**{self.variable_name: x})

    def _propose_host(self, x, p):
        eps, L = self.timestep, self.nsteps
        H0 = self._energy(x) + 0.5 * np.sum(p * p)
        x = x.copy(); p = p.copy()
        p -= 0.5 * eps * self._gradient(x)
        for _ in range(L - 1):
            x += eps * p
            p -= eps * self._gradient(x)
        x += eps * p
        p -= 0.5 * eps * self._gradient(x)
        H1 = self._energy(x) + 0.5 * np.sum(p * p)
        return x, H0, H1

    def _propose_device(self, x, p):
        import torch
        from .gamma_prior import GammaPrior
        pdf = self.pdf
        pdf.sync_engine()
        m = pdf.engine
        dev = torch.device("cuda", m.device)
        t = lambda a: torch.as_tensor(np.atleast_1d(np.asarray(a, dtype=np.float64)), device=dev)
        lam, bet = pdf.tempering()
        x_t, p_t = t(x).reshape(1, -1).clone(), t(p).reshape(1, -1).clone()
        H = torch.zeros(1, 4, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            m.hmc_trajectory_device(1, self.nsteps, x_t, p_t, t(pdf._norm({})), t(lam), t(bet), t(self.timestep), H)
            Hh = H.cpu().numpy()[0]
            x1 = x_t.cpu().numpy()[0]
        # the norm prior (if any) does not depend on the structures: it cancels in H1 - H0
        return x1, Hh[0] + Hh[1], Hh[2] + Hh[3]

    def sample(self):
        from .posteriors import GPUPosterior
        x = self.state
        p = np.random.normal(size=x.shape)
        fused = isinstance(self.pdf, GPUPosterior) and self.pdf._standard() and self.variable_name == "structures"
        if fused and ("norm" in self.pdf.variables):
            fused = False           # norm must be fixed (conditional posterior) for the device path
        x1, H0, H1 = (self._propose_device if fused else self._propose_host)(x, p)
        u = np.random.random()
        dH = H1 - H0
        accept = bool(np.isfinite(dH) and u < np.exp(min(0.0, -dH)))
        self.last_H = (H0, H1)
        self.last_move_accepted = accept
        self.n_samples += 1
        if accept:
            self.n_accepted += 1
            self.state = x1
        if self.n_samples < self._adaption_limit:
            self.timestep *= self._uprate if accept else self._downrate
        return self.state

    def get_last_draw_stats(self):
        return {self.variable_name: dict(accepted=self.last_move_accepted, stepsize=self.timestep,
                                         total_energy=self.last_H)}


class GibbsSampler(object):
    """Sweeps over the variables; each subsampler sees the posterior conditioned on the rest."""

    def __init__(self, pdf, state, subsamplers):
        self.pdf = pdf
        self.state = state if isinstance(state, BinfState) else BinfState(dict(state))
        self.subsamplers = subsamplers
        self._order = [v for v in ("structures", "norm") if v in subsamplers] + \
                      sorted(v for v in subsamplers if v not in ("structures", "norm"))

    def sample(self):
        for var in self._order:
            sub = self.subsamplers[var]
            fixed = {v: val for v, val in self.state.variables.items() if v != var}
            sub.pdf = self.pdf.conditional_factory(**fixed)
            if hasattr(sub, "variable_name"):
                sub.state = np.asarray(self.state.variables[var], dtype=np.float64)
            new = sub.sample()
            self.state.update_variables(**{var: new})
        return self.state

    def get_last_draw_stats(self):
        out = {}
        for sub in self.subsamplers.values():
            if hasattr(sub, "get_last_draw_stats"):
                out.update(sub.get_last_draw_stats())
        return out
