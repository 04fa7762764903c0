"""Replica adaptor (reference: replica.py:12-14): energy of a state under this replica's pdf."""


class CompatibleReplica(object):
    """``get_energy(state) = -pdf.log_prob(**state.variables)``.  The reference derives from
    rexfw.replicas.Replica (absent); the batched multi-GPU driver in :mod:`.rex` supersedes the
    per-process replica objects, this class keeps the call site for single-replica use."""

    def __init__(self, name, state, pdf, sampler_class=None, sampler_params=None, proposers=None,
                 output_folder=None, comm=None):
        self.name, self.state, self.pdf = name, state, pdf
        self.sampler_class, self.sampler_params = sampler_class, sampler_params or {}
        self.proposers, self.output_folder, self.comm = proposers, output_folder, comm

    def get_energy(self, state):
        return -self.pdf.log_prob(**state.variables)
