"""Builds posterior, samplers and replica schedules from a config file -- the public setup
surface of the reference (setup_functions.py), same function names, arguments and config
format (all config values stay strings; booleans are the literals 'True' / 'False').

What differs: ``make_posterior`` returns a :class:`~ensemble_hic_b200.posteriors.GPUPosterior`
whose components share ONE resident GPU model, and ``setup_default_re_master`` returns the
batched multi-GPU replica-exchange driver of :mod:`ensemble_hic_b200.rex` instead of an
rexfw ExchangeMaster (rexfw / mpi4py are not part of this stack).
"""
import configparser
import os

import numpy as np


def parse_config_file(config_file):
    """INI file -> {section: {option: string}} (reference: setup_functions.py:9-35)."""
    config = configparser.ConfigParser(interpolation=None)
    if not config.read(config_file):
        raise IOError("could not read config file %r" % (config_file,))
    out = {}
    for section in config.sections():
        out[section] = {}
        for option in config.options(section):
            try:
                out[section][option] = config.get(section, option)
            except Exception:
                out[section][option] = None
    return out


def update_ensemble_setting(settings):
    """copies [replica] ensemble ('boltzmann' | 'tsallis', default 'boltzmann') into the
    nonbonded prior settings (reference: setup_functions.py:37-62)"""
    rps = settings["replica"]
    if "ensemble" not in rps or rps["ensemble"] == "boltzmann":
        settings["nonbonded_prior"].update(ensemble="boltzmann")
    elif rps["ensemble"] == "tsallis":
        settings["nonbonded_prior"].update(ensemble="tsallis")
    else:
        raise ValueError("ensemble has to be either not present in cfg file, or set to 'tsallis' or 'boltzmann'")
    return settings


def make_posterior(settings, max_replicas=1, exact=None, device=None, fp32=None):
    """``p = make_posterior(parse_config_file('config.cfg'))`` (reference: setup_functions.py:64-98).

    Extra keyword arguments size the resident GPU model: ``max_replicas`` tempered copies can be
    evaluated per call by the batched drivers; ``exact`` selects the bit-exact arithmetic mode, ``fp32``
    the optional FP32 pair arithmetic (1e-5 relative; dense contact lists only)."""
    from . import _models
    from .engine import Model
    from .posteriors import GPUPosterior

    settings = update_ensemble_setting(settings)
    n_beads = int(settings["general"]["n_beads"])
    n_structures = int(settings["general"]["n_structures"])
    priors = make_priors(settings["nonbonded_prior"], settings["backbone_prior"], settings["sphere_prior"],
                         n_beads, n_structures)
    bead_radii = priors["nonbonded_prior"].forcefield.bead_radii
    likelihood = make_likelihood(settings["forward_model"], settings["general"]["error_model"],
                                 settings["data_filtering"], settings["general"]["data_file"],
                                 n_structures, bead_radii)
    if "norm" in [v.strip() for v in settings["general"]["variables"].split(",")]:
        priors.update(norm_prior=make_norm_prior(settings["norm_prior"], likelihood, n_structures))

    # one resident model with every static input of the hot path
    fwm = likelihood.forward_model
    bbp = priors["backbone_prior"]
    ll, ul = bbp._flat_limits()
    sphere = None
    if "sphere_prior" in priors:
        sp = priors["sphere_prior"]
        sphere = (sp["sphere_radius"].value, sp["sphere_k"].value, sp["sphere_center"].value)
    engine = Model(n_structures, bead_radii, fwm.data_points, fwm["contact_distances"].value,
                   alpha=float(settings["forward_model"]["alpha"]),
                   k_nb=float(priors["nonbonded_prior"].forcefield.force_constant),
                   k_bb=bbp["k_bb"].value, bb_lower=ll, bb_upper=ul, mol_ranges=bbp._mol_ranges,
                   sphere=sphere, max_replicas=max_replicas,
                   exact=_models.default_exact() if exact is None else exact,
                   fp32=_models.default_fp32() if fp32 is None else fp32,
                   device=_models.default_device() if device is None else device)
    fwm.attach_engine(engine)
    for p in priors.values():
        if hasattr(p, "attach_engine"):
            p.attach_engine(engine)
    full_posterior = GPUPosterior({likelihood.name: likelihood}, priors, engine=engine)
    return make_conditional_posterior(full_posterior, settings)


def make_norm_prior(norm_prior_settings, likelihood, n_structures):
    """Gamma prior for the scaling factor; 'auto' picks rate = 1/S and
    shape = mean(nonzero counts)/S (reference: setup_functions.py:100-133)."""
    from .gamma_prior import NormGammaPrior
    shape = norm_prior_settings["shape"]
    rate = norm_prior_settings["rate"]
    if shape == rate == "auto":
        rate = 1.0 / n_structures
        dp = likelihood.forward_model.data_points[:, 2]
        shape = np.mean(dp[dp > 0]) / float(n_structures)
    else:
        shape, rate = float(shape), float(rate)
    return NormGammaPrior(shape, rate)


def expspace(min, max, a, N):
    """N exponentially spaced values between min and max with rate a
    (reference: setup_functions.py:135-159)."""
    n = np.arange(1, N + 1)
    return (max - min) / (np.exp(a * (N - 1.0)) - 1.0) * (np.exp(a * (n - 1.0)) - 1.0) + float(min)


def make_replica_schedule(replica_params, n_replicas):
    """{'lammda': array, 'beta': array} for a 'linear' or 'exponential' schedule
    (reference: setup_functions.py:161-217)."""
    l_min, l_max = float(replica_params["lambda_min"]), float(replica_params["lambda_max"])
    b_min, b_max = float(replica_params["beta_min"]), float(replica_params["beta_max"])
    if replica_params["schedule"] == "linear":
        if replica_params["separate_prior_annealing"] == "True":
            n_b = int(np.floor(n_replicas / 2))
            n_l = n_replicas - n_b
            lambdas = np.concatenate((np.zeros(n_b) + l_min, np.linspace(l_min, l_max, n_l)))
            betas = np.concatenate((np.linspace(b_min, b_max, n_b), np.zeros(n_l) + b_max))
            return {"lammda": lambdas, "beta": betas}
        return {"lammda": np.linspace(l_min, l_max, n_replicas), "beta": np.linspace(b_min, b_max, n_replicas)}
    if replica_params["schedule"] == "exponential":
        l_rate, b_rate = float(replica_params["lambda_rate"]), float(replica_params["beta_rate"])
        return {"lammda": expspace(l_min, l_max, l_rate, n_replicas),
                "beta": expspace(b_min, b_max, b_rate, n_replicas)}
    raise ValueError("Schedule has to be either a file name, 'lambda_beta', or 'exponential'")


def load_schedule(replica_params, n_replicas):
    """the schedule selection of run_simulation.py:28-47: linear / exponential / gaussian-cdf
    ('*.py') / pickled dict"""
    sched = replica_params["schedule"]
    if sched in ("linear", "exponential"):
        return make_replica_schedule(replica_params, n_replicas)
    if sched[-3:] == ".py":
        from scipy import stats
        space = np.linspace(0, 1, n_replicas)
        m, s = float(replica_params["gauss_mean"]), float(replica_params["gauss_std"])
        delta = np.concatenate(([0.0], stats.norm.pdf(space, m, s)))
        betas = np.cumsum(delta)
        betas /= betas[-1]
        return {"lammda": betas, "beta": betas}
    import pickle
    with open(sched, "rb") as f:
        return pickle.load(f, encoding="latin1")


def make_subsamplers(posterior, initial_state, structures_hmc_params):
    """{'structures': HMCSampler, ['norm': NormGammaSampler]} (reference: setup_functions.py:219-264)."""
    from .samplers import HMCSampler
    p = posterior
    variables = list(initial_state.keys())
    tl = int(structures_hmc_params["trajectory_length"])
    timestep = float(structures_hmc_params["timestep"])
    adaption_limit = int(structures_hmc_params["adaption_limit"])
    xpdf = p.conditional_factory(**{var: value for var, value in initial_state.items() if var != "structures"})
    subsamplers = dict(structures=HMCSampler(xpdf, initial_state["structures"], timestep, tl,
                                             variable_name="structures",
                                             timestep_adaption_limit=adaption_limit))
    if "norm" in variables:
        from .error_models import PoissonEM
        if isinstance(p.likelihoods["ensemble_contacts"].error_model, PoissonEM):
            from .npsamplers import NormGammaSampler
            subsamplers.update(norm=NormGammaSampler())
        else:
            raise NotImplementedError("Norm sampling only implemented for Poisson error model!")
    return subsamplers


def make_elongated_structures(bead_radii, n_structures):
    """fully stretched chains along x, centred (reference: setup_functions.py:267-290)"""
    r = np.asarray(bead_radii, dtype=np.float64)
    X = np.concatenate(([r[0]], r[0] + np.cumsum(r[1:] + r[:-1])))
    X = X - np.mean(X)
    X = np.array([X, np.zeros(len(r)), np.zeros(len(r))]).T[None, :]
    return X.repeat(n_structures, 0).ravel().astype(float)


def make_random_structures(bead_radii, n_structures):
    """bead positions ~ N(0, (mean(r) * N**0.333)^2) (reference: setup_functions.py:292-314)"""
    d = bead_radii.mean() * len(bead_radii) ** 0.333
    return np.random.normal(scale=d, size=(n_structures, len(bead_radii), 3)).ravel()


def setup_initial_state(initial_state_params, posterior):
    """BinfState with 'structures' (random, or loaded from .npy) and, if sampled, 'norm'
    (reference: setup_functions.py:316-368; 'elongated' means *random* there too)."""
    from ._binf import BinfState
    p = posterior
    n_structures = p.likelihoods["ensemble_contacts"].forward_model.n_structures
    structures = initial_state_params["structures"]
    norm = initial_state_params["norm"]
    if structures == "elongated":
        bead_radii = posterior.priors["nonbonded_prior"].forcefield.bead_radii
        structures = make_random_structures(bead_radii, n_structures)
    else:
        try:
            structures = np.load(structures)
            if len(structures.shape) > 1:
                structures = structures[np.random.choice(structures.shape[0], n_structures, replace=False)]
                structures = structures.ravel()
        except Exception:
            raise ValueError("Couldn't load initial structures from file {}!".format(structures))
    init_state = BinfState({"structures": structures})
    if "norm" in p.variables:
        init_state.update_variables(norm=float(norm))
    return init_state


def make_conditional_posterior(posterior, settings):
    """fix 'norm' at [initial_state] norm when it is not sampled (reference: setup_functions.py:370-392)"""
    variables = [x.strip() for x in settings["general"]["variables"].split(",")]
    if "norm" not in variables:
        return posterior.conditional_factory(norm=settings["initial_state"]["norm"])
    return posterior


def make_backbone_prior(bead_radii, backbone_prior_params, n_beads, n_structures):
    """lower limits 0, upper limits r_j + r_{j+1}, per molecule (reference: setup_functions.py:395-440)"""
    from .backbone_prior import BackbonePrior
    mol_ranges = backbone_prior_params.get("mol_ranges", None)
    if mol_ranges is None:
        mol_ranges = np.array([0, n_beads])
    else:
        mol_ranges = np.loadtxt(mol_ranges).astype(int)
    bb_ll = [np.zeros(mol_ranges[i + 1] - mol_ranges[i] - 1) for i in range(len(mol_ranges) - 1)]
    bb_ul = [np.array([bead_radii[j] + bead_radii[j + 1] for j in range(mol_ranges[i], mol_ranges[i + 1] - 1)])
             for i in range(len(mol_ranges) - 1)]
    return BackbonePrior("backbone_prior", lower_limits=bb_ll, upper_limits=bb_ul,
                         k_bb=float(backbone_prior_params["force_constant"]),
                         n_structures=n_structures, mol_ranges=mol_ranges)


def make_priors(nonbonded_prior_params, backbone_prior_params, sphere_prior_params, n_beads, n_structures):
    """{'nonbonded_prior', 'backbone_prior'[, 'sphere_prior']} (reference: setup_functions.py:442-489)"""
    nb_params = nonbonded_prior_params
    try:
        bead_radii = np.ones(n_beads) * float(nb_params["bead_radii"])
    except ValueError:
        bead_radii = np.loadtxt(nb_params["bead_radii"], dtype=float)
    NBP = make_nonbonded_prior(nb_params, bead_radii, n_structures)
    BBP = make_backbone_prior(bead_radii, backbone_prior_params, n_beads, n_structures)
    priors = {NBP.name: NBP, BBP.name: BBP}
    if sphere_prior_params["active"] == "True":
        SP = make_sphere_prior(sphere_prior_params, bead_radii, n_structures)
        priors[SP.name] = SP
    return priors


def make_nonbonded_prior(nb_params, bead_radii, n_structures):
    """Boltzmann prior over the neighbour-list force field, beta = 1 (reference: setup_functions.py:491-525)"""
    from .forcefields import NBLForceField as ForceField
    forcefield = ForceField(bead_radii, float(nb_params["force_constant"]))
    if "ensemble" not in nb_params or nb_params["ensemble"] == "boltzmann":
        from .nonbonded_prior import BoltzmannNonbondedPrior2
        return BoltzmannNonbondedPrior2("nonbonded_prior", forcefield, n_structures=n_structures, beta=1.0)
    raise NotImplementedError("ensemble %r" % (nb_params["ensemble"],))


def make_sphere_prior(sphere_prior_params, bead_radii, n_structures):
    """radius 'auto' = 2 mean(r) N^(1/3) (reference: setup_functions.py:527-560)"""
    from .sphere_prior import SpherePrior
    radius = sphere_prior_params["radius"]
    if radius == "auto":
        radius = 2 * bead_radii.mean() * len(bead_radii) ** (1 / 3.0)
    else:
        radius = float(radius)
    return SpherePrior("sphere_prior", sphere_radius=radius, sphere_k=float(sphere_prior_params["force_constant"]),
                       n_structures=n_structures, bead_radii=bead_radii)


def parse_data(data_file, data_filtering_params):
    """'i j count' rows -> filtered int array [M,3] (reference: setup_functions.py:562-574):
    drop zero counts iff include_zero_counts == 'False'; sort rows by count (numpy's default,
    non-stable argsort -- kept); drop the lowest fraction; keep |i-j| > ignore_sequential_neighbors.

    The reference multiplies the *string* disregard_lowest by len(data), which only works for
    '0'; the numeric meaning is implemented here (identical for '0')."""
    disregard_lowest = float(data_filtering_params["disregard_lowest"])
    ignore_seq_nbs = int(data_filtering_params["ignore_sequential_neighbors"])
    include_zero_counts = data_filtering_params["include_zero_counts"]
    data = np.loadtxt(data_file, dtype=int)
    if include_zero_counts == "False":
        data = data[data[:, 2] > 0]
    data = data[np.argsort(data[:, 2])]
    data = data[int(disregard_lowest * len(data)):]
    data = data[np.abs(data[:, 0] - data[:, 1]) > ignore_seq_nbs]
    return data


def make_likelihood(forward_model_params, error_model, data_filtering_params, data_file, n_structures, bead_radii):
    """Likelihood('ensemble_contacts', EnsembleContactsFWM('fwm'), PoissonEM('ensemble_contacts_em'),
    lammda=1) conditioned on smooth_steepness = alpha (reference: setup_functions.py:576-629)"""
    from .forward_models import EnsembleContactsFWM
    from .likelihoods import Likelihood
    data = parse_data(data_file, data_filtering_params)
    cd_factor = float(forward_model_params["contact_distance_factor"])
    contact_distances = (bead_radii[data[:, 0]] + bead_radii[data[:, 1]]) * cd_factor
    FWM = EnsembleContactsFWM("fwm", n_structures, contact_distances, data_points=data)
    if error_model == "poisson":
        from .error_models import PoissonEM
        EM = PoissonEM("ensemble_contacts_em", data[:, 2])
    else:
        raise NotImplementedError(error_model)
    L = Likelihood("ensemble_contacts", FWM, EM, 1.0)
    return L.conditional_factory(smooth_steepness=forward_model_params["alpha"])


def setup_default_re_master(n_replicas, sim_path, comm=None):
    """Replica-exchange driver writing the reference's output layout under ``sim_path``
    (samples/, statistics/mcmc_stats.txt, statistics/re_stats.txt, works/).  ``comm`` is unused:
    ranks talk through torch.distributed (reference: setup_functions.py:631-705)."""
    from .rex import ExchangeMaster
    for sub in ("statistics", "works", "samples"):
        os.makedirs(os.path.join(sim_path, sub), exist_ok=True)
    return ExchangeMaster(n_replicas, sim_path)
