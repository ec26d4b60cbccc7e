"""Host-side mirror of the reference's Julia interface for the hot path.

Julia is not installed in this image, so the reference-facing host layer is restated in Python with
the SAME function names, argument meaning, defaults and error behaviour as the Julia functions it
mirrors (cited per function); INTEGRATION.md shows the Julia `ccall` shim that binds the same C
ABI.  Everything numeric goes through `Context` -> libgpar_b200.so; nothing here computes a kernel
matrix, a Cholesky or a filter on the CPU.
"""
import numpy as np
from . import _ffi
from .context import Context
from . import neldermead
from . import lbfgs

# ---- kernel structures (Stheno's EQ(), Matern12(), Matern32(), Matern52()) ---------------------


class Kernel:
    code = None

    def __repr__(self):
        return type(self).__name__ + "()"


class EQ(Kernel):
    code = _ffi.EQ


class Matern12(Kernel):
    code = _ffi.MATERN12


class Matern32(Kernel):
    code = _ffi.MATERN32


class Matern52(Kernel):
    code = _ffi.MATERN52


class ScaledKernel(Kernel):
    """`kernel(k; l, s)` = s * stretch(k, 1/l) (Stheno; cf. optimized.jl:30-31)."""

    def __init__(self, base, l, s):
        self.base, self.l, self.s = base, float(l), float(s)
        self.code = base.code


def kernel(k, l=1.0, s=1.0):
    return ScaledKernel(k, l, s)


class GP:
    """`GP(kernel, GPC())`: zero-mean prior; calling it on inputs gives a FiniteGP."""

    def __init__(self, k):
        self.kernel = k if isinstance(k, ScaledKernel) else ScaledKernel(k, 1.0, 1.0)

    def __call__(self, x, noise=1e-18):
        return FiniteGP(self, to_ColVecs(x), float(noise))


class FiniteGP:
    def __init__(self, gp, x, noise):
        self.gp, self.x, self.noise = gp, x, noise

    def __len__(self):
        return self.x.shape[0]


_default_ctx = {}


def default_context(device=0):
    """One lazily-created Context per device for the functional API."""
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---- src/util.jl ------------------------------------------------------------------------------
def to_ColVecs(inputs):
    """util.jl:16-31.  A list of per-feature vectors (or an (N, D) array of records, or a 1-D vector)
    -> (N, D) C-contiguous records = memory of the reference's D x N column-major ColVecs."""
    if isinstance(inputs, np.ndarray):
        a = np.asarray(inputs, dtype=np.float64)
        return np.ascontiguousarray(a.reshape(-1, 1) if a.ndim == 1 else a)
    if len(inputs) and np.ndim(inputs[0]) == 0:
        return np.ascontiguousarray(np.asarray(inputs, dtype=np.float64).reshape(-1, 1))
    return np.ascontiguousarray(np.stack([np.asarray(c, dtype=np.float64).ravel() for c in inputs], axis=1))


def unpack_gp(params):
    """util.jl:36-43."""
    return tuple(float(np.exp(params[i]) + 1e-3) for i in range(3))


def unpack_gpar(params):
    """util.jl:45-55."""
    return tuple(float(np.exp(params[i]) + 1e-3) for i in range(5))


def get_time_mask(input_length):
    """util.jl:102-106."""
    m = np.zeros(input_length); m[0] = 1.0
    return m


def get_output_mask(input_length):
    """util.jl:111-123 (DomainError for input_length <= 1)."""
    if input_length <= 1:
        raise ValueError("DomainError: Input length must be integer greater than 1")
    m = np.zeros((input_length - 1, input_length))
    for r in range(input_length - 1):
        m[r, r + 1] = 1.0
    return m


def _parse_param(p, rng=None):
    """util.jl:128-134: a missing parameter is drawn with rand() — uniform [0, 1)."""
    if p is None:
        return float((rng or np.random).random())
    return float(p)


def parse_initial_gp_params(i_log_l, i_log_process_var, i_log_noise_sigma, rng=None):
    """util.jl:141-147."""
    return np.array([_parse_param(i_log_l, rng), _parse_param(i_log_process_var, rng), _parse_param(i_log_noise_sigma, rng)])


def parse_initial_gpar_params(i_log_time_l, i_log_time_var, i_log_out_l, i_log_out_var, i_log_noise_sigma, rng=None):
    """util.jl:154-169."""
    return np.array([_parse_param(p, rng) for p in (i_log_time_l, i_log_time_var, i_log_out_l, i_log_out_var, i_log_noise_sigma)])


def _log_pos(v):
    """inverse of exp(p) + 1e-3 for passing positive values through theta-taking entry points"""
    return np.log(np.asarray(v, dtype=np.float64) - 1e-3)


# ---- src/gp/optimized.jl ------------------------------------------------------------------------
class Posterior:
    """`gp | (gp(x, sigma^2) <- y)` (optimized.jl:94,236): callable data holder with the two
    operations the reference's callers use, `marginals` and `mean` (eeg.jl:185-208)."""

    def __init__(self, ctx, X, y, theta, k_time, k_out):
        self.ctx, self.X, self.y, self.theta, self.k_time, self.k_out = ctx, X, np.asarray(y, dtype=np.float64), np.asarray(theta), k_time, k_out

    def _post(self, Xs):
        self.ctx.set_inputs(self.X); self.ctx.set_outputs(self.y)
        mean, var = self.ctx.exact_posterior(self.k_time.code, self.k_out.code, self.theta, to_ColVecs(Xs))
        return mean[0], var

    def marginals(self, Xs):
        """-> (mean, std): Stheno's marginals are Normal(mu, sqrt(var))."""
        mean, var = self._post(Xs)
        return mean, np.sqrt(np.maximum(var, 0.0))

    def mean(self, Xs):
        return self._post(Xs)[0]


def create_optim_gp(input_locations, outputs, kernel_structure=None, i_log_l=None, i_log_process_var=None,
                    i_log_noise_sigma=None, debug=False, ctx=None, rng=None):
    """optimized.jl:19-59 -> (gp, opt_params)."""
    kernel_structure = kernel_structure or EQ()
    ctx = ctx or default_context()
    X = to_ColVecs(input_locations)
    ctx.set_inputs(X); ctx.set_outputs(outputs)

    def nlml(params):   # optimized.jl:28-36
        return -ctx.exact_logpdf(kernel_structure.code, kernel_structure.code, params)[0]

    params = parse_initial_gp_params(i_log_l, i_log_process_var, i_log_noise_sigma, rng)
    if debug:
        print("Generating GP with initial parameters:\n\tl=%s; var=%s; noise=%s" % unpack_gp(params))
    results = neldermead.optimize(nlml, params)
    opt_params = unpack_gp(results.minimizer)
    if debug:
        print("Finished optimizing parameters:\n\tOptimum L: %s \n\tOptimum Process Variance: %s\n\tOptimum noise: %s\n" % opt_params)
    gp = GP(kernel(kernel_structure, l=opt_params[0], s=opt_params[1] ** 2))
    gp._theta = results.minimizer
    return gp, opt_params


def create_optim_gp_post(input_locations, outputs, kernel_structure=None, i_log_l=None, i_log_process_var=None,
                         i_log_noise_sigma=None, debug=False, ctx=None, rng=None):
    """optimized.jl:76-97."""
    kernel_structure = kernel_structure or EQ()
    ctx = ctx or default_context()
    gp, opt_params = create_optim_gp(input_locations, outputs, kernel_structure, i_log_l, i_log_process_var, i_log_noise_sigma, debug, ctx, rng)
    return Posterior(ctx, to_ColVecs(input_locations), outputs, _log_pos(opt_params), kernel_structure, kernel_structure)


def create_optim_gpar(input_locations, outputs, time_kernel=None, out_kernel=None, i_log_time_l=None, i_log_time_var=None,
                      i_log_out_l=None, i_log_out_var=None, i_log_noise_sigma=None, multi_input=True, debug=False, ctx=None, rng=None):
    """optimized.jl:106-183 -> (gpar, opt_params)."""
    time_kernel = time_kernel or EQ(); out_kernel = out_kernel or EQ()
    if not multi_input:
        return create_optim_gp(input_locations, outputs, time_kernel, i_log_time_l, i_log_time_var, i_log_noise_sigma, debug, ctx, rng)
    ctx = ctx or default_context()
    X = to_ColVecs(input_locations)
    if X.shape[1] <= 1:
        get_output_mask(X.shape[1])     # raises like util.jl:112-117
    ctx.set_inputs(X); ctx.set_outputs(outputs)

    def nlml(params):   # optimized.jl:147-154
        return -ctx.exact_logpdf(time_kernel.code, out_kernel.code, params)[0]

    params = parse_initial_gpar_params(i_log_time_l, i_log_time_var, i_log_out_l, i_log_out_var, i_log_noise_sigma, rng)
    results = neldermead.optimize(nlml, params)
    opt_params = unpack_gpar(results.minimizer)
    if debug:
        print("Finished optimizing parameters:\n\tOptimum time L: %s \n\tOptimum time var: %s\n\tOptimum outputs l: %s\n"
              "\tOptimum outputs var: %s\n\tOptimum Noise std: %s\n" % opt_params)
    return ("gpar", time_kernel, out_kernel, opt_params), opt_params


def create_optim_gpar_post(input_locations, outputs, time_kernel=None, out_kernel=None, i_log_time_l=None, i_log_time_var=None,
                           i_log_out_l=None, i_log_out_var=None, i_log_noise_sigma=None, multi_input=True, debug=False, ctx=None, rng=None):
    """optimized.jl:201-239."""
    time_kernel = time_kernel or EQ(); out_kernel = out_kernel or EQ()
    if not multi_input:
        return create_optim_gp_post(input_locations, outputs, time_kernel, i_log_time_l, i_log_time_var, i_log_noise_sigma, debug, ctx, rng)
    ctx = ctx or default_context()
    _, opt_params = create_optim_gpar(input_locations, outputs, time_kernel, out_kernel, i_log_time_l, i_log_time_var, i_log_out_l,
                                      i_log_out_var, i_log_noise_sigma, multi_input, debug, ctx, rng)
    return Posterior(ctx, to_ColVecs(input_locations), outputs, _log_pos(opt_params), time_kernel, out_kernel)


# ---- src/gp/temporal_gp_inference.jl ---------------------------------------------------------
class LGSSM:
    """Handle returned by create_lgssm: the model is rebuilt on the device from these values."""

    def __init__(self, locations, l, process_var, noise_sigma, kernel_structure, noise_vector=None):
        self.t = np.asarray(locations, dtype=np.float64)
        self.l, self.process_var, self.noise_sigma = float(l), float(process_var), float(noise_sigma)
        self.kernel_structure, self.noise_vector = kernel_structure, noise_vector

    @property
    def theta(self):
        return _log_pos([self.l, self.process_var, self.noise_sigma])


def create_lgssm(latent_locations, l, process_var, noise_sigma, kernel_structure, noise_vector=None, debug=False):
    """temporal_gp_inference.jl:15-39."""
    if kernel_structure.code == _ffi.EQ:
        raise ValueError("EQ has no finite-dimensional SDE form (TemporalGPs.to_sde has no method for EQ)")
    return LGSSM(latent_locations, l, process_var, noise_sigma, kernel_structure, noise_vector)


def logpdf(lgssm, y, ctx=None):
    """`logpdf(lgssm, y)` (temporal_gp_inference.jl:78)."""
    ctx = ctx or default_context()
    ctx.set_times(lgssm.t); ctx.set_outputs(y); ctx.set_noise_vector(lgssm.noise_vector)
    out = ctx.lgssm_logpdf(lgssm.kernel_structure.code, lgssm.theta)
    ctx.set_noise_vector(None)
    return out[0] if np.ndim(y) == 1 else out


def decorrelate(lgssm, y, ctx=None):
    """`decorrelate(lgssm, y)` -> (lml, alpha) (dtc.jl:106)."""
    ctx = ctx or default_context()
    ctx.set_times(lgssm.t); ctx.set_outputs(y); ctx.set_noise_vector(lgssm.noise_vector)
    lml, alpha = ctx.lgssm_decorrelate(lgssm.kernel_structure.code, lgssm.theta)
    ctx.set_noise_vector(None)
    return (lml[0], alpha[0]) if np.ndim(y) == 1 else (lml, alpha)


class Gaussian:
    """Element of `smooth`'s output as the reference's callers read it: `.m[1]`, `.P[1]`
    (GPAR_scaled_examples.jl:111,128-129) -> here `.m[0]`, `.P[0]`."""
    __slots__ = ("m", "P")

    def __init__(self, m, P):
        self.m, self.P = (m,), (P,)


def smooth(lgssm, y, ctx=None):
    """`smooth(lgssm, y)` -> (None, y_smooth, lml): the reference only reads the 2nd return."""
    ctx = ctx or default_context()
    ctx.set_times(lgssm.t); ctx.set_outputs(y); ctx.set_noise_vector(lgssm.noise_vector)
    lml, mean, var = ctx.lgssm_smooth(lgssm.kernel_structure.code, lgssm.theta)
    ctx.set_noise_vector(None)
    return None, [Gaussian(m, v) for m, v in zip(mean[0], var[0])], lml[0]


def get_sde_predictions(data_locations, data_outputs, output_locations, kernel_structure=None, i_log_time_l=None, i_log_time_var=None,
                        i_log_noise_sigma=None, debug=True, ctx=None, rng=None, return_arrays=False, optimizer="neldermead",
                        device_merge=False, speculative=False):
    """temporal_gp_inference.jl:45-114 -> (opt_lgssm, output_observations).  optimizer="lbfgs" replaces the
    reference's Nelder-Mead (:82) by L-BFGS on the library's analytic gradient."""
    kernel_structure = kernel_structure or Matern52()
    ctx = ctx or default_context()
    data_locations = np.asarray(data_locations, dtype=np.float64); output_locations = np.asarray(output_locations, dtype=np.float64)
    data_outputs = np.asarray(data_outputs, dtype=np.float64)
    latent_locations = np.concatenate([data_locations, output_locations])                      # :56
    outputs = np.concatenate([data_outputs, np.zeros(len(output_locations))])                  # :58
    sorting_perm = np.argsort(latent_locations, kind="stable")                                 # :61 (Julia's sortperm is stable)
    reverse_perm = np.argsort(sorting_perm, kind="stable")                                     # :62
    s_latent_locations = latent_locations[sorting_perm]; s_outputs = outputs[sorting_perm]

    ctx.set_times(data_locations); ctx.set_outputs(data_outputs); ctx.set_noise_vector(None)

    def nlml(params):    # :69-79 (the LGSSM is built on data_locations as given)
        return -ctx.lgssm_logpdf(kernel_structure.code, params)[0]

    params = parse_initial_gp_params(i_log_time_l, i_log_time_var, i_log_noise_sigma, rng)
    if optimizer == "lbfgs":
        def nlml_fg(p_):
            v, g = ctx.lgssm_logpdf_grad(kernel_structure.code, p_)
            return -v[0], -g[0]
        results = lbfgs.optimize(nlml_fg, params)
    elif speculative:    # the same run; the candidate points of an iteration in ONE pass over the sequence (candidates of gpar_lgssm_logpdf)
        results = neldermead.optimize_speculative(lambda P: -ctx.lgssm_logpdf(kernel_structure.code, P), params)
    else:
        results = neldermead.optimize(nlml, params)                                            # :82
    opt_l, opt_process_var, opt_noise_sigma = unpack_gp(results.minimizer)
    if debug:
        print("Finished optimizing parameters:\n\tOptimum L: %s \n\tOptimum Process Variance: %s\n\tOptimum noise: %s\n"
              % (opt_l, opt_process_var, opt_noise_sigma))
    noise_vector = np.concatenate([np.full(len(data_locations), opt_noise_sigma ** 2), np.full(len(output_locations), 1e10)])   # :93-96
    s_noise_vector = noise_vector[sorting_perm]
    opt_lgssm = create_lgssm(s_latent_locations, opt_l, opt_process_var, opt_noise_sigma, kernel_structure, noise_vector=s_noise_vector)
    if device_merge:     # :55-66, :93-97, :111-112 on the device (gpar_set_merged / gpar_take_test)
        ctx.set_merged(data_locations, data_outputs, output_locations, opt_noise_sigma ** 2)
        ctx.lgssm_smooth(kernel_structure.code, results.minimizer, keep_on_device=True)
        mean, var = ctx.take_test()
        ctx.set_noise_vector(None)
    else:
        ctx.set_times(s_latent_locations); ctx.set_outputs(s_outputs); ctx.set_noise_vector(s_noise_vector)
        _, mean, var = ctx.lgssm_smooth(kernel_structure.code, results.minimizer)              # :109
        ctx.set_noise_vector(None)
        nd = len(data_outputs)
        mean = mean[0][reverse_perm][nd:]; var = var[0][reverse_perm][nd:]                     # :111-112
    if return_arrays:
        return opt_lgssm, (mean, var)
    return opt_lgssm, [Gaussian(m, v) for m, v in zip(mean, var)]


# ---- src/gp/dtc.jl ---------------------------------------------------------------------------
def compute_gpar_dtc_objective(f, u, time_loc, outputs, time_kernel=None, temporal_noise_sigma=0.04, ctx=None, return_A=True):
    """dtc.jl:83-128 -> (dtc, A).  f, u: FiniteGPs of one prior at the inputs / pseudo-inputs;
    cov(u) includes u's noise as jitter (dtc.jl:35,119); time_kernel a (scaled) Matern kernel."""
    ctx = ctx or default_context()
    tk = time_kernel if isinstance(time_kernel, ScaledKernel) else ScaledKernel(time_kernel or Matern52(), 1.0, 1.0)
    ok = f.gp.kernel
    if abs(u.noise - temporal_noise_sigma ** 2) > 1e-15 * max(u.noise, 1e-300):
        raise ValueError("the device path implements the reference's call pattern cov(u) = Kuu + temporal_noise_sigma^2 I "
                         "(dtc.jl:34-35,44): u's noise must equal temporal_noise_sigma^2")
    ctx.set_inputs(f.x); ctx.set_pseudo(u.x); ctx.set_times(time_loc); ctx.set_outputs(outputs)
    theta = _log_pos([tk.l, np.sqrt(tk.s), ok.l, np.sqrt(ok.s), temporal_noise_sigma])
    if return_A:
        return ctx.scaled_dtc(tk.code, ok.code, theta, return_A=True)
    return ctx.scaled_dtc(tk.code, ok.code, theta), None


def get_optim_scaled_gpar_params(input_locations, pseudo_input_locations, time_loc, outputs, out_kernel=None, time_kernel=None,
                                 i_log_time_l=None, i_log_time_var=None, i_log_out_l=None, i_log_out_var=None, i_log_noise_sigma=None,
                                 optimization_time_limit=1000.0, show_optimization_trace=False, debug=False, ctx=None, rng=None,
                                 iterations=1000, return_result=False, optimizer="neldermead", n_restarts=1, speculative=False, group=None):
    """dtc.jl:11-77 -> (time_l, time_var, out_l, out_var, noise_sigma).  group (NEW; a context.Group): the ROWS of this one
    objective are sharded over the group's devices (gpar_group_scaled_dtc_sharded) — for a single output too large for one
    device's time or memory budget; plain Nelder-Mead or optimizer="lbfgs" (sharded gradient).
    A failed Cholesky: the plain Nelder-Mead path raises PosDefException, as the reference's nlml (dtc.jl:29-48, no
    try / catch) would; the NEW paths (lbfgs, n_restarts, speculative, the C++ group fits) read it as +Inf for that point.
    optimizer="lbfgs" replaces the
    reference's Nelder-Mead (:58-61) by L-BFGS on gpar_scaled_dtc_grad.  n_restarts > 1 (NEW): that many Nelder-Mead
    runs in lock-step — the first from the given / drawn initial parameters, the others from theta0 ~ U(0,1)^5 (the
    missing-parameter rule, util.jl:128-134) — every round of candidates in ONE gpar_scaled_dtc_batch call; each run
    performs exactly the operations of a separate Optim.optimize, the best optimum is returned.  speculative (NEW): one
    run whose four candidate points per iteration are evaluated in one batched call (neldermead.optimize_speculative):
    the simplices, the optimum and f_calls of the plain run at ~2/3 of its wall-clock at the reference's sizes."""
    out_kernel = out_kernel or Matern52(); time_kernel = time_kernel or Matern52()
    if group is not None:
        if n_restarts > 1 or speculative:
            raise ValueError("group=: the row-sharded objective is driven by plain Nelder-Mead or L-BFGS, one run")
        row_lo = group.load_row_slices(to_ColVecs(input_locations), to_ColVecs(pseudo_input_locations), time_loc, outputs)

        def nlml(params):
            return -group.scaled_dtc_sharded(time_kernel.code, out_kernel.code, params, row_lo)

        class _Sharded:      # what the L-BFGS branch below asks of a context
            @staticmethod
            def scaled_dtc_grad(k_time, k_out, p_):
                return group.scaled_dtc_sharded(k_time, k_out, p_, row_lo, grad=True)
        ctx = _Sharded
    else:
        ctx = ctx or default_context()
        ctx.set_inputs(to_ColVecs(input_locations)); ctx.set_pseudo(to_ColVecs(pseudo_input_locations))
        ctx.set_times(time_loc); ctx.set_outputs(outputs)

        def nlml(params):    # dtc.jl:29-48; data stays resident on the device across evaluations
            return -ctx.scaled_dtc(time_kernel.code, out_kernel.code, params)

    params = parse_initial_gpar_params(i_log_time_l, i_log_time_var, i_log_out_l, i_log_out_var, i_log_noise_sigma, rng)
    if debug:
        print("Generating scaled GPAR with initial parameters:\n\ti_time_l=%s; i_time_var=%s; i_out_l=%s; i_out_var=%s; i_noise_sigma=%s" % unpack_gpar(params))
    if optimizer == "lbfgs":
        def nlml_fg(p_):
            try:
                v, g = ctx.scaled_dtc_grad(time_kernel.code, out_kernel.code, p_)
            except _ffi.PosDefException:
                return np.inf, np.zeros(5)
            return -v, -g
        results = lbfgs.optimize(nlml_fg, params, iterations=iterations, time_limit=optimization_time_limit, show_trace=show_optimization_trace)
    elif n_restarts > 1:
        rng_ = rng if rng is not None else np.random.default_rng()
        X0 = np.vstack([params] + [rng_.random(5) for _ in range(n_restarts - 1)])

        def nlml_batch(P):       # a failed Cholesky (PosDefException of the reference) is +Inf for that vertex only
            vals, codes = ctx.scaled_dtc_batch(time_kernel.code, out_kernel.code, P)
            return np.where(codes == 0, -vals, np.inf)
        runs = neldermead.optimize_batch(nlml_batch, X0, iterations=iterations)
        results = min(runs, key=lambda r_: r_.minimum)
    elif speculative:
        def nlml_batch1(P):
            vals, codes = ctx.scaled_dtc_batch(time_kernel.code, out_kernel.code, P)
            return np.where(codes == 0, -vals, np.inf)
        results = neldermead.optimize_speculative(nlml_batch1, params, iterations=iterations, time_limit=optimization_time_limit,
                                                  show_trace=show_optimization_trace)
    else:
        results = neldermead.optimize(nlml, params, iterations=iterations, time_limit=optimization_time_limit, show_trace=show_optimization_trace)
    opt_params = unpack_gpar(results.minimizer)
    if debug:
        print("Finished optimizing parameters:\n\tOptimum time L: %s \n\tOptimum time var: %s\n\tOptimum outputs l: %s\n"
              "\tOptimum outputs var: %s\n\tOptimum Noise std: %s\n" % opt_params)
    return (opt_params, results) if return_result else opt_params


# ---- src/gp/gpar_scaled_inference.jl ---------------------------------------------------------
def compute_q_u(input_locations, pseudo_input_locations, time_loc, outputs, out_kernel=None, time_kernel=None,
                temporal_noise_sigma=0.05, debug=False, ctx=None):
    """gpar_scaled_inference.jl:141-196 -> ((m_e, inv(D)), U_u): q_u = MvNormal(m_e, inv(D))."""
    ctx = ctx or default_context()
    ok = out_kernel if isinstance(out_kernel, ScaledKernel) else ScaledKernel(out_kernel or Matern52(), 1.0, 1.0)
    tk = time_kernel if isinstance(time_kernel, ScaledKernel) else ScaledKernel(time_kernel or Matern52(), 1.0, 1.0)
    ctx.set_inputs(to_ColVecs(input_locations)); ctx.set_pseudo(to_ColVecs(pseudo_input_locations))
    ctx.set_times(time_loc); ctx.set_outputs(outputs)
    m_e, Dinv, U_u = ctx.compute_q_u(tk.code, ok.code, [tk.l, np.sqrt(tk.s), ok.l, np.sqrt(ok.s), temporal_noise_sigma])
    return (m_e, Dinv), U_u


def get_gpar_scaled_predictions(input_locations, pseudo_input_locations, time_loc, outputs, inference_time_loc,
                                inference_input_locations, out_kernel_structure=None, time_kernel_structure=None,
                                i_log_time_l=None, i_log_time_var=None, i_log_out_l=None, i_log_out_var=None, i_log_noise_sigma=None,
                                optimization_time_limit=1000.0, debug=False, ctx=None, rng=None, iterations=1000, nsamples=100,
                                opt_params=None, sampler="device", seed=0, device_merge=True, n_restarts=1, speculative=False, group=None):
    """gpar_scaled_inference.jl:20-136 -> (inferred_outputs, inferred_stds) at the inference locations.
    group (NEW; a context.Group): the two N x M stages — the fit and compute_q_u — run with their rows sharded over the group's
    devices (gpar_group_scaled_dtc_sharded / gpar_group_sample_q_u_sharded or _compute_q_u_sharded); the prediction itself needs
    no N x M array and stays on `ctx`.
    n_restarts / speculative: passed to get_optim_scaled_gpar_params (batched candidates; NEW).
    `opt_params` (positive 5-tuple) skips the optimisation (used by the chain driver, which fits all
    outputs in parallel first); `rng` seeds the q_u draws the reference takes from Julia's global RNG."""
    out_kernel_structure = out_kernel_structure or Matern52(); time_kernel_structure = time_kernel_structure or Matern52()
    ctx = ctx or default_context()
    rng = rng or np.random.default_rng()
    X = to_ColVecs(input_locations); Z = to_ColVecs(pseudo_input_locations); Xs = to_ColVecs(inference_input_locations)
    time_loc = np.asarray(time_loc, dtype=np.float64); inference_time_loc = np.asarray(inference_time_loc, dtype=np.float64)
    outputs = np.asarray(outputs, dtype=np.float64)
    # sampler="device" (default): the q_u draws come from the library's seeded Philox stream (gpar_sample_q_u, keyed by
    # `seed`) and never leave the GPU, so the product path has no host numerics; sampler="host" draws them here from
    # `rng` instead (the reference uses Julia's unseeded global RNG, :94) — kept for callers that bring their own draws.
    # device_merge (default): the merge / sort / un-sort protocol (:75-87, :100-103, :132-133) runs on the device too.
    if opt_params is None:
        print("Starting optimization")                                                        # :42 (unconditional in the reference)
        opt_params = get_optim_scaled_gpar_params(X, Z, time_loc, outputs, out_kernel=Matern52(), time_kernel=Matern52(),   # :43-56 hard-codes Matern52
                                                  i_log_time_l=i_log_time_l, i_log_time_var=i_log_time_var, i_log_out_l=i_log_out_l,
                                                  i_log_out_var=i_log_out_var, i_log_noise_sigma=i_log_noise_sigma,
                                                  optimization_time_limit=optimization_time_limit, debug=debug, ctx=ctx, rng=rng,
                                                  iterations=iterations, n_restarts=n_restarts, speculative=speculative, group=group)
    opt_time_l, opt_time_var, opt_out_l, opt_out_var, opt_noise_sigma = opt_params
    params = np.array([opt_time_l, opt_time_var, opt_out_l, opt_out_var, opt_noise_sigma])
    # q(u) ~ p(u | y)  (:63-73)
    W_group = None
    if group is not None:
        row_lo = group.load_row_slices(X, Z, time_loc, outputs)
        if sampler == "device":       # seeded Philox draws on the member that holds the summed statistics: no host numerics
            W_group, _ = group.sample_q_u_sharded(time_kernel_structure.code, out_kernel_structure.code, params, row_lo, seed, nsamples)
        else:
            m_e, Dinv, U_u = group.compute_q_u_sharded(time_kernel_structure.code, out_kernel_structure.code, params, row_lo)
        ctx.set_pseudo(Z)
    else:
        ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(time_loc); ctx.set_outputs(outputs); ctx.set_noise_vector(None)
    if group is not None:
        pass                                       # (m_e, inv(D), U_u) came from the sharded evaluation
    elif sampler == "device":
        ctx.sample_q_u(time_kernel_structure.code, out_kernel_structure.code, params, seed, nsamples)
    else:
        m_e, Dinv, U_u = ctx.compute_q_u(time_kernel_structure.code, out_kernel_structure.code, params)
    # merge + sort (:75-87)
    ntr = len(time_loc)
    time_loc_concat = np.concatenate([time_loc, inference_time_loc])
    sorting_perm = np.argsort(time_loc_concat, kind="stable"); reverse_perm = np.argsort(sorting_perm, kind="stable")
    input_loc_star = np.concatenate([X, Xs], axis=0)[sorting_perm]
    outputs_star = np.concatenate([outputs, np.zeros(len(inference_time_loc))])[sorting_perm]
    noise_vector_star = np.concatenate([np.full(ntr, opt_noise_sigma ** 2), np.full(len(inference_time_loc), 1e10)])[sorting_perm]   # :100-103
    W = W_group
    if sampler != "device":
        # draws eps_j ~ MvNormal(m_e, inv(D)) (:94) and the weights U_u \ eps_j (:96) — tiny M x S host work
        Lc = np.linalg.cholesky(0.5 * (Dinv + Dinv.T))
        eps = m_e[:, None] + Lc @ rng.standard_normal((len(m_e), nsamples))
        from scipy.linalg import solve_triangular
        W = solve_triangular(np.triu(U_u), eps, lower=False)
    if device_merge:     # :75-87, :100-103, :132-133 on the device (gpar_set_merged / gpar_take_test)
        ctx.set_merged(time_loc, outputs, inference_time_loc, opt_noise_sigma ** 2, X=X, Xs=Xs)
        ctx.scaled_predict(time_kernel_structure.code, out_kernel_structure.code, params, W, keep_on_device=True)
        mean, std = ctx.take_test()
        ctx.set_noise_vector(None)
        return mean, std
    ctx.set_inputs(input_loc_star); ctx.set_times(time_loc_concat[sorting_perm]); ctx.set_outputs(outputs_star)
    ctx.set_noise_vector(noise_vector_star)
    mean, std = ctx.scaled_predict(time_kernel_structure.code, out_kernel_structure.code, params, W)      # :110-130 batched
    ctx.set_noise_vector(None)
    return mean[reverse_perm][ntr:], std[reverse_perm][ntr:]                                              # :132-133
