"""Thin object wrapper over the C ABI (one `Context` = one gpar_ctx = one device + stream)."""
import ctypes
import numpy as np
from . import _ffi
from ._ffi import as_f64, dptr


class Context:
    """Owns a `gpar_ctx`.  Data set through the `set_*` methods stays resident on the device across
    evaluations, as the optimiser closures of the reference re-evaluate the objective hundreds of
    times on fixed data (src/gp/dtc.jl:29-61)."""

    def __init__(self, device=0, _borrowed=None):
        self._lib = _ffi.load_library()
        self._owned = _borrowed is None
        if _borrowed is not None:       # a member context of a Group: the group owns the handle
            self._h = ctypes.c_void_p(_borrowed)
            self.device = int(device)
            self.N = self.M = self.D = 0
            self.batch = 0
            return
        h = ctypes.c_void_p()
        st = self._lib.gpar_ctx_create(int(device), ctypes.byref(h))
        if st != _ffi.GPAR_OK:
            raise _ffi.GparError(st, "gpar_ctx_create(device=%d) failed (no usable CUDA device?)" % device)
        self._h = h
        self.device = int(device)
        self.N = self.M = self.D = 0
        self.batch = 0

    # -- plumbing -------------------------------------------------------------------------------
    def _check(self, st):
        if st == _ffi.GPAR_OK:
            return
        msg = self._lib.gpar_last_error(self._h).decode("utf-8", "replace")
        if st == _ffi.GPAR_ERR_NOT_POSDEF:
            raise _ffi.PosDefException(st, msg)
        raise _ffi.GparError(st, msg)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            if getattr(self, "_owned", True):
                self._lib.gpar_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def last_timing(self):
        ms = ctypes.c_double()
        n = ctypes.c_int64()
        self._check(self._lib.gpar_last_timing(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def last_profile(self):
        """(producer_ms, dmma_syrk_kernel_ms, tail_ms) of the last pseudo-point call."""
        out = np.zeros(3)
        self._check(self._lib.gpar_last_profile(self._h, dptr(out), 3))
        return tuple(out.tolist())

    def measure_peaks(self):
        """-> {"dmma_tflops", "dfma_tflops", "hbm_copy_gbs"} measured on this context's device (roofline denominators)."""
        a = ctypes.c_double(); b = ctypes.c_double(); c = ctypes.c_double()
        self._check(self._lib.gpar_measure_peaks(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"dmma_tflops": a.value, "dfma_tflops": b.value, "hbm_copy_gbs": c.value}

    def dense_bench(self, n):
        """Device ms of the dense routines at order n: potrf, trtri, triangular / lower-tile / full gemm, trsv, trsv^T."""
        out = np.zeros(7)
        self._check(self._lib.gpar_dense_bench(self._h, int(n), dptr(out)))
        return dict(zip(["potrf", "trtri", "gemm_tri", "gemm_lower", "gemm_full", "trsv", "trsv_t"], out.tolist()))

    # -- resident data ------------------------------------------------------------------------
    def set_inputs(self, X):
        """X: (N, D) records (memory image of the reference's D x N ColVecs, util.jl:16-31)."""
        X = as_f64(np.atleast_2d(X) if np.ndim(X) > 1 else np.asarray(X, dtype=np.float64).reshape(-1, 1))
        self._check(self._lib.gpar_set_inputs(self._h, dptr(X), X.shape[1], X.shape[0]))
        self.N, self.D = X.shape

    def set_inputs_column(self, d, col=None):
        """Column d of the resident inputs <- col, or (col=None) the chain buffer a Group.broadcast filled."""
        c = None if col is None else as_f64(np.asarray(col).ravel())
        self._check(self._lib.gpar_set_inputs_column(self._h, int(d), dptr(c)))

    def set_merged_test_column(self, d, col=None):
        """Column d of the merged inputs at the test locations <- col (N* values, test order), or the chain buffer."""
        c = None if col is None else as_f64(np.asarray(col).ravel())
        self._check(self._lib.gpar_set_merged_test_column(self._h, int(d), dptr(c)))

    def set_pseudo(self, Z):
        Z = as_f64(np.atleast_2d(Z) if np.ndim(Z) > 1 else np.asarray(Z, dtype=np.float64).reshape(-1, 1))
        self._check(self._lib.gpar_set_pseudo(self._h, dptr(Z), Z.shape[1], Z.shape[0]))
        self.M = Z.shape[0]

    def set_times(self, t):
        t = as_f64(np.asarray(t).ravel())
        self._check(self._lib.gpar_set_times(self._h, dptr(t), t.shape[0]))
        self.Nt = t.shape[0]

    def set_times_range(self, t0, dt, n):
        """The regular grid t0 + k dt (Julia: an AbstractRange) — constant transition matrix per sequence."""
        self._check(self._lib.gpar_set_times_range(self._h, float(t0), float(dt), int(n)))
        self.Nt = int(n)

    def set_outputs(self, y):
        """y: (N,) or (batch, N) — sequence b contiguous."""
        y = as_f64(y)
        if y.ndim == 1:
            y = y[None, :]
        self._check(self._lib.gpar_set_outputs(self._h, dptr(y), y.shape[1], y.shape[0]))
        self.batch, self.Ny = y.shape

    def set_noise_vector(self, r):
        if r is None:
            self._check(self._lib.gpar_set_noise_vector(self._h, None, 0))
            return
        r = as_f64(np.asarray(r).ravel())
        self._check(self._lib.gpar_set_noise_vector(self._h, dptr(r), r.shape[0]))

    def set_merged(self, t, y, ts, sigma2, X=None, Xs=None):
        """Device merge/sort protocol (gpar_set_merged): train (t, y[, X]) + test (ts[, Xs]) -> resident sorted
        times / outputs (0 at test points) / noise vector (sigma2 | 1e10) / inputs."""
        t = as_f64(np.asarray(t).ravel()); y = as_f64(np.asarray(y).ravel()); ts = as_f64(np.asarray(ts).ravel())
        D = 0
        if X is not None:
            X = as_f64(np.atleast_2d(X) if np.ndim(X) > 1 else np.asarray(X, dtype=np.float64).reshape(-1, 1))
            Xs = as_f64(np.atleast_2d(Xs) if np.ndim(Xs) > 1 else np.asarray(Xs, dtype=np.float64).reshape(-1, 1))
            D = X.shape[1]
        self._check(self._lib.gpar_set_merged(self._h, dptr(t), dptr(y), dptr(X), t.shape[0], dptr(ts), dptr(Xs), ts.shape[0], D, float(sigma2)))
        self.Nt = self.Ny = t.shape[0] + ts.shape[0]; self.batch = 1; self._merged_ns = ts.shape[0]
        if D:
            self.N, self.D = self.Nt, D

    def take_test(self, two=True):
        """(a, b) of the last smoother / prediction at the test locations of the merged problem, in test order."""
        a = np.zeros(self._merged_ns); b = np.zeros(self._merged_ns) if two else None
        self._check(self._lib.gpar_take_test(self._h, dptr(a), dptr(b)))
        return (a, b) if two else a

    # -- compute --------------------------------------------------------------------------------
    def dtc_logpdf(self, kernel, theta, vfe=False, jitter=-1.0, grad=False):
        th = as_f64(np.asarray(theta).ravel())
        val = ctypes.c_double()
        g = np.zeros(3) if grad else None
        self._check(self._lib.gpar_dtc_logpdf(self._h, int(kernel), dptr(th), int(bool(vfe)), float(jitter),
                                              ctypes.byref(val), dptr(g)))
        return (val.value, g) if grad else val.value

    def dtc_logpdf_zgrad(self, kernel, theta, vfe=False, jitter=-1.0):
        """-> (value, d/dtheta (3,), d/dZ (M, D)): pseudo-input gradients of the DTC / VFE objective."""
        th = as_f64(np.asarray(theta).ravel())
        val = ctypes.c_double(); g = np.zeros(3); gz = np.zeros((self.M, self.D))
        self._check(self._lib.gpar_dtc_logpdf_zgrad(self._h, int(kernel), dptr(th), int(bool(vfe)), float(jitter),
                                                    ctypes.byref(val), dptr(g), dptr(gz)))
        return val.value, g, gz

    def scaled_dtc(self, k_time, k_out, theta, return_A=False):
        th = as_f64(np.asarray(theta).ravel())
        val = ctypes.c_double()
        A = np.zeros((self.M, self.N), order="F") if return_A else None
        self._check(self._lib.gpar_scaled_dtc(self._h, int(k_time), int(k_out), dptr(th), ctypes.byref(val), dptr(A)))
        return (val.value, A) if return_A else val.value

    def scaled_dtc_batch(self, k_time, k_out, thetas):
        """thetas: (ncand, 5) candidates on the resident data, evaluated concurrently on the context's lanes ->
        (dtc (ncand,), codes (ncand,): non-zero where a Cholesky failed, value NaN)."""
        th = as_f64(np.atleast_2d(thetas))
        out = np.zeros(th.shape[0]); codes = np.zeros(th.shape[0], dtype=np.int32)
        self._check(self._lib.gpar_scaled_dtc_batch(self._h, int(k_time), int(k_out), dptr(th), th.shape[0], dptr(out),
                                                    codes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return out, codes

    def scaled_dtc_grad(self, k_time, k_out, theta):
        """-> (dtc, d dtc / d theta (5,))."""
        th = as_f64(np.asarray(theta).ravel())
        val = ctypes.c_double()
        g = np.zeros(5)
        self._check(self._lib.gpar_scaled_dtc_grad(self._h, int(k_time), int(k_out), dptr(th), ctypes.byref(val), dptr(g)))
        return val.value, g

    # ---- one row slice of a scaled objective, collectives owned by the caller (parallel.scaled_dtc_row_sharded) ----
    # buffers: anything with .data_ptr() (torch CUDA tensors on this context's device) or raw device addresses
    @staticmethod
    def _devptr(buf):
        return ctypes.c_void_p(buf.data_ptr() if hasattr(buf, "data_ptr") else int(buf))

    def scaled_slice_begin(self, k_time, k_out, theta, row_lo, grad=False):
        """The context holds the FULL (t, y), Z and rows [row_lo, row_lo + N) of X -> (summary_count, stats_count)."""
        th = as_f64(np.asarray(theta).ravel())
        sc = ctypes.c_int64(); stc = ctypes.c_int64()
        self._check(self._lib.gpar_scaled_slice_begin(self._h, int(k_time), int(k_out), dptr(th), int(row_lo), int(bool(grad)),
                                                      ctypes.byref(sc), ctypes.byref(stc)))
        return sc.value, stc.value

    def scaled_slice_summary(self, summary):
        self._check(self._lib.gpar_scaled_slice_summary(self._h, self._devptr(summary)))

    def scaled_slice_stats(self, gathered, member, stats):
        self._check(self._lib.gpar_scaled_slice_stats(self._h, self._devptr(gathered), int(member), self._devptr(stats)))

    def scaled_slice_value(self, stats):
        val = ctypes.c_double()
        self._check(self._lib.gpar_scaled_slice_value(self._h, self._devptr(stats), ctypes.byref(val)))
        return val.value

    def scaled_slice_tangent_summary(self, stats, summary2):
        self._check(self._lib.gpar_scaled_slice_tangent_summary(self._h, self._devptr(stats), self._devptr(summary2)))

    def scaled_slice_grad_partial(self, gathered2, member):
        s5 = np.zeros(5)
        self._check(self._lib.gpar_scaled_slice_grad_partial(self._h, self._devptr(gathered2), int(member), dptr(s5)))
        return s5

    def scaled_slice_grad_finish(self, s5_total):
        s5 = as_f64(np.asarray(s5_total).ravel()); val = ctypes.c_double(); g = np.zeros(5)
        self._check(self._lib.gpar_scaled_slice_grad_finish(self._h, dptr(s5), ctypes.byref(val), dptr(g)))
        return val.value, g

    def compute_q_u(self, k_time, k_out, params):
        p = as_f64(np.asarray(params).ravel())
        m_e = np.zeros(self.M)
        Dinv = np.zeros((self.M, self.M), order="F")
        U_u = np.zeros((self.M, self.M), order="F")
        self._check(self._lib.gpar_compute_q_u(self._h, int(k_time), int(k_out), dptr(p), dptr(m_e), dptr(Dinv), dptr(U_u)))
        return m_e, Dinv, U_u

    def sample_q_u(self, k_time, k_out, params, seed, nsamples, return_host=False):
        """Seeded device draws from q_u; the weights W = U_u \\ eps stay resident for scaled_predict(W=None).
        return_host -> (W, eps), each (M, S)."""
        p = as_f64(np.asarray(params).ravel())
        W = np.zeros((self.M, nsamples), order="F") if return_host else None
        E = np.zeros((self.M, nsamples), order="F") if return_host else None
        self._check(self._lib.gpar_sample_q_u(self._h, int(k_time), int(k_out), dptr(p), int(seed), int(nsamples), dptr(W), dptr(E)))
        self._nsamples_resident = int(nsamples)
        return (W, E) if return_host else None

    def scaled_predict(self, k_time, k_out, params, W=None, keep_on_device=False):
        """W: (M, S) — column j = U_u \\ eps_j; None: the resident weights of the last sample_q_u.
        -> (mean, std) over the merged, sorted locations (keep_on_device: nothing, use take_test())."""
        p = as_f64(np.asarray(params).ravel())
        if keep_on_device:
            Wf = None if W is None else np.asfortranarray(W, dtype=np.float64)
            S = self._nsamples_resident if W is None else Wf.shape[1]
            self._check(self._lib.gpar_scaled_predict(self._h, int(k_time), int(k_out), dptr(p), dptr(Wf), S, None, None))
            return None
        mean = np.zeros(self.N); sd = np.zeros(self.N)
        if W is None:
            self._check(self._lib.gpar_scaled_predict(self._h, int(k_time), int(k_out), dptr(p), None, self._nsamples_resident, dptr(mean), dptr(sd)))
            return mean, sd
        W = np.asfortranarray(W, dtype=np.float64)
        self._check(self._lib.gpar_scaled_predict(self._h, int(k_time), int(k_out), dptr(p), W.ctypes.data_as(_ffi._c_double_p),
                                                  W.shape[1], dptr(mean), dptr(sd)))
        return mean, sd

    def lgssm_logpdf(self, kernel, theta):
        """theta: (3,) shared, (batch, 3) one per resident sequence, or — with ONE resident sequence — (ncand, 3)
        hyper-parameter candidates evaluated in one pass -> lml per sequence / candidate."""
        th = as_f64(np.atleast_2d(theta))
        out = np.zeros(max(self.batch, th.shape[0]))
        self._check(self._lib.gpar_lgssm_logpdf(self._h, int(kernel), dptr(th), th.shape[0], dptr(out)))
        return out

    def lgssm_logpdf_grad(self, kernel, theta):
        """-> (lml (batch,), grad (batch, 3)): d lml[b] / d theta of the parameter set sequence b uses."""
        th = as_f64(np.atleast_2d(theta))
        out = np.zeros(self.batch)
        grad = np.zeros((self.batch, 3))
        self._check(self._lib.gpar_lgssm_logpdf_grad(self._h, int(kernel), dptr(th), th.shape[0], dptr(out), dptr(grad)))
        return out, grad

    def lgssm_decorrelate(self, kernel, theta):
        th = as_f64(np.asarray(theta).ravel())
        alpha = np.zeros((self.batch, self.Ny))
        lml = np.zeros(self.batch)
        self._check(self._lib.gpar_lgssm_decorrelate(self._h, int(kernel), dptr(th), dptr(alpha), dptr(lml)))
        return lml, alpha

    def lgssm_smooth(self, kernel, theta, keep_on_device=False):
        th = as_f64(np.asarray(theta).ravel())
        if keep_on_device:      # results stay resident for take_test()
            lml = np.zeros(self.batch)
            self._check(self._lib.gpar_lgssm_smooth(self._h, int(kernel), dptr(th), None, None, dptr(lml)))
            return lml
        mean = np.zeros((self.batch, self.Ny))
        var = np.zeros((self.batch, self.Ny))
        lml = np.zeros(self.batch)
        self._check(self._lib.gpar_lgssm_smooth(self._h, int(kernel), dptr(th), dptr(mean), dptr(var), dptr(lml)))
        return lml, mean, var

    def exact_logpdf(self, k_time, k_out, theta):
        th = as_f64(np.asarray(theta).ravel())
        out = np.zeros(self.batch)
        self._check(self._lib.gpar_exact_logpdf(self._h, int(k_time), int(k_out), dptr(th), th.shape[0], dptr(out)))
        return out

    def exact_logpdf_batch(self, k_time, k_out, thetas):
        """thetas: (ncand, 3 or 5) hyper-parameter candidates on the resident data, one launch -> (lml (ncand, batch),
        codes (ncand,): non-zero where a candidate's Cholesky failed, its lml then NaN)."""
        th = as_f64(np.atleast_2d(thetas))
        out = np.zeros((th.shape[0], self.batch)); codes = np.zeros(th.shape[0], dtype=np.int32)
        self._check(self._lib.gpar_exact_logpdf_batch(self._h, int(k_time), int(k_out), dptr(th), th.shape[1], th.shape[0], dptr(out),
                                                      codes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return out, codes

    def exact_posterior(self, k_time, k_out, theta, Xs):
        th = as_f64(np.asarray(theta).ravel())
        Xs = as_f64(np.atleast_2d(Xs) if np.ndim(Xs) > 1 else np.asarray(Xs, dtype=np.float64).reshape(-1, 1))
        mean = np.zeros((self.batch, Xs.shape[0]))
        var = np.zeros(Xs.shape[0])
        self._check(self._lib.gpar_exact_posterior(self._h, int(k_time), int(k_out), dptr(th), th.shape[0],
                                                   dptr(Xs), Xs.shape[0], dptr(mean), dptr(var)))
        return mean, var


class Group:
    """Several devices of one box from ONE process (gpar_group_*, include/gpar_b200.h): one member Context per
    device; concurrent objective evaluations, whole Nelder-Mead fits of a task list, NCCL all-gather of the scalars
    and broadcast of posterior means down the GPAR chain."""

    def __init__(self, devices):
        self._lib = _ffi.load_library()
        devs = (ctypes.c_int32 * len(devices))(*[int(d) for d in devices])
        h = ctypes.c_void_p()
        st = self._lib.gpar_group_create(devs, len(devices), ctypes.byref(h))
        if st != _ffi.GPAR_OK:
            raise _ffi.GparError(st, "gpar_group_create(%s) failed (devices / NCCL unavailable?)" % (list(devices),))
        self._h = h
        self.devices = [int(d) for d in devices]
        self.members = [Context(d, _borrowed=self._lib.gpar_group_ctx(h, i)) for i, d in enumerate(self.devices)]

    def __len__(self):
        return len(self.members)

    def _check(self, st):
        if st != _ffi.GPAR_OK:
            msg = self._lib.gpar_group_last_error(self._h).decode("utf-8", "replace")
            raise (_ffi.PosDefException if st == _ffi.GPAR_ERR_NOT_POSDEF else _ffi.GparError)(st, msg)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            for m in self.members:
                m._h = None
            self._lib.gpar_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dtc_logpdf(self, kernel, thetas, vfe=False, jitter=-1.0, grad=False):
        """thetas: (ndev, 3), member i evaluates row i on its resident data -> vals (ndev,), [grads (ndev, 3)], codes."""
        th = as_f64(np.atleast_2d(thetas)); n = len(self)
        assert th.shape == (n, 3)
        vals = np.zeros(n); grads = np.zeros((n, 3)) if grad else None; codes = np.zeros(n, dtype=np.int32)
        self._check(self._lib.gpar_group_dtc_logpdf(self._h, int(kernel), dptr(th), int(bool(vfe)), float(jitter), dptr(vals), dptr(grads),
                                                    codes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return (vals, grads, codes) if grad else (vals, codes)

    def dtc_logpdf_sharded(self, kernel, theta, vfe=False, jitter=-1.0, grad=False):
        """ONE DTC / VFE objective over the data slices resident on the members (same pseudo-inputs everywhere):
        per-slice statistics, one NCCL all-reduce, tail on member 0 -> value[, gradient (3,)]."""
        th = as_f64(np.asarray(theta).ravel())
        val = ctypes.c_double(); g3 = np.zeros(3) if grad else None
        self._check(self._lib.gpar_group_dtc_logpdf_sharded(self._h, int(kernel), dptr(th), int(bool(vfe)), float(jitter), ctypes.byref(val), dptr(g3)))
        return (val.value, g3) if grad else val.value

    def load_row_slices(self, X, Z, t, y, bounds=None):
        """Row-shards one scaled problem over the members: every member gets the full (t, y) and Z, member i rows
        bounds[i]:bounds[i+1] of X (default: equal slices, boundaries rounded to multiples of 1024) -> row_lo (ndev,) int64."""
        X = np.asarray(X, dtype=np.float64); X = X[:, None] if X.ndim == 1 else X
        n = len(self); N = X.shape[0]
        if bounds is None:
            bounds = [min(N, (N * i // n + 1023) // 1024 * 1024) for i in range(n)] + [N]
        assert len(bounds) == n + 1 and bounds[0] == 0 and bounds[-1] == N
        for i, m in enumerate(self.members):
            m.set_times(t); m.set_outputs(y); m.set_pseudo(Z); m.set_inputs(np.ascontiguousarray(X[bounds[i]:bounds[i + 1]]))
            m.set_noise_vector(None)
        return np.asarray(bounds[:-1], dtype=np.int64)

    def scaled_dtc_sharded(self, k_time, k_out, theta, row_lo, grad=False):
        """ONE scaled-GPAR objective over the row slices resident on the members (load_row_slices): one all-gather of the
        slice summaries (filter carry), one all-reduce of (beta'beta, beta'alpha), tail on member 0 -> value[, gradient (5,)]
        (gradient: a second all-gather of the tangent summaries)."""
        th = as_f64(np.asarray(theta).ravel()); lo = np.ascontiguousarray(row_lo, dtype=np.int64)
        assert th.shape == (5,) and lo.shape == (len(self),)
        val = ctypes.c_double(); g5 = np.zeros(5) if grad else None
        self._check(self._lib.gpar_group_scaled_dtc_sharded(self._h, int(k_time), int(k_out), dptr(th),
                                                            lo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(val), dptr(g5)))
        return (val.value, g5) if grad else val.value

    def compute_q_u_sharded(self, k_time, k_out, params, row_lo):
        """compute_q_u on the row slices resident on the members -> (m_e (M,), inv(D) (M, M), U_u (M, M)) as Context.compute_q_u."""
        p = as_f64(np.asarray(params).ravel()); lo = np.ascontiguousarray(row_lo, dtype=np.int64)
        M = self.members[0].M
        m_e = np.zeros(M); Dinv = np.zeros((M, M), order="F"); U_u = np.zeros((M, M), order="F")
        self._check(self._lib.gpar_group_compute_q_u_sharded(self._h, int(k_time), int(k_out), dptr(p), lo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                                             dptr(m_e), dptr(Dinv), dptr(U_u)))
        return m_e, Dinv, U_u

    def sample_q_u_sharded(self, k_time, k_out, params, row_lo, seed, nsamples):
        """Seeded device draws from the q(u) of the row-sharded evaluation -> (W = U_u \\ eps, eps), each (M, S)."""
        p = as_f64(np.asarray(params).ravel()); lo = np.ascontiguousarray(row_lo, dtype=np.int64)
        M = self.members[0].M
        W = np.zeros((M, nsamples), order="F"); E = np.zeros((M, nsamples), order="F")
        self._check(self._lib.gpar_group_sample_q_u_sharded(self._h, int(k_time), int(k_out), dptr(p), lo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                                            int(seed), int(nsamples), dptr(W), dptr(E)))
        return W, E

    def fit_sharded(self, k_time, k_out, row_lo, theta0, iterations=200, optimizer="neldermead"):
        """One whole fit on the row-sharded objective (slices resident: load_row_slices), optimiser in the library (C++ twins of
        neldermead.py / lbfgs.py) -> (minimum of the negated objective, minimizer (5,), f_calls)."""
        lo = np.ascontiguousarray(row_lo, dtype=np.int64); th0 = as_f64(np.asarray(theta0).ravel())
        assert th0.shape == (5,) and lo.shape == (len(self),)
        fmin = ctypes.c_double(); xmin = np.zeros(5); calls = ctypes.c_int32()
        self._check(self._lib.gpar_group_fit_sharded(self._h, int(k_time), int(k_out), lo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), dptr(th0),
                                                     {"neldermead": 0, "lbfgs": 1}[optimizer], int(iterations), ctypes.byref(fmin), dptr(xmin), ctypes.byref(calls)))
        return fmin.value, xmin, calls.value

    def scaled_dtc(self, k_time, k_out, thetas, grad=False):
        th = as_f64(np.atleast_2d(thetas)); n = len(self)
        assert th.shape == (n, 5)
        vals = np.zeros(n); grads = np.zeros((n, 5)) if grad else None; codes = np.zeros(n, dtype=np.int32)
        self._check(self._lib.gpar_group_scaled_dtc(self._h, int(k_time), int(k_out), dptr(th), dptr(vals), dptr(grads),
                                                    codes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return (vals, grads, codes) if grad else (vals, codes)

    def fit(self, t, tasks, k_time, k_out, iterations, optimizer="neldermead"):
        """tasks: list of dicts {X (N, D) or None, Z (M, D) or None, y (N,), theta0} -> (minimum, minimizer (ntasks, 5),
        f_calls, member_of)."""
        t = as_f64(np.asarray(t).ravel())
        arr = (_ffi.FitTask * len(tasks))()
        keep = []
        for k, tk in enumerate(tasks):
            y = as_f64(np.asarray(tk["y"]).ravel()); keep.append(y)
            arr[k].y = dptr(y)
            if tk.get("X") is not None:
                X = as_f64(np.atleast_2d(tk["X"])); Z = as_f64(np.atleast_2d(tk["Z"])); keep += [X, Z]
                arr[k].X = dptr(X); arr[k].D = X.shape[1]; arr[k].Z = dptr(Z); arr[k].M = Z.shape[0]
            else:
                arr[k].X = None; arr[k].D = 0; arr[k].Z = None; arr[k].M = 0
            th0 = np.asarray(tk["theta0"], dtype=np.float64)
            for j in range(5):
                arr[k].theta0[j] = float(th0[j]) if j < th0.size else 0.0
        n = len(tasks)
        minimum = np.full(n, np.nan); minimizer = np.full((n, 5), np.nan)
        calls = np.zeros(n, dtype=np.int32); member = np.full(n, -1, dtype=np.int32)
        self._check(self._lib.gpar_group_fit(self._h, dptr(t), t.shape[0], ctypes.cast(arr, ctypes.c_void_p), n, int(k_time), int(k_out),
                                             {"neldermead": 0, "lbfgs": 1}[optimizer], int(iterations), dptr(minimum), dptr(minimizer),
                                             calls.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), member.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return minimum, minimizer, calls, member

    def broadcast(self, src, values=None, n=None):
        """values (host) or the resident result of member src's last smooth / predict -> every member's chain
        buffer; returns the host copy read back from a receiving member."""
        v = None if values is None else as_f64(np.asarray(values).ravel())
        n = int(v.shape[0] if v is not None else n)
        out = np.zeros(n)
        self._check(self._lib.gpar_group_broadcast(self._h, int(src), dptr(v), n, dptr(out)))
        return out
