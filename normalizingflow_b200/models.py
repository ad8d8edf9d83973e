"""NormalizingFlowModel with the reference's interface (nf/models.py:5-40).

    forward(x)  -> (z, prior_logprob, log_det)        models.py:13-20
    inverse(z)  -> (x, log_det)                       models.py:22-29
    sample(n)   -> (x, log_px, z)   detached          models.py:31-35
    evaluate(x) -> log_px           detached          models.py:37-40

Differences that do not change results: under ``torch.no_grad()`` each layer accumulates its
log-det into the running [N] buffer inside its kernel (no separate add pass); consecutive
``Planar`` layers run as one fused stack kernel; a ``GaussianPrior`` is evaluated by the
log-prob reduction kernel, and so is the prior the reference itself builds,
``MultivariateNormal(0, vars*I)`` (applications/src/setup.py:25-30): a zero-mean scaled-identity
``MultivariateNormal`` is recognised once and its ``log_prob`` routed to the same kernel (its ``sample`` stays
torch's, so seeds reproduce the reference's latents).  Any other prior object with ``.sample((n,))`` /
``.log_prob(x)`` is used as is.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn as nn

from . import _lib, _ops
from .flows import Planar, PlanarStack, Radial, RadialStack


class GaussianPrior:
    """N(0, var*I_d) — the prior the reference's hot configs build with
    MultivariateNormal(0, vars*I) (applications/src/setup.py:25-30) — on the log-prob kernel."""

    def __init__(self, dim, var=1.0, device="cuda"):
        self.dim = dim
        self.var = float(var)
        self.device = torch.device(device)
        self.generator = None

    def sample(self, sample_shape=torch.Size()):
        shape = tuple(sample_shape) + (self.dim,)
        z = torch.randn(shape, device=self.device, dtype=torch.float32, generator=self.generator)
        return z * math.sqrt(self.var) if self.var != 1.0 else z

    def log_prob(self, x):
        flat = x.reshape(-1, self.dim)
        if torch.is_grad_enabled() and flat.requires_grad:
            out = _ops.GaussLogprobFn.apply(flat, self.var)
        else:
            out = _ops.gauss_logprob(flat, self.var)
        return out.reshape(x.shape[:-1])


def _scaled_identity_var(prior, why=None):
    """var when ``prior`` is a zero-mean MultivariateNormal with covariance var*I (what the reference's
    hot configs construct, setup.py:25-30), else None.  One host read per prior object.  The Cholesky
    factor torch computed is accepted as sigma*I when its off-diagonal and the spread of its diagonal are
    below 1e-6 of sigma (a device factorisation may differ from the exact sqrt in the last bit)."""
    def no(msg):
        if why is not None:
            why.append(msg)
        return None
    if not isinstance(prior, torch.distributions.MultivariateNormal):
        return no(f"not a MultivariateNormal: {type(prior).__name__}")
    loc, tril = prior.loc, prior._unbroadcasted_scale_tril
    if loc.dim() != 1 or tril.dim() != 2:
        return no(f"batched prior: loc {tuple(loc.shape)}, scale_tril {tuple(tril.shape)}")
    diag = torch.diagonal(tril).double()
    sigma = float(diag.mean())
    off = float((tril.double() - torch.diag(diag)).abs().max())
    spread = float((diag - sigma).abs().max())
    mean = float(loc.abs().max())
    if not (sigma > 0.0) or mean != 0.0 or off > 1e-6 * sigma or spread > 1e-6 * sigma:
        return no(f"not a zero-mean scaled identity: |loc| {mean:.3g}, off-diagonal {off:.3g}, diagonal spread "
                  f"{spread:.3g}, sigma {sigma:.6g}")
    # prefer the caller's own covariance entry when it is available (exact var, not sigma^2 re-rounded)
    cov = prior.__dict__.get("covariance_matrix")
    if cov is not None and cov.dim() == 2:
        return float(cov[0, 0])
    return sigma * sigma


class NormalizingFlowModel(nn.Module):

    def __init__(self, prior, flows, device="cpu", fuse_planar=True):
        super().__init__()
        self.device = device
        self.prior = prior
        self.flows = nn.ModuleList(flows)
        self.fuse_planar = fuse_planar
        self._prior_var_cache = None

    def _prior_var(self):
        """Variance of an isotropic zero-mean Gaussian prior (GaussianPrior or a scaled-identity
        MultivariateNormal), else None."""
        prior = self.prior
        if isinstance(prior, GaussianPrior):
            return prior.var
        c = getattr(self, "_prior_var_cache", None)
        if c is None or c[0] is not prior:
            c = self._prior_var_cache = (prior, _scaled_identity_var(prior))
        return c[1]

    def prior_log_prob(self, z, add=None, add_sign=1.0):
        """prior.log_prob(z) (+ add_sign * add) — nf/models.py:19-20, :34, :39.  Isotropic Gaussian priors
        run on the log-prob reduction kernel with the addition folded in (no-grad calls)."""
        var = self._prior_var()
        if var is not None and z.is_cuda and z.dim() == 2 and not (torch.is_grad_enabled() and z.requires_grad):
            return _ops.gauss_logprob(z, var, add=add, add_sign=add_sign)
        if var is not None and z.is_cuda and z.dim() == 2 and add is None:
            return _ops.GaussLogprobFn.apply(z, var)
        lp = self.prior.log_prob(z)
        return lp if add is None else lp + add_sign * add

    # runs of consecutive Planar layers, and of consecutive Radial layers of one mode, collapse into fused launches
    def _forward_plan(self):
        plan, run = [], []

        def flush():
            if run:
                if len(run) == 1:
                    plan.append(run[0])
                else:
                    plan.append(PlanarStack(list(run)) if type(run[0]) is Planar else RadialStack(list(run)))
                run.clear()
        for f in self.flows:
            fusable = self.fuse_planar and type(f) in (Planar, Radial)
            if fusable and run and (type(run[0]) is not type(f)
                                    or (type(f) is Radial and bool(f.per_sample) != bool(run[0].per_sample))):
                flush()
            if fusable:
                run.append(f)
                continue
            flush()
            plan.append(f)
        flush()
        return plan

    def forward(self, x):
        m, _ = x.shape
        log_det = torch.zeros(m, dtype=torch.float32, device=x.device)     # fp32 always (quirk Q11)
        fused = not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))
        for flow in self._forward_plan():
            if fused and hasattr(flow, "_nfk_step"):
                x, log_det = flow._nfk_step(x, log_det, False)
            else:
                x, ld = flow.forward(x)
                log_det = log_det + ld
        z, prior_logprob = x, self.prior_log_prob(x)
        return z, prior_logprob, log_det

    def inverse(self, z):
        m, _ = z.shape
        log_det = torch.zeros(m, dtype=torch.float32, device=z.device)
        fused = not (torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in self.parameters())))
        for flow in self.flows[::-1]:
            if fused and hasattr(flow, "_nfk_step"):
                z, log_det = flow._nfk_step(z, log_det, True)
            else:
                z, ld = flow.inverse(z)
                log_det = log_det + ld
        x = z
        return x, log_det

    def sample(self, n_samples):
        with torch.no_grad():
            z = self.prior.sample((n_samples,))
            x, log_det = self.inverse(z)
            log_px = self.prior_log_prob(z, add=log_det, add_sign=-1.0)          # models.py:34
        return x.data, log_px.data, z.data

    def evaluate(self, x):
        with torch.no_grad():
            log_px = self._log_px(x)
        return log_px.data

    def _log_px(self, x):
        """log p(x) = prior.log_prob(f(x)) + log_det (models.py:37-39) with the sum folded into the
        log-prob reduction kernel."""
        m, _ = x.shape
        log_det = torch.zeros(m, dtype=torch.float32, device=x.device)
        for flow in self._forward_plan():
            if hasattr(flow, "_nfk_step"):
                x, log_det = flow._nfk_step(x, log_det, False)
            else:
                x, ld = flow.forward(x)
                log_det = log_det + ld
        return self.prior_log_prob(x, add=log_det, add_sign=1.0)


    # ------------------------------------------------------------------------------------
    # host-buffer entry points: batches that live in (pinned) host memory are streamed through
    # the GPU in row chunks, with the host->device copy of chunk i+1 and the device->host copy of
    # chunk i-1 overlapping the kernels of chunk i (three CUDA streams, rotating device buffers).
    # Same results as evaluate()/inverse() on the whole batch: rows are independent.
    # ------------------------------------------------------------------------------------
    def _host_pipe(self, dev, rows, d):
        """Copy streams and rotating device buffers, kept across calls: the next call's first
        host->device copy then overlaps the previous call's last kernels and device->host copy."""
        key = (str(dev), rows, d)
        pipes = self.__dict__.setdefault("_pipes", {})
        pipe = pipes.get(key)
        if pipe is None:
            if len(pipes) >= 4:                  # bound the device memory held by rotating buffers
                self.host_sync()
                pipes.clear()
            nbuf = 3
            pipe = pipes[key] = dict(key=key, s_in=torch.cuda.Stream(dev), s_out=torch.cuda.Stream(dev), nbuf=nbuf,
                                     it=0, free=[None] * nbuf,
                                     bufs=[torch.empty((rows, d), dtype=torch.float32, device=dev)
                                           for _ in range(nbuf)])
        return pipe

    def _param_version(self):
        """Changes whenever a parameter is replaced or written in place (optimizer step, load_state_dict):
        captured chunk graphs hold pointers to weight images packed from one parameter version."""
        return (_lib.param_epoch(),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _stream_rows(self, host_in, fn, outs, chunk_rows, wait=True, tag=None):
        """``tag`` names the computation (``fn``) for the chunk-graph cache: with a tag, the kernels of a
        full-size chunk are captured once per rotating input buffer as ONE CUDA graph and replayed, so
        a chunk costs the host a copy, a graph launch and the result copies instead of ~40 dispatches
        (the pipeline is otherwise host-dispatch-bound as soon as the CPU is busy: measured end-to-end
        rate 36-77 M samples/s eager on the same box).  Ragged last chunks run eagerly."""
        dev = next(self.parameters()).device
        n, d = host_in.shape
        cur = torch.cuda.current_stream(dev)
        pipe = self._host_pipe(dev, min(chunk_rows, n), d)
        s_in, s_out, nbuf, bufs, free = pipe["s_in"], pipe["s_out"], pipe["nbuf"], pipe["bufs"], pipe["free"]
        use_graphs = tag is not None and self.host_graphs and not torch.cuda.is_current_stream_capturing()
        if use_graphs:
            ver = self._param_version()
            if pipe.get("ver") != ver:
                pipe["ver"], pipe["graphs"] = ver, {}
        with torch.no_grad():
            for a in range(0, n, chunk_rows):
                b = min(n, a + chunk_rows)
                k = pipe["it"] % nbuf
                pipe["it"] += 1
                with torch.cuda.stream(s_in):
                    if free[k] is not None:
                        s_in.wait_event(free[k])            # the kernels that read this buffer are done
                    xin = bufs[k][: b - a]
                    xin.copy_(host_in[a:b], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(s_in)
                cur.wait_event(ready)
                slot = None
                if use_graphs and b - a == bufs[k].shape[0]:
                    slot = pipe["graphs"].get((tag, k))
                    if slot is None:
                        slot = self._capture_chunk(fn, bufs[k], cur)
                        pipe["graphs"][(tag, k)] = slot
                    if slot:
                        if slot["drained"] is not None:
                            cur.wait_event(slot["drained"])   # its static outputs have been copied out
                        slot["graph"].replay()
                        res = slot["out"]
                if not slot:
                    res = fn(xin)
                done = torch.cuda.Event()
                done.record(cur)
                free[k] = done
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    for r, o in zip(res, outs):
                        o[a:b].copy_(r, non_blocking=True)
                        if not slot:
                            r.record_stream(s_out)
                    if slot:
                        slot["drained"] = torch.cuda.Event()
                        slot["drained"].record(s_out)
        if wait:
            # the returned host tensors are read by the CPU, not by a stream: block the host until the
            # device->host copies have landed (and keep the current stream ordered after them)
            cur.wait_stream(s_out)
            s_out.synchronize()
        return outs

    def _stream_generated(self, n, fn, outs, chunk_rows, wait=True, tag=None):
        """Like _stream_rows for work that has no host input: ``fn(rows)`` produces ``rows`` result rows on the
        device (e.g. draws latents and pushes them through the flow); results go to the host tensors
        ``outs`` chunk by chunk, the device->host copy of chunk i overlapping the kernels of chunk i+1."""
        dev = next(self.parameters()).device
        cur = torch.cuda.current_stream(dev)
        rows_full = min(chunk_rows, n)
        pipe = self._host_pipe(dev, rows_full, 0)
        s_out, nbuf = pipe["s_out"], pipe["nbuf"]
        # a user-supplied generator is not registered with the capture: run eagerly then
        use_graphs = (tag is not None and self.host_graphs and not torch.cuda.is_current_stream_capturing()
                      and getattr(self.prior, "generator", None) is None)
        if use_graphs:
            ver = self._param_version()
            if pipe.get("ver") != ver:
                pipe["ver"], pipe["graphs"] = ver, {}
        pending = pipe.setdefault("pending", [None] * nbuf)
        with torch.no_grad():
            for a in range(0, n, chunk_rows):
                b = min(n, a + chunk_rows)
                k = pipe["it"] % nbuf
                pipe["it"] += 1
                slot = None
                if use_graphs and b - a == rows_full:
                    slot = pipe["graphs"].get((tag, k))
                    if slot is None:
                        slot = self._capture_chunk(fn, rows_full, cur)
                        pipe["graphs"][(tag, k)] = slot
                    if slot:
                        if slot["drained"] is not None:
                            cur.wait_event(slot["drained"])
                        slot["graph"].replay()
                        res = slot["out"]
                if not slot:
                    res = fn(b - a)
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    for r, o in zip(res, outs):
                        if o is None:
                            continue
                        o[a:b].copy_(r, non_blocking=True)
                        if not slot:
                            r.record_stream(s_out)
                    if slot:
                        slot["drained"] = torch.cuda.Event()
                        slot["drained"].record(s_out)
        if wait:
            cur.wait_stream(s_out)
            s_out.synchronize()
        return outs

    host_graphs = True          # set False to run every chunk eagerly

    def _capture_chunk(self, fn, xin, cur):
        """fn(xin) captured as a CUDA graph on a side stream (one eager call first: weight images are
        packed lazily).  Returns False when capture is not possible, and the chunk then runs eagerly."""
        dev = xin.device if isinstance(xin, torch.Tensor) else next(self.parameters()).device
        try:
            fn(xin)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                out = fn(xin)
            cur.wait_stream(side)
            return dict(graph=graph, out=tuple(out), drained=None)
        except Exception as e:       # e.g. a layer that synchronises; keep the eager path
            torch.cuda.synchronize(dev)
            if not getattr(self, "_warned_eager", False):
                self._warned_eager = True
                warnings.warn("normalizingflow_b200: CUDA-graph capture of a host-pipeline chunk failed "
                              f"({type(e).__name__}: {e}); chunks run eagerly (host-dispatch-bound, slower)",
                              RuntimeWarning, stacklevel=2)
            return False

    def host_sync(self):
        """Block the host until every evaluate_host / inverse_host result has landed (needed before
        reading results of calls made with ``wait=False``)."""
        for pipe in self.__dict__.get("_pipes", {}).values():
            pipe["s_out"].synchronize()

    def evaluate_host(self, x_host, out=None, chunk_rows=131072, wait=True):
        """log p(x) for a batch in host memory -> host tensor [N] (pinned when allocated here).
        ``wait=True`` (default): the result is complete in host memory when the call returns.
        ``wait=False`` returns while copies are in flight, so back-to-back calls overlap completely;
        call ``host_sync()`` before reading any result."""
        if out is None:
            out = torch.empty(x_host.shape[0], dtype=torch.float32).pin_memory()

        def fn(x):
            return (self._log_px(x),)
        self._stream_rows(x_host, fn, (out,), chunk_rows, wait, tag="evaluate")
        return out

    def inverse_host(self, z_host, out_x=None, out_log_px=None, chunk_rows=131072, wait=True):
        """sampling direction for latents in host memory -> (x, log_px) host tensors, as sample()."""
        n, d = z_host.shape
        if out_x is None:
            out_x = torch.empty((n, d), dtype=torch.float32).pin_memory()
        if out_log_px is None:
            out_log_px = torch.empty(n, dtype=torch.float32).pin_memory()

        def fn(z):
            x, log_det = self.inverse(z)
            return x, self.prior_log_prob(z, add=log_det, add_sign=-1.0)
        self._stream_rows(z_host, fn, (out_x, out_log_px), chunk_rows, wait, tag="inverse")
        return out_x, out_log_px

    def sample_host(self, n_samples, out_x=None, out_log_px=None, out_z=None, return_z=True, chunk_rows=131072,
                    wait=True):
        """``sample(n)`` (nf/models.py:31-35) delivered to host memory: the latents are drawn ON THE DEVICE by
        ``prior.sample`` as the reference does, pushed through the inverse flow in row chunks, and
        (x, log_px[, z]) land in (pinned) host tensors while the next chunk computes.  No host->device
        traffic at all; ``return_z=False`` skips the copy of z."""
        d = getattr(self.prior, "dim", None)
        if d is None:
            ev = getattr(self.prior, "event_shape", None)
            d = int(ev[-1]) if ev else int(self.prior.sample((1,)).shape[-1])
        if out_x is None:
            out_x = torch.empty((n_samples, d), dtype=torch.float32).pin_memory()
        if out_log_px is None:
            out_log_px = torch.empty(n_samples, dtype=torch.float32).pin_memory()
        if return_z and out_z is None:
            out_z = torch.empty((n_samples, d), dtype=torch.float32).pin_memory()

        def fn(rows):
            z = self.prior.sample((rows,))
            x, log_det = self.inverse(z)
            return x, self.prior_log_prob(z, add=log_det, add_sign=-1.0), z
        self._stream_generated(n_samples, fn, (out_x, out_log_px, out_z if return_z else None), chunk_rows, wait,
                               tag="sample_z" if return_z else "sample")
        return (out_x, out_log_px, out_z) if return_z else (out_x, out_log_px)
