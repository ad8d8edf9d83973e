#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  Usage, from any directory:

    python /root/repo/oracle/gen_golden.py [--out /root/repo/tests/golden]

It imports ``nf.flows``, ``nf.flows_1``, ``nf.utils``, ``nf.models`` and ``nf.hmc`` from
/root/reference (with empty stub modules for the unused ``MDAnalysis`` / ``lammps``
imports, nf/utils.py:4 and nf/hmc.py:1), feeds them seeded inputs and records inputs,
weights and outputs.  Spline bin indices are captured by wrapping ``nf.utils.searchsorted``
(a module-global looked up at nf/utils.py:94-96).

The repository's own ``nf`` shim must NOT be importable while this runs, so the script
scrubs the repo root from ``sys.path`` first.
"""
import argparse
import math
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NF_REFERENCE", "/root/reference")
sys.path = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != REPO
            and os.path.abspath(p or os.getcwd()) != os.path.join(REPO, "oracle")]
sys.path.insert(0, REF)
sys.modules["MDAnalysis"] = types.ModuleType("MDAnalysis")
_l = types.ModuleType("lammps")
_l.lammps = _l.PyLammps = object
_l.LMP_STYLE_ATOM = _l.LMP_TYPE_ARRAY = 0
sys.modules["lammps"] = _l

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.distributions import MultivariateNormal  # noqa: E402

import nf.utils as ref_utils  # noqa: E402
import nf.flows as ref_flows  # noqa: E402
import nf.flows_1 as ref_flows1  # noqa: E402
import nf.models as ref_models  # noqa: E402
import nf.hmc as ref_hmc  # noqa: E402

assert ref_utils.__file__.startswith(REF), ref_utils.__file__

_captured = []
_orig_ss = ref_utils.searchsorted


def _ss(bin_locations, inputs, eps=1e-6):
    r = _orig_ss(bin_locations, inputs, eps)
    _captured.append(r.clone())
    return r


ref_utils.searchsorted = _ss


def npy(t):
    return t.detach().cpu().numpy()


def sd_np(module, prefix=""):
    return {prefix + k: npy(v) for k, v in module.state_dict().items()}


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def bins_full(x_upper, B, captured):
    """Scatter the captured inside-only bin indices back to [N, F_t], -1 in the tails."""
    inside = (x_upper >= -B) & (x_upper <= B)
    out = torch.full(x_upper.shape, -1, dtype=torch.int64)
    out[inside] = captured
    return out


def run_nsfcl(layer, x, inverse):
    """Returns (out, ld, bins[N,F_t], params[N,F_t,3K-1])."""
    _captured.clear()
    grabbed = {}
    hk = layer.psi.register_forward_hook(lambda m, i, o: grabbed.__setitem__("p", o.detach().clone()))
    with torch.no_grad():
        out, ld = layer.inverse(x) if inverse else layer.forward(x)
    hk.remove()
    x3 = x.reshape(-1, layer.size, layer.dim)
    upper = x3[:, :, layer.unmasked].flatten(start_dim=1)
    bins = bins_full(upper, layer.B, _captured[0])
    n_t = layer.size * (layer.dim - len(layer.mask))
    return out, ld, bins, grabbed["p"].reshape(-1, n_t, 3 * layer.K - 1)


def edge_inputs(x, B):
    """Plant exact boundary / just-outside / NaN-free edge values in the first rows."""
    x = x.clone()
    flat = x.view(-1)
    vals = [B, -B, math.nextafter(B, 10.0), -math.nextafter(B, 10.0), 0.0, B * 0.999999, 7.5, -11.0]
    for i, v in enumerate(vals):
        flat[i * 3] = v
    return x


def fix_nsfcl(out_dir, name, size, dim, K, B, H, masks, N, seed, scale_last=1.0, in_std=1.5):
    rec = dict(size=size, dim=dim, K=K, B=float(B), H=H, N=N, masks=np.array([",".join(map(str, m)) for m in masks]))
    for mi, mask in enumerate(masks):
        torch.manual_seed(seed + mi)
        layer = ref_flows.NSF_CL(size, dim=dim, K=K, B=B, hidden_dim=H, mask=mask)
        if scale_last != 1.0:
            with torch.no_grad():
                layer.psi.network[4].weight.mul_(scale_last)
                layer.psi.network[4].bias.mul_(scale_last)
        x = edge_inputs(torch.randn(N, size * dim, generator=gen(100 + seed + mi)) * in_std, float(B))
        zin = edge_inputs(torch.randn(N, size * dim, generator=gen(200 + seed + mi)) * in_std, float(B))
        z, ld, bins, params = run_nsfcl(layer, x, inverse=False)
        xi, ldi, binsi, paramsi = run_nsfcl(layer, zin, inverse=True)
        p = f"m{mi}."
        rec.update(sd_np(layer, p + "sd."))
        rec.update({p + "x": npy(x), p + "params": npy(params), p + "z": npy(z), p + "ld": npy(ld),
                    p + "bins": npy(bins).astype(np.int8),
                    p + "zin": npy(zin), p + "params_inv": npy(paramsi), p + "x_inv": npy(xi),
                    p + "ld_inv": npy(ldi), p + "bins_inv": npy(binsi).astype(np.int8)})
    np.savez_compressed(os.path.join(out_dir, name), **rec)


def fix_rqs_function(out_dir):
    """Direct calls of unconstrained_RQS (nf/utils.py:27) with raw random tensors."""
    rec = {}
    for tag, K, B, shape in (("k8", 8, 3.0, (40, 7)), ("k5", 5, 1.0, (33,)), ("k32", 32, 3.0, (16, 3))):
        g = gen(7 + K)
        v = torch.randn(*shape, generator=g) * (0.8 * B)
        W = torch.randn(*shape, K, generator=g) * 2
        Hh = torch.randn(*shape, K, generator=g) * 2
        D = torch.randn(*shape, K - 1, generator=g) * 2
        for inv in (False, True):
            _captured.clear()
            out, lad = ref_utils.unconstrained_RQS(v, W, Hh, D, inverse=inv, tail_bound=B)
            bins = bins_full(v, B, _captured[0])
            s = "inv" if inv else "fwd"
            rec.update({f"{tag}.{s}.out": npy(out), f"{tag}.{s}.lad": npy(lad),
                        f"{tag}.{s}.bins": npy(bins).astype(np.int8)})
        rec.update({f"{tag}.v": npy(v), f"{tag}.W": npy(W), f"{tag}.H": npy(Hh), f"{tag}.D": npy(D),
                    f"{tag}.K": K, f"{tag}.B": B})
    np.savez_compressed(os.path.join(out_dir, "rqs_function.npz"), **rec)


def fix_realnvp(out_dir):
    rec = {}
    for tag, d, H, N in (("d2", 2, 16, 64), ("d64", 64, 16, 32), ("d6", 6, 40, 50)):
        torch.manual_seed(11 + d)
        layer = ref_flows.RealNVP(d, hidden_dim=H)
        x = torch.randn(N, d, generator=gen(300 + d))
        zin = torch.randn(N, d, generator=gen(400 + d))
        with torch.no_grad():
            z, ld = layer.forward(x)
            xi, ldi = layer.inverse(zin)
        rec.update(sd_np(layer, tag + ".sd."))
        rec.update({tag + ".x": npy(x), tag + ".z": npy(z), tag + ".ld": npy(ld),
                    tag + ".zin": npy(zin), tag + ".x_inv": npy(xi), tag + ".ld_inv": npy(ldi),
                    tag + ".d": d, tag + ".H": H})
    np.savez_compressed(os.path.join(out_dir, "realnvp.npz"), **rec)


def fix_nsf_ar(out_dir):
    """NSF_AR (nf/flows.py:152-209), the flow type of the shipped experiment YAMLs: forward, inverse,
    bins per dimension, and autograd gradients of sum(z*r) + sum(ld*s) w.r.t. x and every parameter."""
    rec = {}
    for tag, dim, K, B, H, N in (("d6k8", 6, 8, 3.0, 16, 96), ("d4k32", 4, 32, 3, 12, 64), ("d12k8", 12, 8, 4.0, 24, 40)):
        torch.manual_seed(77 + dim)
        layer = ref_flows.NSF_AR(dim, K=K, B=B, hidden_dim=H)
        with torch.no_grad():
            for l in layer.layers:                       # non-uniform bins
                l.network[4].weight.mul_(3.0)
        x = edge_inputs(1.5 * torch.randn(N, dim, generator=gen(500 + dim)), float(B))
        zin = edge_inputs(1.5 * torch.randn(N, dim, generator=gen(600 + dim)), float(B))
        _captured.clear()
        with torch.no_grad():
            z, ld = layer.forward(x)
        fb = [c.clone() for c in _captured]
        _captured.clear()
        with torch.no_grad():
            xi, ldi = layer.inverse(zin)
        ib = [c.clone() for c in _captured]
        _captured.clear()
        rec.update(sd_np(layer, tag + ".sd."))
        rec.update({tag + ".x": npy(x), tag + ".z": npy(z), tag + ".ld": npy(ld), tag + ".zin": npy(zin),
                    tag + ".x_inv": npy(xi), tag + ".ld_inv": npy(ldi), tag + ".dim": dim, tag + ".K": K,
                    tag + ".B": float(B), tag + ".H": H})
        # bins of the inside elements, one capture per dimension (searchsorted sees only inside inputs)
        for name, caps, inp in (("bins", fb, x), ("bins_inv", ib, zin)):
            full = torch.full((N, dim), -1, dtype=torch.int64)
            for i in range(dim):
                inside = (inp[:, i] >= -B) & (inp[:, i] <= B)
                full[inside, i] = caps[i].reshape(-1)
            rec[tag + "." + name] = npy(full)
        # gradients
        r = torch.randn(N, dim, generator=gen(700 + dim))
        s = torch.randn(N, generator=gen(800 + dim))
        xg = x.clone().requires_grad_()
        zz, ll = layer.forward(xg)
        loss = (zz * r).sum() + (ll * s).sum()
        grads = torch.autograd.grad(loss, [xg] + list(layer.parameters()))
        rec[tag + ".r"] = npy(r)
        rec[tag + ".s"] = npy(s)
        rec[tag + ".gx"] = npy(grads[0])
        for (n, _), gv in zip(layer.named_parameters(), grads[1:]):
            rec[tag + ".g." + n] = npy(gv)
    np.savez_compressed(os.path.join(out_dir, "nsf_ar.npz"), **rec)


def fix_checkpoint(out_dir):
    """A checkpoint exactly as applications/src/train.py:39-40 writes it
    ({"model", "optim", "scheduler", "epoch", "loss"}), plus the model's outputs on a seeded batch."""
    torch.manual_seed(5)
    flows = [ref_flows.NSF_CL(4, dim=2, K=8, B=3.0, hidden_dim=8, mask=[i % 2]) for i in range(2)] + \
            [ref_flows.RealNVP(8, hidden_dim=8)]
    prior = MultivariateNormal(torch.zeros(8), torch.eye(8))
    model = ref_models.NormalizingFlowModel(prior, flows)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=10, gamma=0.5)
    x = torch.randn(32, 8, generator=gen(900))
    losses = []
    for _ in range(3):                                   # a few steps of train.py:22-29
        opt.zero_grad()
        z, plp, ld = model(x)
        loss = -torch.mean(plp + ld)
        losses.append(loss.mean().data)
        loss.backward()
        opt.step()
        sched.step()
    torch.save({"model": model.state_dict(), "optim": opt.state_dict(), "scheduler": sched.state_dict(), "epoch": 3,
                "loss": losses}, os.path.join(out_dir, "ref_checkpoint.pth"))
    with torch.no_grad():
        z, plp, ld = model(x)
    np.savez_compressed(os.path.join(out_dir, "ref_checkpoint_out.npz"), x=npy(x), z=npy(z), plp=npy(plp), ld=npy(ld))


def fix_systems(out_dir):
    """Priors / targets of applications/src/systems.py (EinsteinCrystal, LJ.potential, GaussianMixture) run
    unmodified (stub modules for lammps / MDAnalysis / matplotlib, which these three never touch)."""
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, os.path.join(REF, "applications"))
    from src import systems as ref_sys
    rec = {}
    # Einstein crystal: lattice + displaced samples, with and without the periodic wrap
    g0 = gen(1000)
    centers = (torch.rand(12, 3, generator=g0) - 0.5) * 3.6
    for tag, alpha, box in (("ec_free", 50, None), ("ec_box", 1000, 4.0)):
        ec = ref_sys.EinsteinCrystal(centers.tolist(), dim=3, boxlength=box, alpha=alpha)
        x = (centers + torch.randn(40, 12, 3, generator=g0) * (3.0 / alpha) ** 0.5).reshape(40, -1)
        if box is not None:
            x[:5] += box                                   # images one box length away: wrapped back
        rec.update({tag + ".centers": npy(centers), tag + ".alpha": float(alpha), tag + ".box": float(box or 0.0),
                    tag + ".x": npy(x), tag + ".lp": npy(ec.log_prob(x.clone()))})
    # Lennard-Jones: jittered cubic lattice (no overlaps), all four cutoff/shift combinations
    n_side, a0 = 3, 1.15
    grid = torch.stack(torch.meshgrid(*[torch.arange(n_side) * a0] * 3, indexing="ij"), -1).reshape(-1, 3).float()
    box = n_side * a0
    pos = grid[None] + 0.08 * torch.randn(30, n_side ** 3, 3, generator=g0) - box / 2
    rec.update({"lj.pos": npy(pos), "lj.box": float(box)})
    for tag, cutoff, shift in (("lj_nocut", None, True), ("lj_cut_shift", 1.6, True), ("lj_cut", 1.6, False)):
        lj = ref_sys.LJ(boxlength=box, epsilon=1.3, sigma=0.95, cutoff=cutoff, shift=shift)
        rec.update({tag + ".cutoff": float(cutoff or 0.0), tag + ".shift": int(shift),
                    tag + ".U": npy(lj.potential(pos.clone()))})
    # Gaussian mixture
    gm_c = [[-1.0, 0.5], [1.2, -0.3], [0.1, 1.4], [0.0, 0.0]]
    gm_v = [0.5, 0.3, 0.8, 0.2]
    gm = ref_sys.GaussianMixture(gm_c, gm_v, npoints=3, dim=2)
    xg = torch.randn(50, 6, generator=g0) * 1.2
    rec.update({"gm.centers": np.array(gm_c, dtype=np.float32), "gm.vars": np.array(gm_v, dtype=np.float32),
                "gm.x": npy(xg), "gm.lp": npy(gm.log_prob(xg))})
    np.savez_compressed(os.path.join(out_dir, "systems.npz"), **rec)


def fix_bar(out_dir):
    """BAR free-energy estimate of the UNMODIFIED applications/src/bar.py (pure numpy) on seeded work values:
    forward works ~ N(2.0 + 0.5 s^2, s), reverse works ~ N(-2.0 + 0.5 s^2, s) (exact DeltaF = 2.0)."""
    sys.path.insert(0, os.path.join(REF, "applications"))
    from src import bar as ref_bar
    rec = {}
    rng = np.random.default_rng(7)
    for tag, s, nF, nR in (("a", 1.0, 4000, 4000), ("b", 2.5, 6000, 1500), ("c", 0.3, 300, 900)):
        wF = rng.normal(2.0 + 0.5 * s * s, s, nF)
        wR = rng.normal(-2.0 + 0.5 * s * s, s, nR)
        rec[tag + ".wF"], rec[tag + ".wR"] = wF, wR
        rec[tag + ".dF64"] = ref_bar.BAR(wF, wR)
        rec[tag + ".dF32"] = float(ref_bar.BAR(wF.astype(np.float32), wR.astype(np.float32)))
        rec[tag + ".fzero_at_1"] = ref_bar.BARzero(wF, wR, 1.0)
    np.savez_compressed(os.path.join(out_dir, "bar.npz"), **rec)


def init_radial(layer, d, seed):
    with torch.no_grad():
        b = math.sqrt(1 / d)
        g = gen(seed)
        layer.x0.copy_((torch.rand(d, generator=g) * 2 - 1) * b)
        layer.log_alpha.copy_((torch.rand(1, generator=g) * 2 - 1) * b)
        layer.beta.copy_((torch.rand(1, generator=g) * 2 - 1) * b)


def fix_planar_radial(out_dir):
    rec = {}
    for d, N in ((128, 48), (5, 33)):
        tag = f"d{d}"
        torch.manual_seed(21 + d)
        pl = ref_flows1.Planar(d)
        rd = ref_flows1.Radial(d)
        init_radial(rd, d, 31 + d)
        x = torch.randn(N, d, generator=gen(500 + d))
        with torch.no_grad():
            zp, ldp = pl.forward(x)
            zr, ldr = rd.forward(x)
        rec.update(sd_np(pl, tag + ".planar."))
        rec.update(sd_np(rd, tag + ".radial."))
        rec.update({tag + ".x": npy(x), tag + ".planar.z": npy(zp), tag + ".planar.ld": npy(ldp),
                    tag + ".radial.z": npy(zr), tag + ".radial.ld": npy(ldr)})
    np.savez_compressed(os.path.join(out_dir, "planar_radial.npz"), **rec)


def fix_models(out_dir):
    # (1) cfg-2 shaped model, small hidden size
    rec = {}
    d, L, H, N = 64, 8, 16, 64
    torch.manual_seed(0)
    flows = [ref_flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(L)]
    prior = MultivariateNormal(torch.zeros(d), torch.eye(d))
    model = ref_models.NormalizingFlowModel(prior, flows)
    x = torch.randn(N, d, generator=gen(1))
    zin = torch.randn(N, d, generator=gen(2))
    with torch.no_grad():
        z, plp, ld = model.forward(x)
        xi, ldi = model.inverse(zin)
        ev = model.evaluate(x)
        lp_z = prior.log_prob(zin)
    rec.update(sd_np(model, "nsf.sd."))
    rec.update({"nsf.x": npy(x), "nsf.z": npy(z), "nsf.prior_lp": npy(plp), "nsf.ld": npy(ld),
                "nsf.zin": npy(zin), "nsf.x_inv": npy(xi), "nsf.ld_inv": npy(ldi),
                "nsf.evaluate": npy(ev), "nsf.sample_logpx": npy(lp_z - ldi)})
    # (2) cfg-1 shaped model: 8 x RealNVP(2)
    torch.manual_seed(3)
    flows = [ref_flows.RealNVP(2, hidden_dim=12) for _ in range(8)]
    prior = MultivariateNormal(torch.zeros(2), torch.eye(2))
    model = ref_models.NormalizingFlowModel(prior, flows)
    x = torch.randn(128, 2, generator=gen(4))
    with torch.no_grad():
        z, plp, ld = model.forward(x)
        xi, ldi = model.inverse(x)
    rec.update(sd_np(model, "rnvp.sd."))
    rec.update({"rnvp.x": npy(x), "rnvp.z": npy(z), "rnvp.prior_lp": npy(plp), "rnvp.ld": npy(ld),
                "rnvp.x_inv": npy(xi), "rnvp.ld_inv": npy(ldi)})
    # (3) cfg-4 shaped model: planar stack
    torch.manual_seed(5)
    flows = [ref_flows1.Planar(16) for _ in range(6)]
    prior = MultivariateNormal(torch.zeros(16), torch.eye(16))
    model = ref_models.NormalizingFlowModel(prior, flows)
    x = torch.randn(40, 16, generator=gen(6))
    with torch.no_grad():
        z, plp, ld = model.forward(x)
    rec.update(sd_np(model, "planar.sd."))
    rec.update({"planar.x": npy(x), "planar.z": npy(z), "planar.prior_lp": npy(plp), "planar.ld": npy(ld)})
    np.savez_compressed(os.path.join(out_dir, "models.npz"), **rec)


def fix_grads(out_dir):
    """Autograd gradients of the reference layers for custom-backward parity."""
    rec = {}
    # NSF_CL, both directions, two masks
    for mi, (size, dim, mask) in enumerate(((32, 2, [1]), (6, 3, [0, 2]))):
        torch.manual_seed(40 + mi)
        layer = ref_flows.NSF_CL(size, dim=dim, K=8, B=3.0, hidden_dim=12, mask=mask)
        with torch.no_grad():
            layer.psi.network[4].weight.mul_(3.0)
        N = 24
        d = size * dim
        for inv in (False, True):
            s = f"nsf{mi}." + ("inv." if inv else "fwd.")
            x = (torch.randn(N, d, generator=gen(600 + mi + inv)) * 1.5).requires_grad_()
            gz = torch.randn(N, d, generator=gen(610 + mi))
            gl = torch.randn(N, generator=gen(620 + mi))
            grabbed = {}

            def hook(m, i, o):
                o.retain_grad()
                grabbed["p"] = o
            hk = layer.psi.register_forward_hook(hook)
            layer.zero_grad()
            out, ld = layer.inverse(x) if inv else layer.forward(x)
            ((out * gz).sum() + (ld * gl).sum()).backward()
            hk.remove()
            n_t = size * (dim - len(mask))
            rec.update({s + "x": npy(x), s + "gz": npy(gz), s + "gl": npy(gl), s + "out": npy(out),
                        s + "ld": npy(ld), s + "gx": npy(x.grad),
                        s + "params": npy(grabbed["p"]).reshape(N, n_t, 23),
                        s + "gparams": npy(grabbed["p"].grad).reshape(N, n_t, 23)})
            for k, v in layer.named_parameters():
                rec[s + "gw." + k] = npy(v.grad)
        rec.update(sd_np(layer, f"nsf{mi}.sd."))
        rec.update({f"nsf{mi}.size": size, f"nsf{mi}.dim": dim, f"nsf{mi}.mask": np.array(mask)})
    # RealNVP
    torch.manual_seed(50)
    layer = ref_flows.RealNVP(6, hidden_dim=10)
    for inv in (False, True):
        s = "rnvp." + ("inv." if inv else "fwd.")
        x = torch.randn(20, 6, generator=gen(700 + inv)).requires_grad_()
        gz = torch.randn(20, 6, generator=gen(710))
        gl = torch.randn(20, generator=gen(720))
        layer.zero_grad()
        out, ld = layer.inverse(x) if inv else layer.forward(x)
        ((out * gz).sum() + (ld * gl).sum()).backward()
        rec.update({s + "x": npy(x), s + "gz": npy(gz), s + "gl": npy(gl), s + "out": npy(out),
                    s + "ld": npy(ld), s + "gx": npy(x.grad)})
        for k, v in layer.named_parameters():
            rec[s + "gw." + k] = npy(v.grad)
    rec.update(sd_np(layer, "rnvp.sd."))
    # Planar / Radial
    d, N = 16, 20
    torch.manual_seed(60)
    pl = ref_flows1.Planar(d)
    rd = ref_flows1.Radial(d)
    init_radial(rd, d, 61)
    for tag, layer in (("planar.", pl), ("radial.", rd)):
        x = torch.randn(N, d, generator=gen(800)).requires_grad_()
        gz = torch.randn(N, d, generator=gen(810))
        gl = torch.randn(N, generator=gen(820))
        layer.zero_grad()
        out, ld = layer.forward(x)
        ((out * gz).sum() + (ld * gl).sum()).backward()
        rec.update({tag + "x": npy(x), tag + "gz": npy(gz), tag + "gl": npy(gl), tag + "out": npy(out),
                    tag + "ld": npy(ld), tag + "gx": npy(x.grad)})
        for k, v in layer.named_parameters():
            rec[tag + "gw." + k] = npy(v.grad)
        rec.update(sd_np(layer, tag + "sd."))
    np.savez_compressed(os.path.join(out_dir, "grads.npz"), **rec)


class _HarmonicSim:
    """Minimal duck-typed ``simulation`` (hmc.py:15-19, 39-40, 48-50, 63): U = 0.5*k*|q|^2,
    exact leapfrog.  Used to pin the HMC *driver* logic (velocity law, acceptance, bookkeeping)."""

    def __init__(self, nparticles, dim, k=1.0):
        self.nparticles = nparticles
        self.dim = dim
        self.k = k
        self.position = torch.linspace(-1, 1, nparticles * dim)
        self.velocity = torch.zeros(nparticles * dim)

    def get_position(self):
        return self.position

    def get_potential(self):
        return 0.5 * self.k * torch.sum(self.position ** 2)

    def set_position(self, p):
        self.position = p.flatten().clone()

    def set_velocity(self, v):
        self.velocity = v.flatten().clone()

    def integration_step(self, path_len, dt):
        q, v = self.position, self.velocity
        f = -self.k * q
        for _ in range(path_len):
            v = v + 0.5 * dt * f
            q = q + dt * v
            f = -self.k * q
            v = v + 0.5 * dt * f
        self.position, self.velocity = q, v
        return q, self.get_potential()


def fix_hmc(out_dir):
    import contextlib
    import io
    sim = _HarmonicSim(4, 3, k=2.0)
    torch.manual_seed(1234)
    h = ref_hmc.HMC(sim, path_len=5, dt=0.3, dim=3, beta=1.5)
    with contextlib.redirect_stdout(io.StringIO()):
        pos, pot, logp, acc = h.hmc(epochs=12)
    rec = {"positions": npy(pos), "potentials": npy(pot), "last_logp": npy(logp), "accept": acc,
           "nparticles": 4, "dim": 3, "k": 2.0, "path_len": 5, "dt": 0.3, "beta": 1.5, "seed": 1234,
           "epochs": 12}
    np.savez_compressed(os.path.join(out_dir, "hmc.npz"), **rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    torch.set_num_threads(1)
    fix_nsfcl(a.out, "nsfcl_d64.npz", size=32, dim=2, K=8, B=3.0, H=24, masks=[[0], [1]], N=64, seed=0)
    fix_nsfcl(a.out, "nsfcl_d64_stress.npz", size=32, dim=2, K=8, B=3.0, H=24, masks=[[1]], N=64, seed=10,
              scale_last=5.0)
    fix_nsfcl(a.out, "nsfcl_lj38.npz", size=38, dim=3, K=8, B=4.0, H=16,
              masks=[[0], [1], [2], [0, 1], [1, 2], [0, 2]], N=12, seed=20, in_std=2.0)
    fix_nsfcl(a.out, "nsfcl_k32.npz", size=4, dim=3, K=32, B=3, H=16, masks=[[1]], N=64, seed=30, scale_last=4.0)
    fix_rqs_function(a.out)
    fix_realnvp(a.out)
    fix_planar_radial(a.out)
    fix_models(a.out)
    fix_grads(a.out)
    fix_hmc(a.out)
    fix_nsf_ar(a.out)
    fix_checkpoint(a.out)
    fix_systems(a.out)
    fix_bar(a.out)
    tot = sum(os.path.getsize(os.path.join(a.out, f)) for f in os.listdir(a.out))
    print("wrote", sorted(os.listdir(a.out)), "total bytes", tot)


if __name__ == "__main__":
    main()
