"""The B-native boundary: the reference's pybind `_backend` modules served by libsanerf_b200
(sanerf_hq_b200/native_backend.py).  The reference's OWN Python -- gridencoder/grid.py, shencoder/sphere_harmonics.py,
freqencoder/freq.py, nerf/network.py, nerf/renderer.py, byte-compiled and unmodified under oracle/_ref/bytecode -- runs on this
repo's kernels and is compared with the same Python on the reference's compiled kernels (oracle/_ref/*.so).
"""
import importlib

import pytest
import torch

from helpers import O

DEV = "cuda"


def _R(backend):
    from oracle import ref_runtime as R
    if not R.available(backend):
        pytest.skip("oracle/_ref is not staged (python oracle/stage_ref.py; python oracle/build_ref.py)")
    return R


def test_modules_export_the_reference_pybind_names_and_the_reference_binds_to_them():
    """CPU: the 8 functions of the three bindings.cpp files, and the reference's `import _gridencoder as _backend` picks them up."""
    from sanerf_hq_b200 import native_backend as nb
    mods = nb.modules()
    assert sorted(mods) == ["_freqencoder", "_gridencoder", "_shencoder"]
    assert sorted(n for n in vars(mods["_gridencoder"]) if not n.startswith("__")) == \
        ["grad_total_variation", "grad_weight_decay", "grid_encode_backward", "grid_encode_forward"]
    assert sorted(n for n in vars(mods["_shencoder"]) if not n.startswith("__")) == ["sh_encode_backward", "sh_encode_forward"]
    assert sorted(n for n in vars(mods["_freqencoder"]) if not n.startswith("__")) == ["freq_encode_backward", "freq_encode_forward"]
    R = _R("native")
    with R.env("native"):
        grid = importlib.import_module("gridencoder.grid")
        sh = importlib.import_module("shencoder.sphere_harmonics")
        fr = importlib.import_module("freqencoder.freq")
        assert grid.__file__.startswith(R.PYC)                      # the reference's module, not this repo's shim
        for m, name in ((grid, "_gridencoder"), (sh, "_shencoder"), (fr, "_freqencoder")):
            assert m._backend.__name__ == name and "sanerf_hq_b200.native_backend" in (m._backend.__doc__ or "")
        # no CPU path: the reference's CHECK_CUDA behaviour (a RuntimeError), not a silent fallback
        enc = grid.GridEncoder(num_levels=2, log2_hashmap_size=8, desired_resolution=32)
        with pytest.raises(RuntimeError):
            enc(torch.rand(4, 3))
    # outside the environment the names are gone again
    import sys
    assert "_gridencoder" not in sys.modules


@pytest.mark.gpu
def test_reference_encoder_classes_on_native_kernels_match_the_reference_kernels():
    """The reference's GridEncoder / SHEncoder / FreqEncoder classes, forward + backward + regularisers, on both backends."""
    Rn, Rc = _R("native"), _R("cuda")
    g = torch.Generator().manual_seed(3)
    x = torch.rand(5000, 3, generator=g).to(DEV) * 2 - 1          # grid.py:156 maps [-bound, bound] -> [0, 1]
    d = torch.nn.functional.normalize(torch.randn(5000, 3, generator=g), dim=-1).to(DEV)
    go = torch.randn(5000, 16 * 2, generator=g).to(DEV)
    out = {}
    for backend, R in (("cuda", Rc), ("native", Rn)):
        with R.env(backend):
            grid = importlib.import_module("gridencoder.grid")
            sh = importlib.import_module("shencoder.sphere_harmonics")
            fr = importlib.import_module("freqencoder.freq")
            torch.manual_seed(0)
            enc = grid.GridEncoder(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048).to(DEV)
            with torch.no_grad():
                enc.embeddings.uniform_(-1, 1, generator=torch.Generator(device=DEV).manual_seed(1))
            xi = x.clone().requires_grad_(True)
            y = enc(xi, bound=1)
            y.backward(go)
            g_emb = enc.embeddings.grad.clone()
            g_x = xi.grad.clone()
            torch.manual_seed(7)                                  # TV draws its sample points with torch.rand (grid.py:183)
            enc.grad_total_variation(weight=1e-3, B=20000)
            enc.grad_weight_decay(weight=0.1)
            g_reg = enc.embeddings.grad.clone()
            she = sh.SHEncoder(input_dim=3, degree=4)
            di = d.clone().requires_grad_(True)
            ys = she(di)
            ys.sum().backward()
            fe = fr.FreqEncoder(input_dim=3, degree=4)
            fi = x.clone().requires_grad_(True)
            yf = fe(fi)
            (yf * yf).sum().backward()
            out[backend] = dict(y=y.detach(), g_emb=g_emb, g_x=g_x, g_reg=g_reg, ys=ys.detach(), g_d=di.grad.clone(), yf=yf.detach(),
                                g_f=fi.grad.clone())
    a, b = out["cuda"], out["native"]
    assert torch.equal(a["y"], b["y"])                            # grid forward: bit-exact (same FMA order)
    for k, tol in (("ys", 4e-6), ("g_emb", 1e-4), ("g_x", 1e-4), ("g_reg", 1e-4), ("g_d", 1e-4), ("yf", 1e-5), ("g_f", 1e-4)):
        scale = float(a[k].abs().max())
        assert scale > 0, k
        err = float((a[k] - b[k]).abs().max())
        assert err <= tol * scale, f"{k}: {err:.3e} vs scale {scale:.3e}"     # atomics / fast-math sin: summation-order noise


@pytest.mark.gpu
def test_reference_network_and_renderer_on_native_kernels():
    """The reference's NeRFNetwork + NeRFRenderer.render (its Python, unmodified) on this repo's kernels: eval frame and one
    training-mode forward/backward against the same Python on the reference's kernels."""
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    Rn, Rc = _R("native"), _R("cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    opt, opt_rgb = O.default_opt(), O.default_opt()
    opt.with_mask = True                                          # eval frame with the object head; rgb-stage model for training
    params, _ = O.make_params(opt, O.default_specs(2), seed=5)
    params_rgb, _ = O.make_params(opt_rgb, O.default_specs(2), seed=5)
    ro, rd = get_rays(orbit_pose(3), lego_intrinsics(64, 64), 64, 64)
    ro, rd = ro.to(DEV), rd.to(DEV)
    res = {}
    for backend, R in (("cuda", Rc), ("native", Rn)):
        model = R.build_network(opt, params, device=DEV, backend=backend)
        with R.env(backend):
            with torch.no_grad():
                ev = model.render(ro, rd, staged=True, bg_color=1, perturb=False, return_mask=1)
        model = R.build_network(opt_rgb, params_rgb, device=DEV, backend=backend)
        with R.env(backend):
            model.train()
            torch.manual_seed(13)                                 # perturb=True: same random stream on both sides
            tr = model.render(ro[:1024], rd[:1024], staged=False, bg_color=1, perturb=True, update_proposal=True)
            loss = tr["image"].square().mean() + tr["proposal_loss"] + 0.02 * tr["distort_loss"]
            loss.backward()
            grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        res[backend] = (ev, float(loss), grads)
    (ea, la, ga), (eb, lb, gb) = res["cuda"], res["native"]
    for k in ("image", "depth", "weights_sum", "instance_mask_logits"):
        a, b = ea[k].double(), eb[k].double()
        floor = max(1e-3, 0.1 * float(a.square().mean().sqrt())) if k == "instance_mask_logits" else 1e-3
        e = float(((a - b).abs() / a.abs().clamp(min=floor)).max())
        assert e <= 1e-4, f"{k}: {e:.3e}"                         # same Python, same cuBLAS; encoders bit-exact forward
    assert abs(la - lb) <= 1e-5 * abs(la)
    assert set(ga) == set(gb) and len(ga) > 10
    for n in ga:
        scale = float(ga[n].abs().max())
        assert float((ga[n] - gb[n]).abs().max()) <= 2e-3 * scale + 1e-12, n
