"""Run the REFERENCE ITSELF from the staged build outputs under oracle/_ref.  TEST INFRASTRUCTURE ONLY.

    oracle/_ref/bytecode/...      the reference's unmodified Python, byte-compiled by oracle/stage_ref.py
    oracle/_ref/_gridencoder.so   the reference's unmodified CUDA kernels, compiled by oracle/build_ref.py
    oracle/_ref/_shencoder.so, _freqencoder.so

Two backends for the CUDA-only pybind modules that gridencoder/grid.py:9-12 and shencoder/sphere_harmonics.py:9-12 import:
  "cuda"  the reference's compiled kernels  -> the GPU oracle of SURVEY.md 8c/8d and the reference GPU baseline (R-GPU)
  "cpu"   fake modules backed by the C restatement (oracle/sanerf_oracle.c)  -> the reference's Python on the host cores
          (the reference has no CPU encoder); used by bench.py --impl reference.

  "native" the reference's Python over THIS REPO's kernels through the pybind-compatible modules of
          sanerf_hq_b200/native_backend.py -> the B-native boundary test: the reference's own GridEncoder /
          SHEncoder / NeRFNetwork / renderer, unmodified, on libsanerf_b200.

The reference's top-level module names (nerf.renderer, encoding, activation, gridencoder, shencoder, freqencoder) are the
same as this repo's drop-in shims, so the reference is imported inside `env()`, a context manager that swaps those names in
sys.modules / sys.path in and out.  Everything that may trigger an import inside the reference (model construction:
encoding.get_encoder imports lazily) must run inside `with env():`.

Only tests/ and bench.py's reference legs import this module.
"""
import contextlib
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
PYC = os.path.join(REFDIR, "bytecode")
EXT = ".bin"


class _RefFinder(importlib.abc.MetaPathFinder):
    """Resolves the reference's module names to the staged bytecode files (first on sys.meta_path while `env()` is active)."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] not in ("nerf", "encoding", "activation", "gridencoder", "shencoder", "freqencoder"):
            return None
        rel = os.path.join(PYC, *fullname.split("."))
        init = os.path.join(rel, "__init__" + EXT)
        if os.path.exists(init):
            return importlib.util.spec_from_file_location(fullname, init, loader=importlib.machinery.SourcelessFileLoader(fullname, init),
                                                          submodule_search_locations=[rel])
        if os.path.exists(rel + EXT):
            return importlib.util.spec_from_file_location(fullname, rel + EXT,
                                                          loader=importlib.machinery.SourcelessFileLoader(fullname, rel + EXT))
        if os.path.isdir(rel):       # `nerf` has no __init__.py in the reference: a namespace package
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [rel]
            return spec
        return None


_FINDER = _RefFinder()

_REF_ROOTS = ("nerf", "encoding", "activation", "gridencoder", "shencoder", "freqencoder", "_gridencoder", "_shencoder",
              "_freqencoder", "mcubes", "trimesh", "torch_efficient_distloss", "torch_ema", "imageio", "matplotlib", "wandb",
              "tensorboardX")
_state = {"active": None, "depth": 0, "orig": None, "mods": {}}     # mods: backend -> the reference module set loaded with it


def _is_ref_name(name):
    return name.split(".")[0] in _REF_ROOTS


def available(backend="cuda"):
    """True when the staged build outputs this backend needs are present."""
    ok = os.path.exists(os.path.join(PYC, "nerf", "renderer" + EXT)) and os.path.exists(os.path.join(PYC, "encoding" + EXT))
    if backend == "cuda":
        ok = ok and all(os.path.exists(os.path.join(REFDIR, n + ".so")) for n in ("_gridencoder", "_shencoder"))
    if backend == "native":
        from sanerf_hq_b200 import _lib
        ok = ok and os.path.exists(_lib.lib_path())
    return ok


def eff_distloss(w, m, interval):
    """Published definition of torch_efficient_distloss.eff_distloss (requirements.txt:21, not installed here):
    sum_ij w_i w_j |m_i - m_j| + 1/3 sum_i w_i^2 delta_i, mean over rays, in the O(T) prefix-sum form of the package."""
    loss_uni = (1 / 3) * (interval * w.pow(2)).sum(dim=-1).mean()
    wm = w * m
    w_cum, wm_cum = w.cumsum(dim=-1), wm.cumsum(dim=-1)
    loss_bi = 2 * (wm[..., 1:] * w_cum[..., :-1] - w[..., 1:] * wm_cum[..., :-1]).sum(dim=-1).mean()
    return loss_bi + loss_uni


def _stub_modules(backend):
    """Modules the reference imports that this image lacks and the render path never touches, plus the
    fake pybind backends for backend == "cpu"."""
    mods = {}

    def stub(name, **attrs):
        try:
            if name in ("wandb", "imageio", "matplotlib"):
                raise ImportError   # never drag heavyweight optional packages into a test process
            mods[name] = importlib.import_module(name)
        except ImportError:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            mods[name] = m
        return mods[name]

    stub("mcubes")
    stub("trimesh")
    stub("torch_efficient_distloss", eff_distloss=eff_distloss)
    stub("imageio")
    stub("wandb")
    stub("tensorboardX")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl = stub("matplotlib", pyplot=plt)
    mods["matplotlib.pyplot"] = getattr(mpl, "pyplot", plt)

    class ExponentialMovingAverage:   # torch_ema's interface as nerf/trainer.py uses it (:138-140, 1138, 1556, 1730)
        def __init__(self, parameters, decay):
            import torch
            self.decay, self.params = decay, [p for p in parameters]
            self.shadow = [p.detach().clone() for p in self.params]
            self.backup = None
            self._torch = torch

        def update(self):
            with self._torch.no_grad():
                for s, p in zip(self.shadow, self.params):
                    s.mul_(self.decay).add_(p.detach(), alpha=1 - self.decay)

        def store(self):
            self.backup = [p.detach().clone() for p in self.params]

        def copy_to(self):
            with self._torch.no_grad():
                for s, p in zip(self.shadow, self.params):
                    p.copy_(s)

        def restore(self):
            with self._torch.no_grad():
                for b, p in zip(self.backup, self.params):
                    p.copy_(b)
            self.backup = None

        def state_dict(self):
            return {"decay": self.decay, "shadow": self.shadow}

        def load_state_dict(self, sd):
            self.decay, self.shadow = sd["decay"], sd["shadow"]

    stub("torch_ema", ExponentialMovingAverage=ExponentialMovingAverage)

    if backend == "cpu":
        from . import kernels as K
        ge = types.ModuleType("_gridencoder")

        def gef(inputs, embeddings, offsets, outputs, B, D, C, L, max_level, S, H, dy_dx, gridtype, align_corners, interp):
            K.grid_encode_forward(inputs, embeddings, offsets, B, D, C, L, max_level, S, H, dy_dx, gridtype, align_corners, interp,
                                  outputs=outputs)

        ge.grid_encode_forward = gef
        ge.grid_encode_backward = K.grid_encode_backward      # same argument order as gridencoder.h:13
        ge.grad_total_variation = K.grad_total_variation
        ge.grad_weight_decay = K.grad_weight_decay
        sh = types.ModuleType("_shencoder")
        sh.sh_encode_forward = lambda inputs, outputs, B, D, C, dy_dx: K.sh_encode_forward(inputs, B, C, outputs=outputs)
        mods["_gridencoder"], mods["_shencoder"] = ge, sh
    if backend == "native":
        from sanerf_hq_b200 import native_backend
        mods.update(native_backend.modules())
    return mods


@contextlib.contextmanager
def env(backend="cuda"):
    """Inside: `import nerf.renderer`, `encoding`, `gridencoder`, ... resolve to the REFERENCE (oracle/_ref/bytecode) and
    `_gridencoder` / `_shencoder` to the chosen backend.  Outside: to whatever they resolved to before (this repo's shims).
    Re-entrant for the same backend; each backend keeps its own set of reference modules (grid.py binds `_backend` at import
    time), so the "cuda" and the "cpu" reference can live in one process, one at a time."""
    if _state["active"] not in (None, backend):
        raise RuntimeError(f"oracle.ref_runtime: env('{_state['active']}') is active; leave it before entering env('{backend}')")
    if not available(backend):
        raise RuntimeError("oracle.ref_runtime: oracle/_ref is not staged (python oracle/stage_ref.py; python oracle/build_ref.py)")
    _state["depth"] += 1
    if _state["depth"] == 1:
        _state["active"] = backend
        if backend not in _state["mods"]:
            _state["mods"][backend] = _stub_modules(backend)
        _state["orig"] = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_ref_name(k)}
        sys.modules.update(_state["mods"][backend])
        sys.meta_path.insert(0, _FINDER)
        if backend == "cuda":
            sys.path.insert(0, REFDIR)            # the compiled pybind modules _gridencoder / _shencoder / _freqencoder
    try:
        yield
    finally:
        _state["depth"] -= 1
        if _state["depth"] == 0:
            _state["mods"][backend] = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_ref_name(k)}
            sys.modules.update(_state["orig"])
            _state["orig"], _state["active"] = None, None
            if _FINDER in sys.meta_path:
                sys.meta_path.remove(_FINDER)
            if REFDIR in sys.path:
                sys.path.remove(REFDIR)


def modules(backend="cuda"):
    """(nerf.renderer, nerf.network) of the reference."""
    with env(backend):
        r = importlib.import_module("nerf.renderer")
        n = importlib.import_module("nerf.network")
    assert r.__file__.startswith(PYC) and n.__file__.startswith(PYC), (r.__file__, n.__file__)
    return r, n


def build_network(opt, state_dict=None, device="cuda", backend="cuda"):
    """The reference's NeRFNetwork(opt) (nerf/network.py:85), optionally loaded with `state_dict` (strict), in eval mode."""
    with env(backend):
        net = importlib.import_module("nerf.network")
        model = net.NeRFNetwork(opt)
        if state_dict is not None:
            res = model.load_state_dict(state_dict, strict=True)
            assert not res.missing_keys and not res.unexpected_keys
        return model.eval().to(device)


def render(model, rays_o, rays_d, backend="cuda", **kw):
    """model.render(...) of a reference model (nerf/renderer.py:185-219) inside the reference's module environment."""
    with env(backend):
        return model.render(rays_o, rays_d, **kw)


def render_features_by_rows(model, rays_o, rays_d, W, rows_per_call=5, backend="cuda", **kw):
    """800x800 feature render the only way the reference API allows it (staged + return_feats is
    impossible): non-staged calls on chunks of `rows_per_call` image rows with H=rows, W=W, concatenated."""
    import torch
    n = rays_o.shape[0]
    step = rows_per_call * W
    outs = {}
    with env(backend):
        for head in range(0, n, step):
            tail = min(n, head + step)
            part = model.render(rays_o[head:tail], rays_d[head:tail], staged=False, return_feats=1, H=(tail - head) // W, W=W, **kw)
            for k, v in part.items():
                if torch.is_tensor(v):
                    outs.setdefault(k, []).append(v.reshape(tail - head, -1) if k == "samvit" else v)
    return {k: torch.cat(v, 0) for k, v in outs.items()}
