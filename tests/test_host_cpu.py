"""CPU (no GPU needed): host-side logic, C-ABI surface, drop-in module surface, no-fallback behaviour."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

from helpers import REPO, O, make_case

HEADER = os.path.join(REPO, "include", "sanerf_b200.h")

RGB_KEYS = ["aabb_train", "aabb_infer", "grid.embeddings", "grid.offsets", "grid_mlp.net.0.weight", "grid_mlp.net.1.weight",
            "grid_mlp.net.2.weight", "view_mlp.net.0.weight", "view_mlp.net.1.weight", "view_mlp.net.2.weight",
            "prop_encoders.0.embeddings", "prop_encoders.0.offsets", "prop_encoders.1.embeddings", "prop_encoders.1.offsets",
            "prop_mlp.0.net.0.weight", "prop_mlp.0.net.1.weight", "prop_mlp.1.net.0.weight", "prop_mlp.1.net.1.weight"]


def test_library_exports_every_declared_symbol():
    from sanerf_hq_b200 import _lib
    declared = set(re.findall(r"\b(sanerf_[a-z0-9_]+)\s*\(", open(HEADER).read()))
    declared -= {"sanerf_stream_t"}
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib._build.build())
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"libsanerf_b200.so lacks {missing}"
    # every declared symbol has a ctypes prototype, and vice versa
    assert declared == set(_lib.PROTOTYPES)
    L = _lib.load()
    assert L.sanerf_abi_version() == 2
    assert b"C must be" in L.sanerf_error_string(-3)


def test_ctypes_structs_match_the_c_header():
    """sizeof / offsetof of the POD structs as gcc sees the header == the ctypes mirrors."""
    from sanerf_hq_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "sanerf_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(sanerf_grid_t), sizeof(sanerf_model_t), sizeof(sanerf_render_args_t),
         offsetof(sanerf_model_t, grid), offsetof(sanerf_model_t, s_grid), offsetof(sanerf_model_t, aabb),
         offsetof(sanerf_model_t, u65), offsetof(sanerf_render_args_t, f_image));
  printf("%zu %zu\n", offsetof(sanerf_render_args_t, peer_depth), offsetof(sanerf_render_args_t, max_ctas));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        got = [int(v) for v in subprocess.check_output([os.path.join(d, "t")]).split()]
    want = [ctypes.sizeof(_lib.GridT), ctypes.sizeof(_lib.ModelT), ctypes.sizeof(_lib.RenderArgsT), _lib.ModelT.grid.offset,
            _lib.ModelT.s_grid.offset, _lib.ModelT.aabb.offset, _lib.ModelT.u65.offset, _lib.RenderArgsT.f_image.offset,
            _lib.RenderArgsT.peer_depth.offset, _lib.RenderArgsT.max_ctas.offset]
    assert got == want


def test_state_dict_keys_and_shapes_match_reference_checkpoints():
    from sanerf_hq_b200.network import NeRFNetwork
    opt = O.default_opt()
    m = NeRFNetwork(opt)
    assert sorted(m.state_dict().keys()) == sorted(RGB_KEYS)
    assert tuple(m.grid.embeddings.shape) == (6299960, 2)
    assert tuple(m.prop_encoders[0].embeddings.shape) == (383264, 2)
    assert tuple(m.prop_encoders[1].embeddings.shape) == (430080, 2)
    assert m.grid.offsets.dtype == torch.int32
    assert sum(p.numel() for p in m.parameters()) == 14236240
    sp = O.default_specs(2)
    assert torch.equal(m.grid.offsets, sp["grid"].offsets)
    ms = NeRFNetwork(O.default_opt(with_sam=True))
    assert len(ms.state_dict()) == 32 and tuple(ms.samvit_mlp[0].net[2].weight.shape) == (256, 256 + 163)
    assert tuple(ms.s_grid.embeddings.shape) == (5258512, 8)
    mm = NeRFNetwork(O.default_opt(with_mask=True))
    assert len(mm.state_dict()) == 23 and tuple(mm.mask_mlp[0].net[0].weight.shape) == (256, 143)
    # oracle-generated weights load strictly (same names / shapes)
    _, params, _ = make_case(with_sam=True)
    assert not ms.load_state_dict(params, strict=True).missing_keys


@pytest.mark.skipif(not os.path.isdir("/root/reference/nerf"), reason="reference tree not present")
def test_constructor_consumes_rng_like_the_reference():
    """torch.manual_seed(s); NeRFNetwork(opt) must give the reference's initial weights (checked in a
    subprocess so the reference's same-named packages cannot leak into this test session)."""
    code = r'''
import sys, types, torch
sys.path.insert(0, "/root/reference")
for n in ("mcubes", "trimesh", "torch_efficient_distloss"):
    sys.modules[n] = types.ModuleType(n)
sys.modules["torch_efficient_distloss"].eff_distloss = None
for n in ("_gridencoder", "_shencoder"):
    sys.modules[n] = types.ModuleType(n)
import warnings; warnings.filterwarnings("ignore")
from nerf.network import NeRFNetwork
sys.path.insert(0, sys.argv[1])
from oracle import render_oracle as O
torch.manual_seed(5)
m = NeRFNetwork(O.default_opt(with_sam=True, with_mask=True))
print(" ".join(f"{k}:{float(v.double().sum()):.10e}" for k, v in sorted(m.state_dict().items())))
'''
    ref = subprocess.check_output([sys.executable, "-c", code, REPO], stderr=subprocess.DEVNULL).decode().split()
    from sanerf_hq_b200.network import NeRFNetwork
    torch.manual_seed(5)
    m = NeRFNetwork(O.default_opt(with_sam=True, with_mask=True))
    mine = [f"{k}:{float(v.double().sum()):.10e}" for k, v in sorted(m.state_dict().items())]
    assert mine == ref


def test_drop_in_import_names():
    import activation
    import encoding
    import freqencoder
    import gridencoder
    import shencoder
    from gridencoder.grid import GridEncoder as G2
    from nerf.network import NeRFNetwork
    from nerf.renderer import NeRFRenderer, contract, near_far_from_aabb, sample_pdf  # noqa: F401
    assert gridencoder.GridEncoder is G2 and issubclass(NeRFNetwork, NeRFRenderer)
    enc, dim = encoding.get_encoder("hashgrid", desired_resolution=4096)
    assert dim == 32 and isinstance(enc, gridencoder.GridEncoder)
    enc, dim = encoding.get_encoder("sh", degree=4)
    assert dim == 16 and isinstance(enc, shencoder.SHEncoder)
    enc, dim = encoding.get_encoder("frequency", multires=6)
    assert dim == 39 and isinstance(enc, freqencoder.FreqEncoder)
    assert callable(activation.trunc_exp)
    with pytest.raises(NotImplementedError):
        encoding.get_encoder("nope")
    with pytest.raises(AssertionError):
        shencoder.SHEncoder(degree=9)


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors (the reference CHECK_CUDA's, gridencoder.cu:468)."""
    from sanerf_hq_b200.encoders import FreqEncoder, GridEncoder, SHEncoder
    from sanerf_hq_b200.network import NeRFNetwork
    g = GridEncoder(num_levels=2, desired_resolution=32)
    with pytest.raises(RuntimeError):
        g(torch.rand(4, 3))
    with pytest.raises(RuntimeError):
        SHEncoder()(torch.rand(4, 3))
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            FreqEncoder()(torch.rand(4, 3))     # reference moves the input to CUDA (freq.py:22) -> fails without a GPU
    m = NeRFNetwork(O.default_opt(), num_levels=4, hidden_dim=16).eval()
    with pytest.raises(RuntimeError), torch.no_grad():
        m.render(torch.rand(8, 3), torch.rand(8, 3), staged=True)
    with pytest.raises(ValueError):
        g.grad_weight_decay()                   # grad is None (grid.py:202-203)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(REPO, "sanerf_hq_b200")
    for root in (pkg, os.path.join(REPO, "gridencoder"), os.path.join(REPO, "shencoder"), os.path.join(REPO, "freqencoder"),
                 os.path.join(REPO, "nerf")):
        for dp, _, files in os.walk(root):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dp, f)).read()
                    assert "oracle" not in txt.replace("oracle/", "").lower() or f in (), (dp, f)


def test_torch_helpers_match_oracle_on_cpu():
    """The composed path's pure-torch helpers (contract / sample_pdf / near_far) are device-agnostic torch code;
    check them against the oracle here so a GPU failure can be localised."""
    from sanerf_hq_b200 import renderer as R
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(1000, 7, 3, generator=g) - 0.5) * 20
    assert torch.equal(R.contract(x), O.contract(x))
    assert torch.allclose(R.uncontract(R.contract(x)), x, rtol=1e-4, atol=1e-4)
    w = torch.rand(50, 128, generator=g) ** 4
    bins = torch.linspace(0, 1, 129).expand(50, -1)
    assert torch.equal(R.sample_pdf(bins, w, 65), O.sample_pdf(bins, w, 65)[0])
    o = torch.randn(100, 3, generator=g)
    d = torch.randn(100, 3, generator=g)
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1])
    n1, f1 = R.near_far_from_aabb(o, d, aabb, 0.2)
    n2, f2 = O.near_far_from_aabb(o, d, aabb, 0.2)
    assert torch.equal(n1, n2) and torch.equal(f1, f2)


def test_head_entry_points_validate_arguments_before_any_launch():
    """The C ABI returns a status instead of launching on bad arguments (the reference never checks): these paths return before
    the first CUDA call, so they can be exercised without a GPU -- the dummy pointers are never dereferenced."""
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    g = _lib.GridT()
    g.embeddings = 0x1000
    g.num_levels, g.level_dim = 16, 8
    for l in range(17):
        g.offset[l] = l * 524288
    for l in range(16):
        g.res[l] = 16 + l
    P = ctypes.c_void_p

    def mask_head(rec, n_inst, n_rays=4, w0=0x3000):
        return lib.sanerf_mask_head(P(rec), P(0x2000), ctypes.byref(g), P(w0), P(0x4000), P(0x5000), n_inst, n_rays, P(0x6000), P(0x7000), None)

    assert mask_head(0x10000, 2, n_rays=0) == 0          # empty batch: nothing to do
    assert mask_head(0x10004, 2) != 0                    # records not 16-byte aligned (TMA bulk copies)
    assert mask_head(0x10000, 0) != 0 and mask_head(0x10000, 17) != 0    # 1 <= n_inst <= 16
    assert mask_head(0x10000, 2, w0=0) != 0              # null weight
    g.level_dim = 2
    assert mask_head(0x10000, 2) != 0                    # the object grid has 8 channels per level
    g.level_dim, g.num_levels = 8, 12
    assert mask_head(0x10000, 2) != 0                    # ... and 16 levels
    assert lib.sanerf_error_string(mask_head(0x10004, 2))   # every status has a message
    ws = (ctypes.c_void_p * 5)(*[0x1000] * 5)
    assert lib.sanerf_samvit_mlp(P(0x1000), ws, ws, P(0x1000), P(0x1000), 0, P(0x1000), P(0x1000), None) == 0
    assert lib.sanerf_samvit_mlp(None, ws, ws, P(0x1000), P(0x1000), 128, P(0x1000), P(0x1000), None) != 0
    bad = (ctypes.c_void_p * 5)(0x1000, 0x1000, None, 0x1000, 0x1000)
    assert lib.sanerf_samvit_mlp(P(0x1000), bad, ws, P(0x1000), P(0x1000), 128, P(0x1000), P(0x1000), None) != 0
