"""Drop-in for the reference's `nerf/renderer.py`.

`nerf/` deliberately has NO __init__.py: like the reference's it is a PEP 420 namespace package, so with
this repo ahead of the reference on sys.path, `nerf.renderer` / `nerf.network` resolve here while
`nerf.trainer`, `nerf.provider`, `nerf.utils`, `nerf.gui` still resolve to the reference's files and
inherit / call the new renderer unchanged.
"""
from sanerf_hq_b200.renderer import (NeRFRenderer, contract, distort_loss, near_far_from_aabb,  # noqa: F401
                                     proposal_loss, sample_pdf, uncontract)
