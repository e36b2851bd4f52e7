"""sanerf_hq_b200 -- B200-native (sm_100a) implementation of the SANeRF-HQ volumetric-render hot path.

Layout:
  csrc/         hand-written CUDA kernels + the C ABI (include/sanerf_b200.h) -> lib/libsanerf_b200.so
  _lib.py       ctypes binding of the C ABI (no torch types cross the boundary)
  encoders.py   GridEncoder / SHEncoder / FreqEncoder  (reference gridencoder/, shencoder/, freqencoder/)
  renderer.py   NeRFRenderer.render/run                (reference nerf/renderer.py)
  network.py    NeRFNetwork, MLP, SkipConnMLP          (reference nerf/network.py)
  encoding.py, activation.py                           (reference encoding.py, activation.py)
  parallel.py   ray-sharded multi-GPU render + all-gather

The reference's import names are provided by thin top-level packages of this repo
(`gridencoder`, `shencoder`, `freqencoder`, `nerf.renderer`, `nerf.network`, `encoding`,
`activation`), so nerf/trainer.py and main.py run unchanged with this repo first on sys.path.
"""
__version__ = "0.1.0"
