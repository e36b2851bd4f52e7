"""CPU oracle for the SANeRF-HQ render hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product path (sanerf_hq_b200/, gridencoder/, shencoder/,
freqencoder/, nerf/) never does.
"""
