"""chess2rt_b200 — B200-native (sm_100a) implementation of Chess2RT's per-pixel render loop.

The product is ``libc2rt.so`` (hand-written CUDA kernels behind the C ABI of ``include/c2rt.h``)
plus ``libc2rt_host.so`` (C++ mirror of the reference's scene API: loaders, object model, scene
flattener, renderer entry points).  This Python package is a thin ctypes harness over those two
libraries for tests and ``bench.py``; it contains no rendering code and no CPU fallback — importing
:mod:`chess2rt_b200.api` raises if the libraries have not been built.
"""
from .api import (  # noqa: F401
    C2rtError, HostScene, Stats, band_rows_owned, device_count, init, lib, host_lib, render_device,
    read_ray_counters, deinterleave, measure_fma_peak, srgb_table, rng_u31, shutdown,
)
