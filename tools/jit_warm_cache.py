#!/usr/bin/env python
"""Fill the on-disk cache of run-time specialised kernels (mbt_gym_b200/_jit_cache/) WITHOUT a GPU: every configuration
of the committed fixtures (both precisions, float32 I/O, the fused-rollout kernels the tests use) and of bench.py's
workloads.  NVRTC cross-compiles for sm_100a here; the cache travels to the GPU box with the tree, so the first
`mbt_create` there loads a cubin instead of compiling one.  `python tools/jit_warm_cache.py`  (also run by build())."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(verbose=True):
    from mbt_gym_b200 import _abi, _lib
    from tests.helpers import Golden, copy_config, golden_names

    t0, n = time.time(), 0
    for name in golden_names():
        g = Golden(name)
        for prec in (_abi.MBT_F64, _abi.MBT_F32):
            cfg = g.config(prec)
            _lib.jit_precompile(cfg, 0)
            n += 1
            for pol in (_abi.MBT_POL_FIXED, _abi.MBT_POL_AVELLANEDA_STOIKOV):
                if pol == _abi.MBT_POL_AVELLANEDA_STOIKOV and cfg.dynamics != _abi.MBT_DYN_LIMIT:
                    continue
                _lib.jit_precompile(cfg, 1, pol)
                n += 1
        _lib.jit_precompile(copy_config(g.config(_abi.MBT_F64), io_precision=_abi.MBT_IO_F32), 0)
        n += 1
    if verbose:
        cache = os.path.join(ROOT, "mbt_gym_b200", "_jit_cache")
        print(f"{n} kernels requested, {len(os.listdir(cache)) if os.path.isdir(cache) else 0} cubins in {cache} "
              f"({time.time() - t0:.1f} s)")


if __name__ == "__main__":
    main()
