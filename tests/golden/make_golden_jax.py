"""Regenerates the per-stage reference fixtures with the reference's OWN runtime: real `jax` (the reference pins jax==0.3.23,
requirements.txt:1), no NumPy stand-in.  Closes the third-party pin DESIGN.md section 4 leaves open: with the shim, `jax.random.*`
(Threefry, `normal` via XLA's erf_inv, `beta` via Marsaglia-Tsang with per-element key splitting) and XLA's float32 kernels are this
repo's own restatement; with real JAX they are the reference's.

    pip install "jax==0.3.23" "jaxlib<=0.3.25"          # on a machine with network access
    MPCMMD_REFERENCE_ROOT=/path/to/MPC-MMD python tests/golden/make_golden_jax.py

writes tests/golden/ref_stages_jax.npz (same schema as ref_stages.npz + `meta.jax_version`).  tests/test_reference_stages.py picks the
file up automatically when it exists and holds the CPU oracle (`-m "not gpu"`) and the CUDA stage entry points (`-m gpu`) to the same
1e-4 on it.  Not runnable in the build container (no jax wheel, no network); the shim-based ref_stages.npz stays the committed fixture.
"""
import os
import runpy
import sys

try:
    import jax
except ImportError:
    sys.exit("make_golden_jax.py needs the reference's runtime: pip install jax==0.3.23 (plus a matching jaxlib); "
             "use make_golden_ref.py for the NumPy stand-in")
if "jax_shim" in os.path.abspath(getattr(jax, "__file__", "") or ""):
    sys.exit("make_golden_jax.py: `jax` resolved to tests/golden/jax_shim -- remove it from PYTHONPATH")
os.environ["MPCMMD_GOLDEN_JAX"] = "real"
print("real jax", jax.__version__, "(reference pins 0.3.23)")
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "make_golden_ref.py"), run_name="__main__")
