"""Runtime switches of the mirror modules."""
import os

# Orbit propagator inside the dynamics residual.  The reference picks by torch.cuda.is_available()
# (BA_filtering.py:16-19): CPU `predict` = 1 s RK4 steps, `predict_gpu` = steps of up to 100 s, which
# differ by metres.  Parity is pinned to the CPU reference, so "step1s" is the default.
propagator = os.environ.get("VINSAT_PROPAGATOR", "step1s")   # "step1s" | "skip100"
device = int(os.environ.get("VINSAT_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def mode():
    from . import _lib
    return _lib.MODE_SKIP100 if propagator == "skip100" else _lib.MODE_STEP1S
