"""Runtime switches of the mirror modules."""
import os

# Orbit propagator inside the dynamics residual.  The reference picks by torch.cuda.is_available()
# (BA_filtering.py:16-19): CPU `predict` = 1 s RK4 steps, `predict_gpu` = steps of up to 100 s, which
# differ by metres.  Parity is pinned to the CPU reference, so "step1s" is the default.
propagator = os.environ.get("VINSAT_PROPAGATOR", "step1s")   # "step1s" | "skip100"
device = int(os.environ.get("VINSAT_DEVICE", os.environ.get("LOCAL_RANK", "0")))


_warned = False


def mode():
    """Propagator mode for the library.  Says once (on stderr, through `warnings`) when the default is in use: the
    reference running on this same CUDA host would take `predict_gpu` (100 s steps), i.e. differ by metres."""
    global _warned
    from . import _lib
    if propagator != "skip100" and "VINSAT_PROPAGATOR" not in os.environ and not _warned:
        _warned = True
        import warnings
        warnings.warn("vinsat_b200: using the reference's CPU propagator (`predict`, 1 s RK4 steps), the one parity is pinned to; "
                      "the reference itself would pick `predict_gpu` (100 s steps) on a CUDA host -- set VINSAT_PROPAGATOR=skip100 "
                      "(or vinsat_b200.config.propagator) to mirror that", stacklevel=2)
    return _lib.MODE_SKIP100 if propagator == "skip100" else _lib.MODE_STEP1S
