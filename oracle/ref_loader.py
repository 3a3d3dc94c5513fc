"""Import the UNMODIFIED VINSat reference (pure Python) so it can pin the oracle.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` in the build
container, where ``/root/reference`` is mounted read-only.  Nothing in the product
path (``vinsat_b200/``), in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may
import this: ``/root/reference`` does not exist on the GPU box.

The reference imports three modules that are not installed here
(``torch_scatter``, ``ipdb``, ``matplotlib.pyplot``); SURVEY.md section 8(c) shows that
stubbing them is enough to import ``BA.BA_utils``, ``BA.BA_filtering``,
``trajgen_pipe`` and ``od_pipe``.  The ``torch_scatter`` stub restates the
published semantics of ``scatter_sum`` / ``scatter_mean`` (segmented sum / mean
along ``dim`` into ``dim_size`` slots) as used at ``BA/BA_utils.py:1376-1382``.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("VINSAT_REF", "/root/reference")


def _stub_modules():
    import torch

    ts = types.ModuleType("torch_scatter")

    def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
        dim = dim % src.dim()
        size = list(src.shape)
        size[dim] = int(dim_size if dim_size is not None else int(index.max()) + 1)
        res = torch.zeros(size, dtype=src.dtype, device=src.device)
        return res.index_add_(dim, index.to(torch.long), src)

    def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
        dim = dim % src.dim()
        s = scatter_sum(src, index, dim, None, dim_size)
        ones = torch.ones(src.shape[dim], dtype=src.dtype, device=src.device)
        cnt = torch.zeros(s.shape[dim], dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index.to(torch.long), ones).clamp_(min=1)
        shape = [1] * s.dim()
        shape[dim] = -1
        return s / cnt.view(shape)

    ts.scatter_sum = scatter_sum
    ts.scatter_mean = scatter_mean
    sys.modules.setdefault("torch_scatter", ts)

    ipdb = types.ModuleType("ipdb")
    ipdb.set_trace = lambda *a, **k: None
    sys.modules.setdefault("ipdb", ipdb)

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "estimation", "BA"))


def load():
    """Returns a namespace with the reference modules (CPU `predict` path forced)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    import torch

    _stub_modules()
    # pin the oracle to the CPU propagator (`predict`, 1 s RK4 steps).
    torch.cuda.is_available = lambda: False
    est = os.path.join(REF_ROOT, "estimation")
    if est not in sys.path:
        sys.path.insert(0, est)
    import BA.BA_utils as BA_utils
    import BA.BA_filtering as BA_filtering
    import trajgen_pipe

    cwd = os.getcwd()
    os.chdir(est)  # od_pipe reads landmarks/intrinsics.csv relative to cwd
    try:
        import od_pipe
    finally:
        os.chdir(cwd)
    ns = types.SimpleNamespace(BA_utils=BA_utils, BA_filtering=BA_filtering,
                               trajgen_pipe=trajgen_pipe, od_pipe=od_pipe, est_dir=est)
    return ns


def load_satcam():
    """Import the UNMODIFIED ``sim/SatCam.py`` (and ``sim/orbit_gen.py``).  Its geometry is pure NumPy; the
    modules it imports but that are absent here (astropy, rasterio, pyproj) are only reached by
    ``get_corner_lonlats`` (astropy, SatCam.py:181), ``lonlat_to_pixel_coords`` (pyproj, :194-199) and the image
    windowing (:264-660).  Six empty stub modules make the import succeed; callers that need
    ``get_corner_lonlats`` patch the single astropy call with a stated closed form (tests/golden/make_golden_satcam.py).
    Returns (SatCam module, orbit_gen module, sim dir)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "sim")):
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    _stub_modules()
    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            try:
                __import__(name)
                return sys.modules[name]
            except Exception:
                m = types.ModuleType(name)
                sys.modules[name] = m
        for k, v in attrs.items():
            if not hasattr(m, k):
                setattr(m, k, v)
        return m

    class EarthLocation:                      # placeholder; patched by the caller where needed
        @staticmethod
        def from_geocentric(*a, **k):
            raise NotImplementedError("astropy is not installed; patch get_corner_lonlats")

    astropy = stub("astropy")
    astropy.coordinates = stub("astropy.coordinates", EarthLocation=EarthLocation)
    rasterio = stub("rasterio")
    rasterio.merge = stub("rasterio.merge", merge=None)
    rasterio.windows = stub("rasterio.windows", from_bounds=None, transform=None)
    stub("pyproj")
    sim = os.path.join(REF_ROOT, "sim")
    if sim not in sys.path:
        sys.path.insert(0, sim)
    import SatCam as SatCamMod
    import orbit_gen
    return SatCamMod, orbit_gen, sim
