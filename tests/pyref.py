"""Loader of the reference's OWN, unmodified Python layers (test infrastructure).

oracle/Makefile (`make -C oracle pyref`, part of `ref`; run by __graft_entry__.build() wherever /root/reference exists)
stages two files byte for byte under oracle/_ref/pyref/ -- build output like libgof_ref.so: git-ignored, shipped to the
GPU box with the tree:

  pyref/diff_gof_rasterization/__init__.py   <- RAST/diff_gof_rasterization/__init__.py   (autograd bridge + nn.Module)
  pyref/ref_gaussian_renderer.py             <- src/gaussian_renderer/__init__.py          (render_predicted_more_v2_gof)

`reference_rasterizer_package()` imports the first one as a package whose `_C` extension is the product's `_C` shim
(`from . import _C`, :15, resolves through sys.modules), i.e. the reference's `_RasterizeGaussians` /
`GaussianRasterizer_GOF` running over libgof_b200.  `reference_renderer(over)` imports the second one with
`diff_gof_rasterization` (:10) bound to either the product's drop-in package or to that reference package.
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")
RAST_INIT = os.path.join(PYREF, "diff_gof_rasterization", "__init__.py")
RENDERER = os.path.join(PYREF, "ref_gaussian_renderer.py")


def available() -> bool:
    return os.path.exists(RAST_INIT) and os.path.exists(RENDERER)


def _load(name, path, package_dir=None):
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[package_dir] if package_dir else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_rasterizer_package():
    """The reference's diff_gof_rasterization/__init__.py, unmodified, over the product's `_C`."""
    name = "ref_diff_gof_rasterization"
    if name in sys.modules:
        return sys.modules[name]
    from f3d_gaus_b200.diff_gof_rasterization import _C
    sys.modules[name + "._C"] = _C
    return _load(name, RAST_INIT, os.path.dirname(RAST_INIT))


def reference_renderer(over: str):
    """The reference's src/gaussian_renderer/__init__.py, unmodified.  over = 'dropin': on the product's
    diff_gof_rasterization package (f3d_gaus_b200.install_drop_in); over = 'refpy': on the reference's own Python
    package, which in turn sits on the product's `_C`."""
    import f3d_gaus_b200
    saved = sys.modules.get("diff_gof_rasterization")
    try:
        if over == "dropin":
            f3d_gaus_b200.install_drop_in()
        else:
            sys.modules["diff_gof_rasterization"] = reference_rasterizer_package()
        return _load("ref_gaussian_renderer_" + over, RENDERER)
    finally:
        if saved is not None:
            sys.modules["diff_gof_rasterization"] = saved
        else:
            sys.modules.pop("diff_gof_rasterization", None)
