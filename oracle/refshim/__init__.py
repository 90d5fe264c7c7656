"""TEST INFRASTRUCTURE: load the UNMODIFIED reference sources from /root/reference in a
container that has neither xarray nor h5py nor the plotting stack.

`load_reference()` registers stand-ins in sys.modules (xarray -> xarray_standin,
h5py -> the repo's pure-Python HDF5 reader, empty stubs for the plotting stack that
transport.py imports at module top and never uses on the stepping path) and
returns the reference package with its real linalg / transport / utilities /
constituents / io.hdf modules executed from where they lie.  Nothing is copied.
Only tools/make_golden.py uses this; /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path
from unittest import mock

REFERENCE_SRC = Path("/root/reference/src")


def load_reference(src: Path = REFERENCE_SRC):
    if "clearwater_riverine" in sys.modules and getattr(sys.modules["clearwater_riverine"], "_refshim", False):
        return sys.modules["clearwater_riverine"]
    if not (src / "clearwater_riverine" / "linalg.py").is_file():
        raise FileNotFoundError(f"reference sources not found under {src}")

    from . import xarray_standin
    sys.modules["xarray"] = xarray_standin

    repo_root = Path(__file__).resolve().parents[2]
    if str(repo_root) not in sys.path:
        sys.path.insert(0, str(repo_root))
    from clearwater_riverine_b200.io import hdf5_mini
    h5 = types.ModuleType("h5py")
    h5.File = lambda path, mode="r": hdf5_mini.File(path)
    sys.modules["h5py"] = h5

    for name in ["holoviews", "geoviews", "geopandas", "shapely", "shapely.geometry",
                 "matplotlib", "matplotlib.pyplot", "zarr", "netCDF4"]:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = mock.MagicMock(name=name)

    _tolerate_trailing_space_in_timestamps()

    pkg = types.ModuleType("clearwater_riverine")
    pkg.__path__ = [str(src / "clearwater_riverine")]
    pkg._refshim = True
    sys.modules["clearwater_riverine"] = pkg
    for sub in ["variables", "linalg", "utilities", "io.hdf", "io.inputs", "io.outputs", "io.config",
                "mesh", "constituents", "transport"]:
        mod = importlib.import_module(f"clearwater_riverine.{sub}")
        setattr(pkg, sub.split(".")[0], sys.modules[f"clearwater_riverine.{sub.split('.')[0]}"])
    pkg.ClearwaterRiverine = sys.modules["clearwater_riverine.transport"].ClearwaterRiverine
    return pkg


def _tolerate_trailing_space_in_timestamps():
    """HEC-RAS writes 'Time Date Stamp' as '01JAN2023 12:00:00 ' (trailing blank).  The
    pandas the reference was written against ignored it in io/hdf.py:156; pandas 3 raises.
    Strip it before parsing -- the parsed datetimes are identical."""
    import pandas as pd
    if getattr(pd.to_datetime, "_refshim", False):
        return
    orig = pd.to_datetime

    def to_datetime(arg, *a, **k):
        if isinstance(arg, pd.Series) and arg.dtype == object or str(getattr(arg, "dtype", "")) in ("str", "string"):
            try:
                arg = arg.str.strip()
            except Exception:
                pass
        return orig(arg, *a, **k)

    to_datetime._refshim = True
    pd.to_datetime = to_datetime
