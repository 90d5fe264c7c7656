"""Result writers for `ClearwaterRiverine.finalize(save=True, output_filepath=...)` ("next" row N4).

The reference hands its xarray Dataset to zarr or netCDF4 (reference io/outputs.py:10-89, transport.py:385-395);
neither library is a dependency here.  Same factory-by-extension shape, two self-contained formats:

  .npz  numpy archive: every mesh variable under its reference name (variables.py) plus `attrs_json`
  .nc   NetCDF-3 classic through scipy.io.netcdf_file -- opens with xarray / netCDF4 / ncdump like the reference's file;
        dimensions time / nface / nedge as in the reference's Dataset, time as seconds since the first stamp

`.zarr` raises with that explanation.  The boundary table goes to `<stem>_boundary_data.csv` as in the reference.
"""
from __future__ import annotations

import errno
import json
import os
from pathlib import Path
from typing import Any, Dict

import numpy as np


def _dims_of(name: str, a: np.ndarray, T: int, F: int, E: int):
    """Dimension names of a mesh variable from its shape (time, nface, nedge as in the reference's Dataset)."""
    table = {T: "time", F: "nface", E: "nedge"}
    if name == "time":
        return ("time",)
    dims = []
    for ax, sz in enumerate(a.shape):
        nm = table.get(sz)
        if nm is None or nm in dims or (nm == "time" and ax != 0):
            nm = f"{name}_dim{ax}"
        dims.append(nm)
    return tuple(dims)


class NpzWriter:
    """numpy archive (always available)."""
    def write(self, mesh: Dict[str, Any], output_file_path):
        attrs = {k: (v if isinstance(v, (int, float, str, bool)) else str(v)) for k, v in getattr(mesh, "attrs", {}).items()
                 if k != "boundary_data"}
        arrays = {k: np.asarray(v) for k, v in mesh.items()}
        if "time" in arrays and arrays["time"].dtype.kind == "M":
            arrays["time"] = arrays["time"].astype("datetime64[ns]").astype(np.int64)
            attrs["time_units"] = "nanoseconds since 1970-01-01"
        np.savez_compressed(output_file_path, attrs_json=np.array(json.dumps(attrs)), **arrays)


class NetCDF3Writer:
    """NetCDF-3 classic via scipy (no netCDF4 / HDF5 library needed)."""
    def write(self, mesh: Dict[str, Any], output_file_path):
        from scipy.io import netcdf_file
        T = len(mesh["time"])
        F = int(getattr(mesh, "attrs", {}).get("n_face", 0)) or int(np.asarray(mesh["volume"]).shape[-1])
        E = int(np.asarray(mesh["edges_face1"]).shape[0])
        with netcdf_file(str(output_file_path), "w", version=2) as nc:
            nc.createDimension("time", T); nc.createDimension("nface", F); nc.createDimension("nedge", E)
            for k, v in getattr(mesh, "attrs", {}).items():
                if isinstance(v, (int, float, str, np.integer, np.floating)):
                    setattr(nc, k, v)
            for name, arr in mesh.items():
                a = np.asarray(arr)
                units = None
                if name == "time":
                    if a.dtype.kind == "M":
                        t0 = a[0]
                        units = f"seconds since {np.datetime_as_string(t0, unit='s').replace('T', ' ')}"
                        a = (a - t0) / np.timedelta64(1, "s")
                    a = a.astype(np.float64)
                if a.dtype == np.bool_:
                    a = a.astype(np.int8)
                if a.dtype == np.int64:
                    a = a.astype(np.int32) if np.abs(a).max(initial=0) < 2**31 else a.astype(np.float64)
                dims = _dims_of(name, a, T, F, E)
                for d, sz in zip(dims, a.shape):
                    if d not in nc.dimensions:
                        nc.createDimension(d, sz)
                var = nc.createVariable(name, a.dtype, dims)
                var[...] = a
                if units:
                    var.units = units


class ClearWaterRiverineOutput:
    """Reference io/outputs.py:23-49: the output directory must exist."""
    def __init__(self, output_file_path: str, mesh) -> None:
        self.output_file_path = output_file_path
        if not Path(output_file_path).parents[0].is_dir():
            raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), output_file_path)
        self.mesh = mesh

    def write_mesh(self, writer) -> None:
        writer.write(self.mesh, self.output_file_path)


class ClearWaterRiverineOutputFactory:
    def get_writer(self, output_file_path):
        ext = Path(output_file_path).suffix
        if ext == ".npz":
            return NpzWriter()
        if ext == ".nc":
            return NetCDF3Writer()
        if ext == ".zarr":
            raise ValueError("Cannot save as .zarr here (zarr is not a dependency of this package): use .nc (NetCDF-3 classic, "
                             "opens with xarray / netCDF4) or .npz")
        raise ValueError(f"Cannot save as {ext}.")


writing_factory = ClearWaterRiverineOutputFactory()


def save_mesh(mesh, output_file_path) -> None:
    """Reference utilities.py `save_clearwater_xarray` + io/outputs.py ClearWaterRiverineWriter.write_mesh."""
    ClearWaterRiverineOutput(str(output_file_path), mesh).write_mesh(writing_factory.get_writer(str(output_file_path)))
