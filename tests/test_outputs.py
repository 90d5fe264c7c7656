"""finalize(save=True) writers (next row N4; reference io/outputs.py:10-89, transport.py:385-395): host only."""
import json

import numpy as np
import pytest

from clearwater_riverine_b200.io.outputs import save_mesh
from clearwater_riverine_b200.transport import ModelMesh


def _mesh():
    T, F, E = 5, 8, 7
    rng = np.random.default_rng(0)
    m = ModelMesh()
    m.attrs.update({"diffusion_coefficient": 0.1, "nreal": 1, "n_face": F})
    m["edges_face1"] = np.arange(E, dtype=np.int32) % 2
    m["edges_face2"] = np.arange(E, dtype=np.int32) + 1
    m["face_x"], m["face_y"] = rng.random(F), rng.random(F)
    m["time"] = np.datetime64("2023-01-01T12:00:00") + np.arange(T) * np.timedelta64(300, "s")
    m["face_flow"] = rng.random((T, E)).astype(np.float32)
    m["volume"] = rng.random((T, F)).astype(np.float32)
    m["dt"] = np.append(np.full(T - 1, 300.0), np.nan)
    c = np.full((T, F), np.nan); c[:, :2] = rng.random((T, 2))
    m["tracer"] = c
    return m


def test_npz_round_trip(tmp_path):
    m = _mesh()
    save_mesh(m, tmp_path / "out.npz")
    z = np.load(tmp_path / "out.npz")
    assert np.array_equal(z["tracer"], m["tracer"], equal_nan=True) and z["face_flow"].dtype == np.float32
    attrs = json.loads(str(z["attrs_json"]))
    assert attrs["diffusion_coefficient"] == 0.1 and attrs["nreal"] == 1
    assert np.array_equal(z["time"].astype("datetime64[ns]"), m["time"].astype("datetime64[ns]"))


def test_netcdf3_round_trip(tmp_path):
    from scipy.io import netcdf_file
    m = _mesh()
    save_mesh(m, tmp_path / "out.nc")
    with netcdf_file(str(tmp_path / "out.nc"), "r", mmap=False) as nc:
        assert nc.dimensions["time"] == 5 and nc.dimensions["nface"] == 8 and nc.dimensions["nedge"] == 7
        assert nc.variables["tracer"].dimensions == ("time", "nface")
        assert nc.variables["face_flow"].dimensions == ("time", "nedge")
        assert np.array_equal(nc.variables["tracer"][:], m["tracer"], equal_nan=True)
        assert np.array_equal(nc.variables["time"][:], np.arange(5) * 300.0)
        assert nc.variables["time"].units.decode().startswith("seconds since 2023-01-01 12:00:00")
        assert abs(nc.diffusion_coefficient - 0.1) < 1e-15


def test_unknown_extensions_and_missing_directory(tmp_path):
    m = _mesh()
    with pytest.raises(ValueError):
        save_mesh(m, tmp_path / "out.zarr")
    with pytest.raises(ValueError):
        save_mesh(m, tmp_path / "out.xyz")
    with pytest.raises(FileNotFoundError):
        save_mesh(m, tmp_path / "nope" / "out.npz")
