"""attach(model): run a LIVE reference `ClearwaterRiverine` object's update() on the GPU.

The reference has no plugin interface; its seam is the body of `ClearwaterRiverine.update()` (reference
transport.py:201-276) and the state it touches.  `attach` reads that state from the object as the reference built it --
the connectivity, the derived coefficients `advection_coeff` / `coeff_to_diffusion` (utilities.py:513-541: the very
arrays the reference's linalg would read, so both paths see identical inputs), `edge_velocity`, `volume`, `dt`, and every
constituent's `input_array` -- hands it to the CUDA library once, and replaces the bound `update` with one that

  * applies `update_concentration` overrides to row `time_step` of `mesh[name]` and to the device (transport.py:219-236),
  * takes the step on the device (assembly, right-hand sides, solve for all constituents, mass flux),
  * writes row `time_step + 1` of `mesh[name]` (real cells, boundary values re-imposed, other ghost cells left NaN:
    transport.py:252-264) and row `time_step` of the three `*_mass_flux` arrays of every constituent (267-273),
  * increments `time_step`.

Everything else of the object (plotting, saving, `finalize`, the mesh Dataset) keeps working on the same arrays.
`model.lhs` and `constituent.b` are NOT refreshed (the matrix lives on the device; `stepper.backend.get_lhs()` returns it).
Works on anything shaped like the reference object: `mesh[name].values` (or plain arrays), `mesh.nreal` / attrs,
`constituent_dict[name].input_array / advection_mass_flux / diffusion_mass_flux / total_mass_flux`, `time_step`.
"""
from __future__ import annotations

import types
import warnings
from typing import Any, Dict, Optional

import numpy as np

from .backend import CWR_OK, STATUS_NAMES, SolverWarning, TransportBackend


def _values(x) -> np.ndarray:
    return np.asarray(getattr(x, "values", x))


def _attr(mesh, name):
    attrs = getattr(mesh, "attrs", None)
    if attrs is not None and name in attrs:
        return attrs[name]
    return getattr(mesh, name)


def extract_model_arrays(model) -> Dict[str, Any]:
    """The arrays of the reference object the step reads, under plain names (no device needed)."""
    mesh = model.mesh
    f1 = np.ascontiguousarray(_values(mesh["edges_face1"]), dtype=np.int32)
    f2 = np.ascontiguousarray(_values(mesh["edges_face2"]), dtype=np.int32)
    vol = np.ascontiguousarray(_values(mesh["volume"]), dtype=np.float32)
    out = {
        "f1": f1, "f2": f2, "n_face": int(vol.shape[1]), "n_time": int(vol.shape[0]),
        "n_real": int(_attr(mesh, "nreal")) + 1,
        "diffusion_coefficient": float(_attr(mesh, "diffusion_coefficient")),
        "adv": np.ascontiguousarray(_values(mesh["advection_coeff"]), dtype=np.float32),
        "cdiff": np.ascontiguousarray(_values(mesh["coeff_to_diffusion"]), dtype=np.float64),
        "vel": np.ascontiguousarray(_values(mesh["edge_velocity"]), dtype=np.float32),
        "vol": vol,
        "dt": np.ascontiguousarray(_values(mesh["dt"]), dtype=np.float64),
        "constituents": list(model.constituent_dict.keys()),
        "inputs": {name: np.ascontiguousarray(c.input_array, dtype=np.float64) for name, c in model.constituent_dict.items()},
        "time_step": int(getattr(model, "time_step", 0)),
    }
    if out["n_real"] != int(f1.max()) + 1:
        raise ValueError("mesh.nreal does not match max(edges_face1) (reference io/hdf.py:268-269)")
    return out


class GpuStepper:
    """What `attach` leaves on the model as `model._cwr_b200`."""

    def __init__(self, model, device: int = 0, **options):
        a = extract_model_arrays(model)
        if a["time_step"] != 0:
            raise ValueError("attach() before the first update(): the device starts from the initial conditions")
        self.model = model
        self.names = a["constituents"]
        self.n = a["n_real"]
        hint = np.nanmean(a["adv"][:: max(1, a["n_time"] // 32)], axis=0, dtype=np.float64).astype(np.float32)
        self.backend = TransportBackend(a["f1"], a["f2"], a["n_face"], a["n_time"], len(self.names), a["diffusion_coefficient"],
                                        device=device, flow_hint=hint, **options)
        self.backend.set_hydro(0, a["adv"], a["cdiff"], a["vel"], a["vol"], a["dt"])
        for k, name in enumerate(self.names):
            self.backend.set_inputs(k, a["inputs"][name])
        self.index = {name: k for k, name in enumerate(self.names)}
        self.solver_info = []
        self._original_update = model.update

    def update(self, update_concentration: Optional[dict] = None):
        model, n = self.model, self.n
        t = model.time_step
        if isinstance(update_concentration, dict):
            for name in update_concentration:                       # transport.py:219-229
                if name not in model.constituent_dict:
                    print(f"WARNING: {name} is not being used in the model.")
                    print("Please review the constituent names in the update dictionary")
            for name, values in update_concentration.items():
                if name not in model.constituent_dict:
                    continue
                v = np.asarray(_values(values), dtype=np.float64)[0:n]
                _values(model.mesh[name])[t][0:n] = v                # transport.py:233-236
                self.backend.set_state(self.index[name], t, v)
        info = self.backend.step(t)
        self.solver_info.append((info.iterations, info.max_relres, info.status))
        if info.status != CWR_OK:
            warnings.warn(f"step {t}: {STATUS_NAMES.get(info.status, info.status)} "
                          f"({info.iterations} iterations, relres {info.max_relres:.3e})", SolverWarning)
        for name, k in self.index.items():
            con = model.constituent_dict[name]
            row = self.backend.get_state(k, t + 1)                   # real cells, BC ghost values, NaN elsewhere
            out = _values(model.mesh[name])
            out[t + 1, 0:n] = row[0:n]
            set_cols = np.nonzero(con.input_array[t + 1])[0]         # transport.py:258-264
            out[t + 1, set_cols] = con.input_array[t + 1][set_cols]
            if self.backend.options.mass_flux and getattr(con, "total_mass_flux", None) is not None:
                fa, fd, ft = self.backend.get_mass_flux(k, t)
                con.advection_mass_flux[t] = fa; con.diffusion_mass_flux[t] = fd; con.total_mass_flux[t] = ft
        model.time_step += 1                                          # transport.py:276

    def detach(self):
        self.model.update = self._original_update
        self.backend.close()
        if getattr(self.model, "_cwr_b200", None) is self:
            del self.model._cwr_b200


def attach(model, device: int = 0, **options) -> GpuStepper:
    """Replace `model.update` by the GPU step.  Returns the stepper (also kept as `model._cwr_b200`); `stepper.detach()`
    restores the reference's own update()."""
    stepper = GpuStepper(model, device=device, **options)
    model._cwr_b200 = stepper
    model.update = types.MethodType(lambda self, update_concentration=None: stepper.update(update_concentration), model)
    return stepper
