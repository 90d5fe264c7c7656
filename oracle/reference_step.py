"""CPU ORACLE -- test infrastructure only, never the product path.

A numpy/scipy restatement of ClearWater-Riverine's per-timestep implicit
advection-diffusion step.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
(clearwater_riverine_b200) never does and fails loudly without its CUDA library.

Parity pin: this restatement is checked bit-for-bit against the reference's own
linalg.LHS / linalg.RHS / ClearwaterRiverine.update / _mass_flux code, executed
in the build container through a minimal xarray stand-in
(oracle/refshim, tools/make_golden.py); the resulting vectors are committed
under tests/golden/ and tests/test_oracle_golden.py replays them.

Every function cites the reference lines it restates (paths relative to
/root/reference/src/clearwater_riverine/).  The sparse solve is the real thing:
scipy.sparse.csr_matrix + scipy.sparse.linalg.spsolve (SuperLU), exactly as
transport.py:215-218,249 call them.

dtype policy: adv, vel, vol are the float32 arrays
HEC-RAS stores; cdiff and dt are float64; V/dt is evaluated in float64
(numpy >= 2 promotion of `float32_array / np.float64`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.linalg import spsolve


# --------------------------------------------------------------------------
# Inputs produced upstream of the step (utilities.py:513-541)
# --------------------------------------------------------------------------
def derive_coefficients(face_flow, edge_velocity, face_x, face_y, f1, f2, diffusion_coefficient, time_seconds):
    """adv, area, dist, cdiff, dt exactly as WQVariableCalculator.calculate does.

    utilities.py:513-516  adv  = face_flow * sign(abs(vel))             (float32)
    utilities.py:518-522  area = (adv / vel).fillna(0)                  (float32)
    utilities.py:261-280  dist = sqrt(dx^2 + dy^2) of cell centres      (float64)
    utilities.py:304      cdiff = area * D / dist   -> f64(f32(area*D)) / dist;
                          ghost edges are NOT zeroed (the mask at 294-301 is dead)
    utilities.py:537-541  dt = diff(time) in seconds, last entry NaN
    """
    face_flow = np.asarray(face_flow, dtype=np.float32)
    edge_velocity = np.asarray(edge_velocity, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        adv = face_flow * np.sign(np.abs(edge_velocity))
        area = adv / edge_velocity
    area = np.where(np.isnan(area), np.float32(0), area).astype(np.float32)
    x1, y1 = face_x[f1], face_y[f1]
    x2, y2 = face_x[f2], face_y[f2]
    dist = np.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2)
    cdiff = area * float(diffusion_coefficient)  # python float is "weak": product stays float32
    assert cdiff.dtype == np.float32
    cdiff = cdiff / dist  # float32 / float64 -> float64
    dt = np.diff(np.asarray(time_seconds, dtype=np.float64))
    dt = np.append(dt, np.nan)
    return adv.astype(np.float32), area, dist, cdiff.astype(np.float64), dt


@dataclass
class HydroMesh:
    """Plain-array view of what the step reads from the reference's xarray mesh."""
    f1: np.ndarray            # (E,) int   edges_face1  (always a real cell)
    f2: np.ndarray            # (E,) int   edges_face2  (ghost iff > nreal)
    n_face: int               # F  real + ghost cells
    adv: np.ndarray           # (T,E) float32  advection_coeff
    cdiff: np.ndarray         # (T,E) float64  coeff_to_diffusion
    vel: np.ndarray           # (T,E) float32  edge_velocity
    vol: np.ndarray           # (T,F) float32  volume
    dt: np.ndarray            # (T,)  float64  dt  (last = NaN)
    diffusion_coefficient: float
    nreal: int = field(init=False)   # io/hdf.py:268-269  nreal = max(edges_face1)

    def __post_init__(self):
        self.f1 = np.asarray(self.f1, dtype=np.int64)
        self.f2 = np.asarray(self.f2, dtype=np.int64)
        self.nreal = int(self.f1.max())

    @property
    def n(self) -> int:       # linalg.py:32  nreal_count = nreal + 1
        return self.nreal + 1

    @property
    def n_edge(self) -> int:
        return len(self.f1)

    @property
    def n_time(self) -> int:
        return len(self.dt)


# --------------------------------------------------------------------------
# LHS   (linalg.py:17-156)
# --------------------------------------------------------------------------
class LHS:
    def __init__(self, mesh: HydroMesh):
        # linalg.py:28-32
        self.internal_edges = np.where((mesh.f1 <= mesh.nreal) & (mesh.f2 <= mesh.nreal))[0]
        self.internal_edge_count = len(self.internal_edges)
        self.real_edges_face1 = np.where(mesh.f1 <= mesh.nreal)[0]
        self.real_edges_face2 = np.where(mesh.f2 <= mesh.nreal)[0]
        self.nreal_count = mesh.nreal + 1
        self._is_internal = np.zeros(mesh.n_edge, dtype=bool)
        self._is_internal[self.internal_edges] = True

    def update_values(self, mesh: HydroMesh, t: int):
        """COO triplets in the reference's block order (linalg.py:59-156).

        The reference over-allocates its arrays (linalg.py:69-74);
        the surplus entries are (0, 0, 0.0) triplets, which add +0.0 to A[0,0] and
        are omitted here.
        """
        n = self.nreal_count
        adv_t = mesh.adv[t]
        cdiff_t = mesh.cdiff[t]
        # linalg.py:61-66
        flow_out = np.where(adv_t > 0)[0]
        flow_out_internal = np.where((adv_t > 0) & self._is_internal)[0]
        flow_in = np.where((adv_t < 0) & self._is_internal)[0]
        vol_next = mesh.vol[t + 1]
        empty_cells = np.where((vol_next == 0) & (np.arange(len(vol_next)) < n))[0][0:n]

        rows, cols, coef = [], [], []

        def block(r, c, v):
            rows.append(np.asarray(r, dtype=np.float64))
            cols.append(np.asarray(c, dtype=np.float64))
            coef.append(np.asarray(v, dtype=np.float64))

        # linalg.py:77-81  dry cells: dummy 1 on the diagonal
        block(empty_cells, empty_cells, np.ones(len(empty_cells)))
        # linalg.py:84-89  V[t+1]/dt[t] on the diagonal (float32 / float64 scalar -> float64)
        seconds = np.float64(mesh.dt[t])
        block(np.arange(n), np.arange(n), vol_next[0:n] / seconds)
        # linalg.py:92-97  cdiff on the diagonal of face1 for every edge (ghost edges included)
        e1 = self.real_edges_face1
        block(mesh.f1[e1], mesh.f1[e1], cdiff_t[e1])
        # linalg.py:99-103 cdiff on the diagonal of face2 for internal edges
        e2 = self.real_edges_face2
        block(mesh.f2[e2], mesh.f2[e2], cdiff_t[e2])
        # linalg.py:107-122 upwind advection, outflow from face1
        if len(flow_out) > 0:
            block(mesh.f1[flow_out], mesh.f1[flow_out], adv_t[flow_out])
            block(mesh.f2[flow_out_internal], mesh.f1[flow_out_internal], adv_t[flow_out_internal] * -1)
        # linalg.py:124-141 upwind advection, inflow into face1 (internal edges only)
        if len(flow_in) > 0:
            block(mesh.f1[flow_in], mesh.f2[flow_in], adv_t[flow_in])
            block(mesh.f2[flow_in], mesh.f2[flow_in], adv_t[flow_in] * -1)
        # linalg.py:145-156 diffusion off-diagonals
        ie = self.internal_edges
        block(mesh.f1[ie], mesh.f2[ie], -1 * cdiff_t[ie])
        block(mesh.f2[ie], mesh.f1[ie], -1 * cdiff_t[ie])

        self.rows = np.concatenate(rows)
        self.cols = np.concatenate(cols)
        self.coef = np.concatenate(coef)

    def to_csr(self) -> csr_matrix:
        # transport.py:215-218 (COO -> CSR, duplicates summed, float indices cast)
        n = self.nreal_count
        return csr_matrix((self.coef, (self.rows, self.cols)), shape=(n, n))


# --------------------------------------------------------------------------
# RHS   (linalg.py:158-406)
# --------------------------------------------------------------------------
class RHS:
    def __init__(self, mesh: HydroMesh, input_array: np.ndarray):
        # linalg.py:172-175
        self.nreal_count = mesh.nreal + 1
        self.input_array = input_array
        self.vals = np.zeros(self.nreal_count)
        self.ghost_cells = np.where(mesh.f2 > mesh.nreal)[0]

    def update_values(self, solution: np.ndarray, mesh: HydroMesh, t: int):
        # linalg.py:193-201
        solver = np.zeros(mesh.n_face)
        solver[0:self.nreal_count] = solution
        nz = self.input_array[t].nonzero()
        solver[nz] = self.input_array[t][nz]
        self.vals[:] = self._calculate_rhs(mesh, t, solver[0:self.nreal_count])

    def _calculate_load(self, mesh, t, concentrations):
        # linalg.py:227-241  float32 volume * float64 c / float64 dt
        return mesh.vol[t][0:self.nreal_count] * concentrations / np.float64(mesh.dt[t])

    def _calculate_rhs(self, mesh, t, concentrations):
        # linalg.py:262-275  ghost terms evaluated at t+1
        load = self._calculate_load(mesh, t, concentrations)
        n = self.nreal_count
        ghost_in = self._ghost_cell(mesh, t + 1, flowing_in=True)[0:n]
        ghost_out = self._ghost_cell(mesh, t + 1, flowing_in=False)[0:n]
        return load + ghost_in + ghost_out

    def _ghost_cell(self, mesh, t, flowing_in: bool):
        """linalg.py:354-406 (+ 278-352).

        Known divergence (SURVEY App. A.2 hazard): the reference pairs
        `nonzero(|coef|)` with cells positionally (linalg.py:349-351) and raises a
        shape-mismatch ValueError when a selected edge has a coefficient of exactly
        zero.  Here the (possibly zero) value is assigned per edge instead; the two
        agree whenever the reference does not raise.
        """
        advection = flowing_in                      # linalg.py:301-308
        cond = np.less if flowing_in else np.greater
        velocity_indices = np.where(cond(mesh.vel[t], 0))[0]            # 372
        index_list = np.intersect1d(velocity_indices, self.ghost_cells)  # 373
        internal = mesh.f1[index_list]                                  # 374
        external = mesh.f2[index_list]                                  # 375
        mult = np.zeros(mesh.n_face)
        mult[internal] = self.input_array[t][external]                  # 377-378 (last edge wins)
        adv_face = np.zeros(mesh.n_face)
        diff_face = np.zeros(mesh.n_face)
        if len(index_list) != 0:
            if advection:
                adv_face[internal] = np.abs(mesh.adv[t][index_list])    # 381-388 / 348-351
            if mesh.diffusion_coefficient != 0:                         # 390
                diff_face[internal] = np.abs(mesh.cdiff[t][index_list])  # 391-397
        add = adv_face + diff_face if flowing_in else diff_face         # 399-402
        return add * mult                                               # 404


# --------------------------------------------------------------------------
# Step driver   (transport.py:201-276, 406-429; constituents.py:19-76)
# --------------------------------------------------------------------------
class Constituent:
    """Array contract of constituents.py:19-76: input_array (T,F) holds the IC in
    row 0 and BC concentrations in ghost-cell columns; 0 means "not set"."""

    def __init__(self, name: str, mesh: HydroMesh, input_array: np.ndarray,
                 initial_row: Optional[np.ndarray] = None):
        T, F, E = mesh.n_time, mesh.n_face, mesh.n_edge
        self.name = name
        self.input_array = np.asarray(input_array, dtype=np.float64)
        assert self.input_array.shape == (T, F)
        self.advection_mass_flux = np.zeros((T, E))   # constituents.py:28-30
        self.diffusion_mass_flux = np.zeros((T, E))
        self.total_mass_flux = np.zeros((T, E))
        self.concentration = np.full((T, F), np.nan)  # constituents.py:39-48
        # constituents.py:94-98: row 0 is written by set_initial_conditions BEFORE the BCs are
        # merged into input_array (constituents.py:51-59), so it holds the IC cells and zeros
        # elsewhere -- not the ghost-cell BC values, not NaN.  Default: the real-cell part of
        # input_array[0] (what every fixture's IC CSV lists).
        if initial_row is None:
            initial_row = np.zeros(F)
            initial_row[0:mesh.n] = self.input_array[0][0:mesh.n]
        self.concentration[0] = initial_row
        self.b = RHS(mesh, self.input_array)


class OracleRiverine:
    """Array-level mirror of ClearwaterRiverine (transport.py:68) for the hot path."""

    def __init__(self, mesh: HydroMesh, inputs: Dict[str, np.ndarray]):
        self.mesh = mesh
        self.lhs = LHS(mesh)
        self.constituent_dict = {k: Constituent(k, mesh, v) for k, v in inputs.items()}
        self.time_step = 0
        self.last_A: Optional[csr_matrix] = None

    def update(self, update_concentration: Optional[Dict[str, np.ndarray]] = None):
        mesh, t, n = self.mesh, self.time_step, self.mesh.n
        self.lhs.update_values(mesh, t)                      # transport.py:208-211
        A = self.lhs.to_csr()                                # 215-218
        self.last_A = A
        for name, c in self.constituent_dict.items():        # 231
            if isinstance(update_concentration, dict) and name in update_concentration:
                c.concentration[t][0:n] = np.asarray(update_concentration[name])[0:n]   # 233-236
                x = np.asarray(update_concentration[name])[0:n]
            else:
                x = c.concentration[t][0:n]                  # 238
            c.b.update_values(x, mesh, t)                    # 241-246
            x = spsolve(A, c.b.vals)                         # 249
            c.concentration[t + 1][0:n] = x                  # 252-257
            nz = np.nonzero(c.input_array[t + 1])[0]         # 258
            c.concentration[t + 1][nz] = c.input_array[t + 1][nz]   # 259-264
            mass_flux(mesh, c.concentration, c.advection_mass_flux,
                      c.diffusion_mass_flux, c.total_mass_flux, t)  # 267-273
        self.time_step += 1                                  # 276


def mass_flux(mesh: HydroMesh, output, advection_mass_flux, diffusion_mass_flux, total_mass_flux, t):
    """transport.py:406-429 (NaN ghost concentrations propagate, as in the reference)."""
    adv_t = mesh.adv[t]
    negative = adv_t < 0
    parent = output[t + 1][mesh.f1]
    neighbor = output[t + 1][mesh.f2]
    dt = np.float64(mesh.dt[t])
    advection_mass_flux[t] = np.where(negative, adv_t * neighbor, adv_t * parent) * dt
    diffusion_mass_flux[t] = mesh.cdiff[t] * (neighbor - parent) * dt
    total_mass_flux[t] = advection_mass_flux[t] + diffusion_mass_flux[t]


# --------------------------------------------------------------------------
# Mass balance   (postproc_util.py:21-166)  -- acceptance metric of test_final_mass.py
# --------------------------------------------------------------------------
def mass_balance(mesh: HydroMesh, concentration, total_mass_flux, face_flow, bc_faces: Dict[str, np.ndarray]):
    """Start/end mass over real cells and per-boundary volume/mass in (<=0) / out (>=0).

    postproc_util.py:36-59 (mass start/end), 86-143 (per BC line sums),
    153-165 (closure).  `face_flow` is the raw (T,E) float32 HEC-RAS Face Flow.
    """
    n = mesh.n
    T = mesh.n_time
    out = {}
    vol0, vol1 = mesh.vol[0][0:n], mesh.vol[T - 1][0:n]
    out["Vol_start"] = vol0.sum()
    out["Mass_start"] = (vol0 * concentration[0][0:n]).sum()
    out["Vol_end"] = vol1.sum()
    out["Mass_end"] = (vol1 * concentration[T - 1][0:n]).sum()
    tot_v_in = tot_v_out = tot_m_in = tot_m_out = 0.0
    for name, faces in bc_faces.items():
        faces = np.asarray(faces, dtype=np.int64)
        edge_vol = face_flow[:, faces] * mesh.dt[:, None]          # NaN in the last row ...
        edge_mass = total_mass_flux[:, faces]
        out[f"{name}_vol"] = np.nansum(edge_vol)                    # ... skipped by xarray's sum(skipna)
        out[f"{name}_mass"] = edge_mass.sum()
        v_in = np.nansum(np.where(edge_vol <= 0, edge_vol, 0.0))
        v_out = np.nansum(np.where(edge_vol >= 0, edge_vol, 0.0))
        m_in = np.where(edge_mass <= 0, edge_mass, edge_mass * 0).sum()
        m_out = np.where(edge_mass >= 0, edge_mass, edge_mass * 0).sum()
        out[f"{name}_in_vol"], out[f"{name}_out_vol"] = v_in, v_out
        out[f"{name}_in_mass"], out[f"{name}_out_mass"] = m_in, m_out
        tot_v_in += v_in; tot_v_out += v_out; tot_m_in += m_in; tot_m_out += m_out
    out["bcTotalVolInAll"], out["bcTotalVolOutAll"] = tot_v_in, tot_v_out
    out["bcTotalMassInAll"], out["bcTotalMassOutAll"] = tot_m_in, tot_m_out
    out["vol_end_calc"] = out["Vol_start"] - tot_v_in - tot_v_out
    out["mass_end_calc"] = out["Mass_start"] - tot_m_in - tot_m_out
    out["error_vol"] = out["vol_end_calc"] - out["Vol_end"]
    out["error_mass"] = out["mass_end_calc"] - out["Mass_end"]
    return out
