"""TEST INFRASTRUCTURE -- a tiny stand-in for the slice of the xarray API that the
reference's hot path touches, so that the UNMODIFIED reference sources under
/root/reference can be executed in the build container (xarray/h5py are not
installed and cannot be).  Used only by tools/make_golden.py to generate the
golden vectors in tests/golden/; nothing in the product imports it.

Design: DataArray is an ndarray subclass, so every arithmetic, comparison,
ufunc, indexing and dtype-promotion rule is numpy's own -- which is what xarray
delegates to for same-dimension operands (all the reference uses).  Only the
label/metadata conveniences (dims, attrs, .values, .isel, .loc, .fillna,
Dataset lookup, accessors) are re-implemented.
"""
from __future__ import annotations

import copy as _copy

import numpy as np

__version__ = "standin"


def _as_dims(dims, ndim):
    if dims is None:
        return tuple(f"dim_{i}" for i in range(ndim))
    if isinstance(dims, str):
        return (dims,)
    return tuple(dims)


class _Loc:
    def __init__(self, da):
        self._da = da

    def _index(self, key):
        da = self._da
        idx = [slice(None)] * da.ndim
        for dim, label in key.items():
            ax = da.dims.index(dim)
            coord = da._coords.get(dim)
            lab = np.asarray(label)
            if coord is None or lab.dtype.kind in "iu":
                idx[ax] = lab if lab.ndim else int(lab)       # positional == label for range indexes
            else:
                cv = np.asarray(coord)
                if lab.ndim == 0:
                    pos = np.nonzero(cv == lab)[0]
                    if len(pos) == 0:
                        raise KeyError(label)
                    idx[ax] = int(pos[0])
                else:
                    idx[ax] = np.array([int(np.nonzero(cv == v)[0][0]) for v in lab])
        return tuple(idx)

    def __setitem__(self, key, value):
        np.asarray(self._da)[self._index(key)] = np.asarray(value)

    def __getitem__(self, key):
        return self._da[self._index(key)]


class DataArray(np.ndarray):
    def __new__(cls, data=None, coords=None, dims=None, attrs=None, name=None):
        if hasattr(data, "to_numpy"):
            data = data.to_numpy()
        arr = np.asarray(data)
        obj = arr.view(cls)
        obj.dims = _as_dims(dims, arr.ndim)
        obj.attrs = dict(attrs) if attrs else {}
        obj._coords = {}
        if coords:
            for k, v in dict(coords).items():
                if isinstance(v, tuple):
                    obj._coords[k] = DataArray(v[1], dims=v[0])
                else:
                    obj._coords[k] = DataArray(v, dims=(k,))
        return obj

    def __array_finalize__(self, obj):
        self.dims = getattr(obj, "dims", ())
        self.attrs = getattr(obj, "attrs", {})
        self._coords = getattr(obj, "_coords", {})

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        attrs = self.__dict__.get("attrs", {})
        if item in attrs:
            return attrs[item]
        raise AttributeError(item)

    @property
    def values(self):
        return self.view(np.ndarray)

    @property
    def data(self):
        return self.view(np.ndarray)

    @property
    def loc(self):
        return _Loc(self)

    def isel(self, **kw):
        idx = [slice(None)] * self.ndim
        for dim, i in kw.items():
            idx[self.dims.index(dim)] = i
        return self[tuple(idx)]

    def sel(self, **kw):
        return self.isel(**kw)

    def fillna(self, value):
        # modern xarray: python scalars promote weakly (float32 stays float32)
        out = np.where(np.isnan(self.values), np.asarray(value, dtype=self.dtype), self.values)
        return DataArray(out, dims=self.dims, attrs=self.attrs)

    def where(self, cond, other=np.nan):
        return DataArray(np.where(np.asarray(cond), self.values, other), dims=self.dims)

    def sum(self, *a, **k):
        return DataArray(np.nansum(self.values, *a, **k))

    def to_numpy(self):
        return self.values

    def to_list(self):
        return self.values.tolist()


def where(cond, x, y):
    return DataArray(np.where(np.asarray(cond), np.asarray(x), np.asarray(y)))


_ACCESSORS = {}


def register_dataset_accessor(name):
    def deco(cls):
        _ACCESSORS[name] = cls
        return cls
    return deco


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        object.__setattr__(self, "_vars", {})
        object.__setattr__(self, "_coords", {})
        object.__setattr__(self, "_sizes", {})
        object.__setattr__(self, "attrs", dict(attrs) if attrs else {})
        object.__setattr__(self, "_accessor_cache", {})

    # -- mapping ----------------------------------------------------------------
    def _note(self, da):
        for d, s in zip(da.dims, da.shape):
            self._sizes[d] = s

    def __setitem__(self, key, value):
        if not isinstance(value, DataArray):
            value = DataArray(value)
        self._vars[key] = value
        self._note(value)
        for k, c in value._coords.items():
            self._coords[k] = c
            self._note(c)

    def __getitem__(self, key):
        if key in self._vars:
            da = self._vars[key]
            for d in da.dims:          # dataset-level coordinates index its variables
                if d in self._coords and d not in da._coords:
                    da._coords = {**da._coords, d: self._coords[d]}
            return da
        if key in self._coords:
            return self._coords[key]
        if key in self._sizes:     # dimension without coordinate -> range index
            return DataArray(np.arange(self._sizes[key]), dims=(key,))
        raise KeyError(key)

    def __contains__(self, key):
        return key in self._vars or key in self._coords

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        if item in _ACCESSORS:
            cache = self._accessor_cache
            if item not in cache:
                cache[item] = _ACCESSORS[item](self)
            return cache[item]
        try:
            return self[item]
        except KeyError:
            pass
        if item in self.attrs:
            return self.attrs[item]
        raise AttributeError(item)

    def __setattr__(self, key, value):
        if key == "attrs":
            object.__setattr__(self, "attrs", dict(value))
        else:
            raise AttributeError(f"cannot set {key} on Dataset stand-in")

    def assign_coords(self, coords=None, **kw):
        new = Dataset()
        new._vars.update(self._vars)
        new._coords.update(self._coords)
        new._sizes.update(self._sizes)
        object.__setattr__(new, "attrs", self.attrs)      # xarray keeps attrs on assign_coords
        for k, v in {**(coords or {}), **kw}.items():
            if not isinstance(v, DataArray):
                v = DataArray(v, dims=(k,))
            new._coords[k] = v
            new._note(v)
        return new

    def copy(self, deep=False):
        return _copy.deepcopy(self) if deep else self.assign_coords()

    @property
    def data_vars(self):
        return self._vars

    def keys(self):
        return list(self._vars.keys())


def open_zarr(*a, **k):
    raise NotImplementedError("xarray stand-in: no I/O")


open_dataset = open_zarr
