"""Output files of the driver (CSR.py:784-879): wakes per CSR step, run statistics, particle dumps.

The reference writes HDF5 through ``h5py`` (and openPMD particle files through ``pmd_beamphysics``).  ``CSR2D`` here
writes through the SAME calls -- ``create_group`` / ``create_dataset`` / ``attrs[...]`` with the reference's group,
dataset and attribute names -- on a store opened by ``open_store``:

* with ``h5py`` importable the store IS an ``h5py.File`` and the files are the reference's ``*.h5`` files (readable by
  its ``postprocessor.py``);
* offline (this image has no HDF5 library at all) the store is an ``NpzStore``: the same tree kept in memory and saved
  as ``<name>.npz`` with ``/``-joined paths as keys and ``path@attr`` keys for attributes.  ``load_store`` reads either
  kind back into one nested dictionary, so downstream code is independent of the container.
"""
from __future__ import annotations

import os

import numpy as np

try:
    import h5py as _h5py
except Exception:  # not installed offline
    _h5py = None

HAVE_H5PY = _h5py is not None


class _Node:
    def __init__(self, store, path):
        self._store, self._path = store, path
        self.attrs = _Attrs(store, path)

    def _join(self, name):
        return f"{self._path}/{name}" if self._path else name

    def create_group(self, name):
        path = self._join(name)
        if path in self._store._groups:
            raise ValueError(f"Unable to create group (name already exists): {path}")
        self._store._groups.add(path)
        return _Node(self._store, path)

    def create_dataset(self, name, data=None, shape=None):
        path = self._join(name)
        if path in self._store._data:
            raise ValueError(f"Unable to create dataset (name already exists): {path}")
        arr = np.asarray(data)
        if shape is not None and tuple(shape) != arr.shape:
            arr = arr.reshape(shape)
        self._store._data[path] = arr
        return arr

    def __getitem__(self, name):
        path = self._join(name)
        if path in self._store._data:
            return self._store._data[path]
        if path in self._store._groups:
            return _Node(self._store, path)
        raise KeyError(path)

    def keys(self):
        pre = self._path + "/" if self._path else ""
        names = {p[len(pre):].split("/")[0] for p in list(self._store._data) + list(self._store._groups) if p.startswith(pre)}
        return sorted(n for n in names if n)


class _Attrs:
    def __init__(self, store, path):
        self._store, self._path = store, path

    def __setitem__(self, key, value):
        self._store._attrs[f"{self._path}@{key}"] = np.asarray(value)

    def __getitem__(self, key):
        v = self._store._attrs[f"{self._path}@{key}"]
        return v.item() if v.shape == () else v


class NpzStore(_Node):
    """h5py.File look-alike for the subset the driver uses, persisted as one .npz file."""

    def __init__(self, filename, mode="a"):
        self.filename = filename
        self._data, self._attrs, self._groups = {}, {}, set()
        if mode in ("a", "r", "r+") and os.path.isfile(filename):
            with np.load(filename, allow_pickle=False) as z:
                for k in z.files:
                    if k == "__groups__":
                        self._groups = set(str(g) for g in z[k])
                    elif "@" in k:
                        self._attrs[k] = z[k]
                    else:
                        self._data[k] = z[k]
        self._mode = mode
        _Node.__init__(self, self, "")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._mode != "r":
            payload = dict(self._data)
            payload.update(self._attrs)
            payload["__groups__"] = np.array(sorted(self._groups), dtype=str)
            tmp = self.filename + ".tmp.npz"
            np.savez(tmp, **payload)
            os.replace(tmp, self.filename)


def store_path(basename: str) -> str:
    """File name for a store called `basename` (no extension): .h5 with h5py, .npz without."""
    return basename + (".h5" if HAVE_H5PY else ".npz")


def open_store(filename: str, mode: str = "a"):
    """h5py.File(filename, mode) when h5py is importable and the name ends in .h5, else an NpzStore."""
    if HAVE_H5PY and filename.endswith(".h5"):
        return _h5py.File(filename, mode)
    return NpzStore(filename, mode)


def load_store(filename: str) -> dict:
    """Nested dict {name: array | dict, "@attrs": {...}} of a file written through open_store (either container)."""
    def walk(node):
        out = {}
        attrs = {}
        if HAVE_H5PY and filename.endswith(".h5"):
            attrs = {k: node.attrs[k] for k in node.attrs}
            for k in node.keys():
                out[k] = walk(node[k]) if isinstance(node[k], _h5py.Group) else np.asarray(node[k])
        else:
            pre = node._path + "@"
            attrs = {k[len(pre):]: (v.item() if v.shape == () else v) for k, v in node._store._attrs.items()
                     if k.startswith(pre) and "/" not in k[len(pre):]}
            for k in node.keys():
                child = node[k]
                out[k] = walk(child) if isinstance(child, _Node) else child
        if attrs:
            out["@attrs"] = attrs
        return out
    with open_store(filename, "r") as st:
        return walk(st)


def dict2hdf5(hf, dic, group=None):
    """tools.py:69-77 of the reference: nested dict -> groups / datasets."""
    for key, item in dic.items():
        if not isinstance(item, dict):
            (group if group is not None else hf).create_dataset(key, data=item)
        else:
            dict2hdf5(hf, item, hf.create_group(key))
