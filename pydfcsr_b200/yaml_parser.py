"""YAML -> ordered dict, same contract as the reference's parse_yaml (yaml_parser.py:30-50):
accepts a path or an open stream; mapping order is preserved (the lattice is an ordered list of
named elements, lattice.py:23-24)."""
from __future__ import annotations

import os
from collections import OrderedDict

import yaml


def full_path(path: str) -> str:
    return os.path.abspath(os.path.expandvars(path))


class _OrderedLoader(yaml.SafeLoader):
    pass


def _mapping(loader, node):
    loader.flatten_mapping(node)
    return OrderedDict(loader.construct_pairs(node))


_OrderedLoader.add_constructor(yaml.resolver.BaseResolver.DEFAULT_MAPPING_TAG, _mapping)


def parse_yaml(source):
    if isinstance(source, dict):
        return source
    if isinstance(source, str):
        path = full_path(source)
        if not os.path.exists(path):
            raise FileNotFoundError(f"input file does not exist: {source}")
        with open(path) as fh:
            return yaml.load(fh, _OrderedLoader)
    out = yaml.load(source, _OrderedLoader)
    if not isinstance(out, dict):
        raise ValueError(f"could not parse {source!r} as a YAML mapping")
    return out
