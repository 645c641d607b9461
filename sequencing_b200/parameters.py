"""Serializable parameter trees (host-side configuration plumbing; no arithmetic).

Mirrors the public surface of the reference's ``sequencing/parameters.py:31-413``
(``Parameterized`` with recursive get/set, ``temporarily_set``, dict/JSON round trips, and the
typed ``*Parameter`` attr factories) so that ``Mode``/``Pulse``/``System`` objects written against
the reference keep working.  Implementation is attrs-based like the reference so that user
subclasses decorated with ``@attr.s`` behave identically.
"""
import importlib
import json
import logging
from contextlib import contextmanager

import attr
import numpy as np


class NumpyJSONEncoder(json.JSONEncoder):
    """JSON encoder that understands numpy scalars and arrays."""

    def default(self, obj):
        if isinstance(obj, np.generic):
            return obj.item()
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


def _revive(value):
    """Rebuild nested Parameterized objects from their ``as_dict`` form."""
    if isinstance(value, dict):
        if "cls" in value:
            module_name, _, cls_name = value["cls"].rpartition(".")
            klass = getattr(importlib.import_module(module_name), cls_name)
            return klass.from_dict(value)
        return {k: _revive(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_revive(v) for v in value]
    return value


@attr.s
class Parameterized(object):
    """A named object whose attrs fields are its (possibly nested) parameters."""

    name = attr.ib(type=str)
    cls = attr.ib(type=str, default="")
    logger = logging.getLogger("Parameterized")

    def initialize(self):
        """Post-init hook; subclasses extend it and call ``super().initialize()``."""

    def __attrs_post_init__(self):
        self.cls = f"{type(self).__module__}.{type(self).__name__}"
        self.initialize()

    # -- recursive access ----------------------------------------------------------------
    def get_param(self, address, *default, delimiter="."):
        obj = self
        for part in address.split(delimiter):
            obj = getattr(obj, part, *default)
        return obj

    def set_param(self, address, value, delimiter="."):
        head, _, leaf = address.rpartition(delimiter)
        target = self.get_param(head, delimiter=delimiter) if head else self
        setattr(target, leaf, value)

    def get(self, *addresses, delimiter="."):
        return {a: self.get_param(a, delimiter=delimiter) for a in addresses}

    def set(self, **kwargs):
        delimiter = kwargs.pop("delimiter", "__")
        for address, value in kwargs.items():
            self.set_param(address, value, delimiter=delimiter)

    @contextmanager
    def temporarily_set(self, **kwargs):
        delimiter = kwargs.pop("delimiter", "__")
        saved = self.get(*kwargs, delimiter=delimiter)
        try:
            self.set(delimiter=delimiter, **kwargs)
            yield
        finally:
            self.set(delimiter=delimiter, **saved)

    # -- (de)serialisation ---------------------------------------------------------------
    def as_dict(self, json_friendly=True):
        return attr.asdict(self, retain_collection_types=True)

    def to_json(self, dumps=False, json_path=None):
        d = self.as_dict(json_friendly=True)
        if dumps:
            if json_path is not None:
                raise ValueError("If dumps is True, json_path must be None.")
            return json.dumps(d, indent=2, sort_keys=True, cls=NumpyJSONEncoder)
        path = json_path or f"{self.name}.json"
        if not path.endswith(".json"):
            path += ".json"
        with open(path, "w") as f:
            json.dump(d, f, indent=2, sort_keys=True, cls=NumpyJSONEncoder)

    @classmethod
    def from_dict(cls, d):
        known = attr.fields_dict(cls)
        return cls(**{k: _revive(v) for k, v in d.items() if k in known})

    @classmethod
    def from_json(cls, json_path=None, json_str=None):
        if (json_path is None) == (json_str is None):
            raise ValueError("Provide exactly one of json_path or json_str.")
        if json_str is not None:
            return cls.from_dict(json.loads(json_str))
        if not json_path.endswith(".json"):
            json_path += ".json"
        with open(json_path, "r") as f:
            return cls.from_dict(json.load(f))


def _typed(converter):
    def factory(default, unit=None, **kwargs):
        if unit is not None:
            kwargs["metadata"] = {"unit": str(unit)}
        return attr.ib(default=default, converter=converter, **kwargs)

    return factory


IntParameter = _typed(int)
FloatParameter = _typed(float)


def StringParameter(default, **kwargs):
    return attr.ib(default=default, converter=str, **kwargs)


def BoolParameter(default, **kwargs):
    return attr.ib(default=default, converter=bool, **kwargs)


def NanosecondParameter(default, base=IntParameter, **kwargs):
    return base(default, unit="ns", **kwargs)


def GigahertzParameter(default, base=FloatParameter, **kwargs):
    return base(default, unit="GHz", **kwargs)


def RadianParameter(default, base=FloatParameter, **kwargs):
    return base(default, unit="radian", **kwargs)


def _container(kind):
    def factory(default=None, factory=kind, **kwargs):
        if default is not None:
            return attr.ib(default, **kwargs)
        return attr.ib(factory=factory, **kwargs)

    return factory


DictParameter = _container(dict)
ListParameter = _container(list)
