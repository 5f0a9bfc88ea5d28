import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields


class BaseOutput(OrderedDict):
    """Dataclass-style output container (attribute and key access), like diffusers.utils.BaseOutput."""

    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()
