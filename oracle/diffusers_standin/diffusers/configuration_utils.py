import functools
import inspect
from types import SimpleNamespace


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return getattr(self, "_standin_config", SimpleNamespace())


def register_to_config(init):
    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        self._standin_config = SimpleNamespace(**cfg)
        init(self, *args, **kwargs)

    return wrapper
