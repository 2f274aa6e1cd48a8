"""Name -> class registries, the plugin seam of the reference (fdbm/util/registry.py:5-34,
fdbm/backbones/shared.py:22, fdbm/bridge.py:11): `BackboneRegistry.get_by_name("ncsnpp_v2")(**kw)`."""
import warnings
from typing import Callable


class Registry:
    def __init__(self, managed_thing: str):
        self.managed_thing = managed_thing
        self._registry = {}

    def register(self, name: str) -> Callable:
        def wrap(cls):
            if name in self._registry:
                warnings.warn(f"{self.managed_thing} with name '{name}' doubly registered, old class will be replaced.")
            self._registry[name] = cls
            return cls
        return wrap

    def get_by_name(self, name: str):
        if name not in self._registry:
            raise ValueError(f"{self.managed_thing} with name '{name}' unknown.")
        return self._registry[name]

    def get_all_names(self):
        return list(self._registry.keys())


BackboneRegistry = Registry("Backbone")
BridgeRegistry = Registry("Bridge")
