"""Name -> class registry with the reference's semantics (det3d/utils/registry.py:6-80)."""
import inspect


class Registry:
    """``@REG.register_module`` adds a class under its ``__name__``; duplicates raise KeyError (:28-41)."""

    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __repr__(self):
        return f"{type(self).__name__}(name={self._name}, items={list(self._module_dict)})"

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError(f"module must be a class, but got {type(cls)}")
        if cls.__name__ in self._module_dict:
            raise KeyError(f"{cls.__name__} is already registered in {self._name}")
        self._module_dict[cls.__name__] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    """Construct ``registry[cfg["type"]](**rest_of_cfg)`` (det3d/utils/registry.py:49-80).

    ``type`` may be a registered name (KeyError when unknown) or a class (TypeError for anything
    else); ``default_args`` only fill keys the cfg lacks; the cfg itself is not modified.
    """
    assert isinstance(cfg, dict) and "type" in cfg
    assert isinstance(default_args, dict) or default_args is None
    kwargs = dict(cfg)
    obj_type = kwargs.pop("type")
    if isinstance(obj_type, str):
        cls = registry.get(obj_type)
        if cls is None:
            raise KeyError(f"{obj_type} is not in the {registry.name} registry")
    elif inspect.isclass(obj_type):
        cls = obj_type
    else:
        raise TypeError(f"type must be a str or valid type, but got {type(obj_type)}")
    for key, value in (default_args or {}).items():
        kwargs.setdefault(key, value)
    return cls(**kwargs)
