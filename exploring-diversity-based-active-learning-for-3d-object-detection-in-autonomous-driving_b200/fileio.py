"""``load`` / ``dump`` by file extension with the reference's formats (det3d/torchie/fileio/io.py:15-82).

json via ``json`` (handlers/json_handler.py:7-14), pkl/pickle via ``pickle`` with dump protocol 2
(handlers/pickle_handler.py:13-19), yaml/yml via PyYAML.  Keyword arguments are forwarded, so
``dump(buffer, path, indent=4)`` writes the indented initial buffer of tools/active_select.py:70.
"""
import json
import pickle
from pathlib import Path

_TEXT = {"json", "yaml", "yml"}
_KNOWN = _TEXT | {"pkl", "pickle"}


def _format_of(file, file_format):
    if file_format is None and isinstance(file, str):
        file_format = file.split(".")[-1]
    if file_format not in _KNOWN:
        raise TypeError(f"Unsupported format: {file_format}")
    return file_format


def _load_obj(f, fmt, **kwargs):
    if fmt == "json":
        return json.load(f)
    if fmt in ("yaml", "yml"):
        import yaml
        kwargs.setdefault("Loader", yaml.SafeLoader)
        return yaml.load(f, **kwargs)
    return pickle.load(f, **kwargs)


def _dump_obj(obj, f, fmt, **kwargs):
    if fmt == "json":
        json.dump(obj, f, **kwargs)
    elif fmt in ("yaml", "yml"):
        import yaml
        yaml.safe_dump(obj, f, **kwargs)
    else:
        kwargs.setdefault("protocol", 2)
        pickle.dump(obj, f, **kwargs)


def load(file, file_format=None, **kwargs):
    if isinstance(file, Path):
        file = str(file)
    fmt = _format_of(file, file_format)
    if isinstance(file, str):
        with open(file, "r" if fmt in _TEXT else "rb") as f:
            return _load_obj(f, fmt, **kwargs)
    if hasattr(file, "read"):
        return _load_obj(file, fmt, **kwargs)
    raise TypeError('"file" must be a filepath str or a file-object')


def dump(obj, file=None, file_format=None, **kwargs):
    if isinstance(file, Path):
        file = str(file)
    if file is None and file_format is None:
        raise ValueError("file_format must be specified since file is None")
    fmt = _format_of(file, file_format)
    if file is None:
        if fmt == "json":
            return json.dumps(obj, **kwargs)
        if fmt in ("yaml", "yml"):
            import yaml
            return yaml.safe_dump(obj, **kwargs)
        kwargs.setdefault("protocol", 2)
        return pickle.dumps(obj, **kwargs)
    if isinstance(file, str):
        with open(file, "w" if fmt in _TEXT else "wb") as f:
            _dump_obj(obj, f, fmt, **kwargs)
    elif hasattr(file, "write"):
        _dump_obj(obj, file, fmt, **kwargs)
    else:
        raise TypeError('"file" must be a filename str or a file-object')
