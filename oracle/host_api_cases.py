"""oracle/host_api_cases.py -- TEST INFRASTRUCTURE ONLY.

One scenario table for the host-side API the selection path goes through before and after the kernels:
``Registry`` / ``build_from_cfg`` (det3d/utils/registry.py:6-80) and ``load`` / ``dump``
(det3d/torchie/fileio/io.py:15-82 with its json / pickle / yaml handlers).  ``run_all(api, workdir)`` executes every
scenario against an ``api`` namespace -- the reference's modules in oracle/make_golden_host_api.py, this package's
in tests/test_host.py -- and returns JSON-able outcomes: the value, or the exception type and ``args``.
"""
import hashlib
import io
import os
from pathlib import Path

import numpy as np


def _outcome(fn):
    try:
        return {"ok": fn()}
    except Exception as e:   # noqa: BLE001 - the exception type and message ARE the outcome
        return {"raises": type(e).__name__, "args": [str(a) for a in e.args]}


def _sha(data):
    return hashlib.sha256(data if isinstance(data, bytes) else data.encode()).hexdigest()


def registry_cases(api):
    Registry, build = api["Registry"], api["build_from_cfg"]
    reg = Registry("selector")

    class Alpha:
        def __init__(self, x, y=2):
            self.x, self.y = x, y

    class Beta(Alpha):
        pass

    out = {}
    out["register_returns_class"] = _outcome(lambda: reg.register_module(Alpha) is Alpha)
    out["register_second"] = _outcome(lambda: reg.register_module(Beta) is Beta)
    out["repr"] = _outcome(lambda: repr(reg))
    out["name"] = _outcome(lambda: reg.name)
    out["module_dict_keys"] = _outcome(lambda: list(reg.module_dict))
    out["get_known"] = _outcome(lambda: reg.get("Alpha") is Alpha)
    out["get_unknown"] = _outcome(lambda: reg.get("Nope"))
    out["register_duplicate"] = _outcome(lambda: reg.register_module(Alpha))
    out["register_function"] = _outcome(lambda: reg.register_module(len))
    out["register_instance"] = _outcome(lambda: reg.register_module(Alpha(1)))

    def attrs(obj):
        return [type(obj).__name__, obj.x, obj.y]

    cfg = {"type": "Alpha", "x": 1}
    out["build_by_name"] = _outcome(lambda: attrs(build(cfg, reg)))
    out["cfg_untouched"] = _outcome(lambda: cfg == {"type": "Alpha", "x": 1})
    out["build_by_class"] = _outcome(lambda: attrs(build({"type": Beta, "x": 3, "y": 4}, reg)))
    out["defaults_fill_only_missing"] = _outcome(lambda: attrs(build({"type": "Alpha", "x": 1}, reg, {"x": 9, "y": 5})))
    out["defaults_none"] = _outcome(lambda: attrs(build({"type": "Alpha", "x": 1}, reg, None)))
    out["unknown_name"] = _outcome(lambda: build({"type": "Gamma"}, reg))
    out["type_is_int"] = _outcome(lambda: build({"type": 3}, reg))
    out["type_is_none"] = _outcome(lambda: build({"type": None}, reg))
    out["missing_type"] = _outcome(lambda: build({"x": 1}, reg))
    out["cfg_not_dict"] = _outcome(lambda: build([("type", "Alpha")], reg))
    out["defaults_not_dict"] = _outcome(lambda: build({"type": "Alpha", "x": 1}, reg, [("y", 1)]))
    return out


def fileio_cases(api, workdir):
    load, dump = api["load"], api["dump"]
    buffer = {"0": [], "600": [3, 1, 2], "1200": [3, 1, 2, 40, 7]}
    infos = [{"token": "t0", "gt_names": np.array(["car", "bus"]), "car_from_global": np.eye(4)},
             {"token": "t1", "gt_names": np.array([], dtype="<U1"), "car_from_global": np.arange(16.0).reshape(4, 4)}]
    plain = {"a": [1, 2.5, "x", None, True], "b": {"c": []}}
    p = lambda name: os.path.join(workdir, name)   # noqa: E731
    out = {}

    def dumped(obj, name, **kw):
        dump(obj, p(name), **kw)
        return _sha(open(p(name), "rb").read())

    out["json_buffer_bytes"] = _outcome(lambda: dumped(buffer, "b.json"))
    out["json_buffer_indent4_bytes"] = _outcome(lambda: dumped({"0": []}, "b0.json", indent=4))
    out["json_roundtrip"] = _outcome(lambda: load(p("b.json")) == buffer)
    out["json_from_Path"] = _outcome(lambda: load(Path(p("b.json"))) == buffer)
    out["json_to_str"] = _outcome(lambda: dump(buffer, file_format="json"))
    out["json_to_str_sorted"] = _outcome(lambda: dump(plain, file_format="json", sort_keys=True))
    out["json_fileobj"] = _outcome(lambda: load(io.StringIO('{"0": [], "5": [1]}'), file_format="json"))
    out["json_dump_fileobj"] = _outcome(lambda: (lambda f: (dump(buffer, f, file_format="json"), f.getvalue())[1])(io.StringIO()))
    out["pkl_infos_bytes"] = _outcome(lambda: dumped(infos, "i.pkl"))
    out["pkl_protocol_header"] = _outcome(lambda: list(open(p("i.pkl"), "rb").read()[:2]))
    out["pkl_roundtrip"] = _outcome(lambda: [[i["token"], i["gt_names"].tolist(), i["car_from_global"].tolist()] for i in load(p("i.pkl"))])
    out["pickle_extension"] = _outcome(lambda: dumped(buffer, "b.pickle"))
    out["pkl_to_str"] = _outcome(lambda: _sha(dump(buffer, file_format="pkl")))
    out["pkl_protocol_kwarg"] = _outcome(lambda: list(dump(buffer, file_format="pkl", protocol=4)[:2]))
    out["pkl_fileobj"] = _outcome(lambda: load(io.BytesIO(dump(buffer, file_format="pickle")), file_format="pkl") == buffer)
    out["yaml_plain_text"] = _outcome(lambda: dump(plain, file_format="yaml"))
    out["yaml_roundtrip"] = _outcome(lambda: (dump(plain, p("c.yml")), load(p("c.yml")) == plain)[1])
    out["load_unknown_extension"] = _outcome(lambda: load(p("x.txt")))
    out["load_no_extension_format"] = _outcome(lambda: load(io.StringIO("{}")))
    out["load_bad_file_type"] = _outcome(lambda: load(3, file_format="json"))
    out["load_missing_file"] = _outcome(lambda: type(load(p("missing.json"))))
    out["dump_no_file_no_format"] = _outcome(lambda: dump(buffer))
    out["dump_unknown_format"] = _outcome(lambda: dump(buffer, p("x.npy")))
    out["dump_bad_file_type"] = _outcome(lambda: dump(buffer, 3, file_format="json"))
    out["dump_returns_none"] = _outcome(lambda: dump(buffer, p("r.json")) is None)
    # base_selector.py:69-72 names the infos subset with str.replace on the extension
    out["explicit_format_wins"] = _outcome(lambda: (dump(buffer, p("weird.dat"), file_format="json"), load(p("weird.dat"), file_format="json") == buffer)[1])
    for k, v in out.items():   # the temp directory is not part of the outcome
        if "args" in v:
            v["args"] = [a.replace(workdir, "<tmp>") for a in v["args"]]
    return out


def run_all(api, workdir):
    return {"registry": registry_cases(api), "fileio": fileio_cases(api, workdir)}


CLI_ARGV = {   # scenario id -> argv after the program name (tools/active_select.py:22-63)
    "defaults": ["--config", "c.py"],
    "budget_and_workdir": ["--config", "c.py", "--budget", "600", "--work_dir", "/w"],
    "pred_flag": ["--config", "c.py", "--pred"],
    "launcher_none": ["--config", "c.py", "--launcher", "none"],
    "launcher_bogus": ["--config", "c.py", "--launcher", "ssh"],
    "local_rank": ["--config", "c.py", "--local_rank", "3", "--gpus", "8"],
    "budget_not_int": ["--config", "c.py", "--budget", "many"],
    "txt_result": ["--config", "c.py", "--txt_result", "1", "--checkpoint", "x.pth"],
    "unknown_flag": ["--config", "c.py", "--resume"],
}
CLI_KEYS = ("config", "work_dir", "txt_result", "gpus", "budget", "pred", "launcher", "local_rank")


def cli_cases(parse_argv):
    """``parse_argv(argv) -> namespace``; outcome = the shared keys of the namespace and the LOCAL_RANK the parser
    exports when the launcher has not set one (:60-61), or the ``SystemExit`` code of argparse."""
    import contextlib
    out = {}
    for name, argv in CLI_ARGV.items():
        saved = os.environ.pop("LOCAL_RANK", None)
        try:
            with contextlib.redirect_stderr(io.StringIO()):
                ns = vars(parse_argv(list(argv)))
            out[name] = {"ok": {k: ns[k] for k in CLI_KEYS}, "LOCAL_RANK": os.environ.get("LOCAL_RANK")}
        except SystemExit as e:
            out[name] = {"raises": "SystemExit", "args": [str(e.code)]}
        finally:
            os.environ.pop("LOCAL_RANK", None)
            if saved is not None:
                os.environ["LOCAL_RANK"] = saved
    return out
