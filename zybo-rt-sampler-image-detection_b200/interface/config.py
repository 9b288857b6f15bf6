"""`interface.config` -- the constants module the reference generates from config.json.

The reference writes PC/interface/config.py with build_config.py at *build* time
(PC/src/build_config.py:16-72); here the same names are produced at *import* time
from config.json (sizes are run-time values for the CUDA library).  Every key of
the "general", "python" and "c" sections becomes a module attribute; "expression"
entries are evaluated in the module namespace exactly like the generated file
would; "imports" are imported.

The file read is $BF_CONFIG_JSON if set, else <package>/src/config.json.  A
reference PC/src/config.json can be used unchanged.
"""
import importlib
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG_PATH = os.environ.get("BF_CONFIG_JSON",
                             os.path.join(os.path.dirname(_HERE), "src", "config.json"))


def _load(path):
    g = globals()
    with open(path) as f:
        data = json.load(f)
    for lib in data.get("python", {}).get("imports", []):
        g[lib] = importlib.import_module(lib)
    deferred = []
    for section in ("general", "python", "c"):
        for key, val in data.get(section, {}).items():
            if key == "imports":
                continue
            if key == "expression":
                deferred.extend(val.items())
            else:
                g[key] = val
    for key, expr in deferred:
        g[key] = eval(expr, g)          # noqa: S307 - same semantics as the generated module
    extra = data.get("b200", {})
    # directions.pyx:15-16 hard-codes these two (they ignore N_MICROPHONES/ACTIVE_ARRAYS)
    g["GEOMETRY_N_MICS"] = extra.get("GEOMETRY_N_MICS", 256)
    g["GEOMETRY_N_ARRAYS"] = extra.get("GEOMETRY_N_ARRAYS", 4)
    g["FIR_FUSED"] = extra.get("FIR_FUSED", -1)
    g.setdefault("MIC_GAIN", 128)
    g.setdefault("N_TAPS", 8)
    g.setdefault("SKIP_N_MICS", 1)


def reload(path=None, **overrides):
    """Re-read config.json (or `path`) and apply keyword overrides; returns this module.
    The reference needs `make clean && make` for this (PC/LÄS_DETTA.md:16-23)."""
    global CONFIG_PATH
    if path is not None:
        CONFIG_PATH = path
    _load(CONFIG_PATH)
    globals().update(overrides)
    if "N_SAMPLES" in overrides or "N_MICROPHONES" in overrides:
        globals()["BUFFER_LENGTH"] = globals()["N_SAMPLES"] * globals()["N_MICROPHONES"]
    import sys
    return sys.modules[__name__]


_load(CONFIG_PATH)
