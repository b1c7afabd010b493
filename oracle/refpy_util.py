"""oracle/refpy_util.py — TEST INFRASTRUCTURE (not product code): the reference's own Python surface, importable where /root/reference is absent.

`make -C oracle pyref` leaves in oracle/_ref/ (git-ignored build outputs that travel to the GPU box):
  MARLEnv.so          the reference's pybind11 module (cpp/bindings.cpp over the unmodified sources + the RNG shim)
  refpy/{env,utils,cpp_backend}.pyc.bin   byte-compiled copies of the reference's env.py / utils.py / cpp_backend.py

Only tests/ and bench.py's CPU-baseline leg import this module; the product package never does.

`load_reference_env(backend)` executes the reference's env.py with its `import cpp_backend` bound to
  * "MARLEnv": the reference's own cpp_backend.py, which imports MARLEnv.so  -> the reference, end to end;
  * a module object: e.g. this repo's marl_traffic_intersection_b200.cpp_backend -> the reference's env.py on the CUDA stepper.
"""
from __future__ import annotations

import ctypes
import importlib.machinery
import importlib.util
import os
import sys

REF_OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFPY = os.path.join(REF_OUT, "refpy")


def have_pyref() -> bool:
    return os.path.exists(os.path.join(REF_OUT, "MARLEnv.so")) and os.path.exists(os.path.join(REFPY, "env.pyc.bin"))


def _load_pyc(name: str, alias: str):
    path = os.path.join(REFPY, name + ".pyc.bin")
    loader = importlib.machinery.SourcelessFileLoader(alias, path)
    spec = importlib.util.spec_from_loader(alias, loader)
    mod = importlib.util.module_from_spec(spec)
    mod.__file__ = path
    loader.exec_module(mod)
    return mod


def marlenv_module():
    if REF_OUT not in sys.path:
        sys.path.insert(0, REF_OUT)
    import MARLEnv  # noqa: N813
    return MARLEnv


def marlenv_seed(seed: int, env_id: int, tick: int):
    """Position the traffic stream the shimmed TrafficFlow.cpp reads (oracle/ref_pystream.cpp) before an env.step()."""
    lib = ctypes.CDLL(marlenv_module().__file__)
    lib.isxpy_seed.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
    lib.isxpy_seed.restype = None
    lib.isxpy_seed(int(seed), int(env_id), int(tick))


def load_reference_env(backend):
    """Returns the module object of the reference's env.py, executed with `cpp_backend` = backend."""
    saved = {k: sys.modules.get(k) for k in ("cpp_backend", "utils")}
    try:
        sys.modules["utils"] = _load_pyc("utils", "utils")
        if backend == "MARLEnv":
            marlenv_module()
            sys.modules["cpp_backend"] = _load_pyc("cpp_backend", "cpp_backend")
            tag = "ref_env_over_marlenv"
        else:
            sys.modules["cpp_backend"] = backend
            tag = "ref_env_over_" + backend.__name__.replace(".", "_")
        return _load_pyc("env", tag)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
