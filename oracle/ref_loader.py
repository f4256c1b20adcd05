"""TEST INFRASTRUCTURE ONLY.  Loader for the unmodified reference simulator.

Imports /root/reference/libs_m/ct/src/data/cancer_sim/cancer_simulation.py as a free
standing module.  Its imports (cancer_simulation.py:15-23) include matplotlib and seaborn,
which are absent here and only used for plotting, so empty stub modules are planted in
``sys.modules`` first.  The file is *not* imported through the ``src`` package because
``src/data/__init__.py`` pulls in jax.

Only usable in the build container: /root/reference does not exist on the GPU box, so
nothing that runs there (``-m gpu`` tests, smoke(), bench.py) may call this.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("INSITE_REFERENCE_ROOT", "/root/reference")
REF_SIM = os.path.join(REF_ROOT, "libs_m/ct/src/data/cancer_sim/cancer_simulation.py")


def reference_available() -> bool:
    return os.path.isfile(REF_SIM)


def _plant_stubs():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "seaborn" not in sys.modules:
        sns = types.ModuleType("seaborn")
        sns.set = lambda *a, **k: None
        sys.modules["seaborn"] = sns


_cached = None


def load_reference_sim():
    """Return the reference ``cancer_simulation`` module (tqdm silenced)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference simulator not found at {REF_SIM}")
    _plant_stubs()
    spec = importlib.util.spec_from_file_location("_insite_reference_cancer_simulation", REF_SIM)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.tqdm = lambda it, **kw: it      # silence progress bars; iteration unchanged
    _cached = mod
    return mod


REF_CONT = os.path.join(REF_ROOT, "libs_m/ct/src/data/continuous/continuous.py")


def load_reference_continuous():
    """The reference's EQ_5 simulators (continuous/continuous.py), unmodified.  Its only extra import is
    ``src.data.pkpd.pkpd_simulation.Equation`` (:21), whose package pulls in jax: a stub module with the same IntEnum
    (pkpd_simulation.py:51-60) is planted instead."""
    import enum
    if not os.path.isfile(REF_CONT):
        raise FileNotFoundError(f"reference continuous simulator not found at {REF_CONT}")
    _plant_stubs()
    if "src.data.pkpd.pkpd_simulation" not in sys.modules:
        class Equation(enum.IntEnum):
            EQ_4_A = 1; EQ_4_B = 2; EQ_4_C = 3; EQ_4_D = 4; EQ_5_A = 5; EQ_5_B = 6; EQ_5_C = 7; EQ_5_D = 8; EQ_4_M = 9
        for name in ("src", "src.data", "src.data.pkpd"):
            sys.modules.setdefault(name, types.ModuleType(name))
        m = types.ModuleType("src.data.pkpd.pkpd_simulation")
        m.Equation = Equation
        sys.modules["src.data.pkpd.pkpd_simulation"] = m
    spec = importlib.util.spec_from_file_location("_insite_reference_continuous", REF_CONT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.tqdm = lambda it, **kw: it
    return mod
