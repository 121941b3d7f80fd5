"""Import the live reference (read-only checkout) for fixture generation.

Only usable where ``/root/reference`` exists (the build container); the GPU box
has no such path, so nothing on the ``-m gpu`` / smoke / bench paths may call
this.  Two third-party modules the reference imports but never executes on the
entity-alignment path are absent from the image (POT ``ot`` and ``torchtext``,
SURVEY.md §8c); empty stand-ins are registered so the modules import.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GNN_MTL_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "layers"))


def _register_stand_ins():
    if "ot" not in sys.modules:
        ot = types.ModuleType("ot")
        ot.gromov = types.ModuleType("ot.gromov")
        sys.modules["ot"] = ot
        sys.modules["ot.gromov"] = ot.gromov
    if "torchtext" not in sys.modules:
        tt = types.ModuleType("torchtext")
        ttd = types.ModuleType("torchtext.data")
        for name in ("Dataset", "BucketIterator", "Field", "Example"):
            setattr(ttd, name, type(name, (), {}))
        tt.data = ttd
        sys.modules["torchtext"] = tt
        sys.modules["torchtext.data"] = ttd


class _RefModules:
    """Lazy handle: ``ref.layers``, ``ref.ot_loss``, ``ref.eval_utils`` …"""

    _names = {
        "layers": "layers.layers",
        "encoders": "models.encoders",
        "decoders": "models.decoders",
        "ot_loss": "utils.ot_loss",
        "eval_utils": "utils.eval_utils",
        "data_utils": "utils.data_utils",
        "sinkhorn_loss": "SinkhornOT.sinkhorn_loss",
        "cderivation": "SinkhornOT.cderivation",
        "models_ea": "models.models_ea",
        "iterative_projection": "SinkhornOT.iterative_projection",
    }

    def __getattr__(self, key):
        if key not in self._names:
            raise AttributeError(key)
        mod = load(self._names[key])
        setattr(self, key, mod)
        return mod


def load(dotted: str):
    """Import ``dotted`` from the reference tree without polluting the caller's
    package namespace (the reference uses top-level names ``utils``, ``models``
    that would shadow nothing in this repo, but keep it contained anyway)."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _register_stand_ins()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module(dotted)


ref = _RefModules()
