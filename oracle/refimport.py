"""Import recipe for the live reference -- TEST INFRASTRUCTURE.

Used only by `oracle/gen_golden.py` and by tests that pin the oracle while the reference tree
is present (the build container).  Nothing here runs on the GPU box."""
import os
import sys
import types


def reference_root():
    for p in (os.environ.get("EZAUDIO_REF"), "/root/reference"):
        if p and os.path.isdir(os.path.join(p, "src", "models")):
            return p
    return None


def import_reference():
    """Returns a namespace with MaskDiT, DiTControlNet, OobleckDecoder or None if absent."""
    root = reference_root()
    if root is None:
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    for name, attrs in (("alias_free_torch", ["Activation1d"]), ("vector_quantize_pytorch", ["ResidualVQ", "FSQ"]),
                        ("audiotools", ["AudioSignal", "STFTParams"])):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m
    import contextlib
    import io
    import warnings
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.models.conditioners import MaskDiT
        from src.models.controlnet import DiTControlNet
        from src.modules.stable_vae.models.autoencoders import OobleckDecoder, OobleckEncoder
    ns.MaskDiT, ns.DiTControlNet, ns.OobleckDecoder, ns.OobleckEncoder = MaskDiT, DiTControlNet, OobleckDecoder, OobleckEncoder
    # the conditions package __init__ pulls librosa/julius; load the one file the shipped ControlNet config uses
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_energy", os.path.join(root, "src/models/conditions/energy.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ns.EnergyExtractor = mod.EnergyExtractor
    return ns


def build(cls, sd, **kw):
    import contextlib
    import copy
    import io
    import warnings
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = cls(**copy.deepcopy(kw))
    m.load_state_dict(sd, strict=True)
    return m.eval()
