"""Import the reference's own hot-path modules UNMODIFIED from /root/reference, over oracle/shims.

Only usable in the build container (the GPU box has no /root/reference): used by make_golden.py
and by the not-gpu tests that are skipped when the reference tree is absent."""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("RF_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "train_flux", "flux"))


def load():
    """Returns a namespace with the reference modules: block, transformer, generate, lora_controller."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_REPO, _SHIMS, os.path.join(REFERENCE_ROOT, "train_flux")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # `flux` package of the reference: import submodules individually (its __init__ is empty)
    mods = {}
    for name in ("lora_controller", "block", "transformer", "condition", "generate"):
        try:
            mods[name] = importlib.import_module(f"flux.{name}")
        except Exception as e:  # condition.py needs cv2 etc.; not required for the oracle
            if name in ("block", "transformer", "lora_controller", "generate"):
                raise
            mods[name] = e

    class NS:
        pass
    ns = NS()
    ns.__dict__.update(mods)
    return ns


def load_tts_utils():
    """tts/utils.py of the reference (get_noises, prepare_latents_for_flux)."""
    for p in (_REPO, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    spec = importlib.util.spec_from_file_location("ref_tts_utils", os.path.join(REFERENCE_ROOT, "tts", "utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m
