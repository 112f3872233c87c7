"""B200-native (sm_100a) implementation of the audio->pose inference-and-evaluation hot path of
Xukai-UoA/Audio-to-Motion-Generation, behind the reference's own Python call surface.

The directory name carries the upstream repository name (hyphens), so import it by string::

    import importlib
    a2m = importlib.import_module("audio-to-motion-generation_b200")
    a2m.install_dropin()                       # optional: expose the reference's module names
    from pose_video.mel_features import log_mel_spectrogram
    from real_motion_model import SelfAttention_G
    from motion_evaluation import compute_pck

Everything numerical runs in hand-written CUDA kernels inside ``liba2m_b200.so`` (C ABI in
``include/a2m_b200.h``); there is no CPU, Triton or eager-PyTorch fallback.
"""
import importlib
import sys

__version__ = "0.1.0"

_DROPIN = {
    "pose_video": ".pose_video",
    "pose_video.mel_features": ".pose_video.mel_features",
    "pose_video.audio_repr": ".pose_video.audio_repr",
    "motion_evaluation": ".motion_evaluation",
    "normalization_tools": ".normalization_tools",
    "model_layers": ".model_layers",
    "real_motion_model": ".real_motion_model",
}


# registered only when asked for by name: "pats" is a large package of the reference (data loaders, skeleton, ...) of
# which only the audio feature extraction has a B200 implementation -- shadowing it wholesale would hide the rest
_DROPIN_ON_REQUEST = {
    "pats.data_loading.audio": ".pats_audio",
}


def install_dropin(names=None):
    """Register this package's modules under the reference's top-level module names, so code written
    against the reference (``from pose_video.mel_features import ...``) runs on the B200 path."""
    installed = {}
    table = dict(_DROPIN)
    table.update({k: v for k, v in _DROPIN_ON_REQUEST.items() if names is not None and k in names})
    for public, relative in table.items():
        if names is not None and public not in names:
            continue
        try:
            mod = importlib.import_module(relative, __name__)
        except ModuleNotFoundError as e:
            if e.name and e.name.endswith(relative.lstrip(".")):
                continue             # module not built yet in this round
            raise
        sys.modules[public] = mod
        installed[public] = mod
    return installed


def load_library():
    """Load liba2m_b200.so now (raises ImportError with build instructions if it is missing)."""
    from . import _cabi
    return _cabi.lib()
