"""In-process shims that let the UNMODIFIED reference package import in this image.  Test infrastructure.

Nothing under the reference tree is edited: missing optional dependencies are stubbed in ``sys.modules``, the Pillow
constant the reference still uses is aliased, and the ImageNet download is replaced by a random-init constructor.
The reference is looked up in ``$FP_REFERENCE_DIR``, ``baseline/_ref`` (scripts/install_reference.sh; travels to the GPU
box, never committed) and ``/root/reference`` (build container only)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for p in (os.environ.get('FP_REFERENCE_DIR'), os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if p and os.path.isfile(os.path.join(p, 'footprints', 'network.py')):
            return p
    return None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference root after making ``import footprints...`` work; None when no checkout is available."""
    ref = find_reference()
    if ref is None:
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from PIL import Image
    if not hasattr(Image, 'ANTIALIAS'):            # removed in Pillow 10; predict_simple.py:42 still names it
        Image.ANTIALIAS = Image.LANCZOS
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        import numpy as np

        def get_cmap(name=None, lut=None):
            return lambda x: np.stack([x, x, x, np.ones_like(x)], -1)
        plt = _stub('matplotlib.pyplot', get_cmap=get_cmap)
        _stub('matplotlib', pyplot=plt)
    try:
        import tensorboardX  # noqa: F401
    except ImportError:
        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, name):
                return lambda *a, **k: None
        _stub('tensorboardX', SummaryWriter=SummaryWriter)
    try:
        import skimage  # noqa: F401
    except ImportError:
        meas = _stub('skimage.measure', label=lambda *a, **k: None)
        _stub('skimage', measure=meas)
    import torchvision
    import footprints.network as fn
    fn.resnet34 = lambda pretrained=True: torchvision.models.resnet34(weights=None)     # no ImageNet download
    return ref
