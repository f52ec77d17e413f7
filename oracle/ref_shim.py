"""Import shim for the UNMODIFIED reference (test infrastructure only).

Only usable in the build container, where /root/reference is mounted.  It is
used by oracle/make_golden.py to generate the committed fixtures under
tests/golden/ and by tests that are skipped when /root/reference is absent.
Nothing on the product path, in `-m gpu` tests, smoke() or bench.py imports it.

The reference imports `matplotlib` (utils.py:1,6-7) and `ray` (datasets.py:3),
neither of which is installed; both are irrelevant to the hot path, so they are
replaced by inert stand-ins before `model`, `utils`, `datasets` are imported.
"""
import os
import sys
import types
import tempfile

REFERENCE_ROOT = os.environ.get("PWR_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "datasets.py"))


def _stub_modules():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        fm = types.ModuleType("matplotlib.font_manager")

        class _FontProperties:
            def __getattr__(self, name):
                return lambda *a, **k: None

        fm.FontProperties = _FontProperties
        mpl.rc = lambda *a, **k: None
        mpl.pyplot = plt
        mpl.font_manager = fm
        plt.cm = types.SimpleNamespace(jet=None)
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.font_manager"] = fm
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")
        ray.remote = lambda obj: obj
        ray.init = lambda *a, **k: None
        ray.shutdown = lambda *a, **k: None
        ray.get = lambda x: x
        sys.modules["ray"] = ray
    if "torchvision" not in sys.modules:
        try:
            import torchvision  # noqa: F401
        except Exception:
            sys.modules["torchvision"] = types.ModuleType("torchvision")


def load():
    """Return the reference's (model, utils, datasets) modules, unmodified."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub_modules()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        utils = importlib.import_module("utils")
        model = importlib.import_module("model")
        datasets = importlib.import_module("datasets")
    return model, utils, datasets


def make_synthetic_dataset(datasets, samples, fx, fy, halfu, halfv, cube, joints,
                           test_only=False, msra_com=False, augment=False):
    """Drive the reference's own HandDataset.process_single_data on in-memory
    synthetic samples: `samples` is a list of dicts(frame, uvd, com, cube).
    Only build_data()/load_from_text() are overridden (they are file I/O)."""
    tmp = tempfile.mkdtemp(prefix="pwr_ref_")
    with open(os.path.join(tmp, "train.txt"), "w") as f:
        f.write("\n".join(str(i) for i in range(len(samples))))

    class _Synthetic(datasets.HandDataset):
        def build_data(self):
            return

        def decode_line_txt(self, string):
            return string.strip(), None

        def load_from_text(self, text):
            s = samples[int(text.strip())]
            if msra_com:
                return s["frame"], s["uvd"], None, None
            return s["frame"], s["uvd"], s["com"], s["cube"]

    # augment: train.py's defaults (rotation, scale, shift on; flip off), train.py:35-38
    return _Synthetic(fx, fy, halfu, halfv, tmp, 1.5, 128, 7, 64, test_only,
                      augment, augment, augment, False, cube, joints)
