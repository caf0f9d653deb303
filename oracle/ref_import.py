"""Import harness for the UNMODIFIED reference (test infrastructure only).

Only usable where /root/reference exists (the build container).  It fabricates
empty stub modules for the front-end pip packages the reference imports
transitively but never calls on the arithmetic path, then builds
the three reference model classes from the reference's own YAML configs.

Nothing under conan_b200/ may import this file.  It is used by
oracle/make_golden.py (fixture generation) and by tests that are skipped when
/root/reference is absent.
"""
import contextlib
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("CONAN_REFERENCE_ROOT", "/root/reference")

_STUBS = {
    "librosa", "pyloudnorm", "skimage", "matplotlib", "h5py", "textgrid", "webrtcvad",
    "resemblyzer", "g2p_en", "nltk", "parselmouth", "pycwt", "torchdyn", "chardet",
    "soundfile", "resampy", "pyworld", "torchcrepe",
}


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        sys.modules[m.__name__] = m
        setattr(self, k, m)
        return m


class _Finder:
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in _STUBS or name.startswith("modules.parallel_wavegan"):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "modules", "Conan"))


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REF_ROOT)
    _installed = True


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def build_reference_models():
    """Returns (hp, conan, emformer_distill, hifigan_generator, voc_hp) built by the
    reference's own constructors from its own YAMLs (random init, eval mode)."""
    install()
    import torch
    with _cwd(REF_ROOT):
        from utils.commons.hparams import set_hparams
        hp = set_hparams(config="egs/conan_emformer.yaml", print_hparams=False)
        voc_hp = set_hparams(config="egs/hifi_16k320_shuffle.yaml", print_hparams=False,
                             global_hparams=False)
        from modules.Conan.Conan import Conan
        from modules.Emformer.emformer import EmformerDistillModel
        from modules.vocoder.hifigan.hifigan_causal import HifiGanGenerator
        torch.manual_seed(1234)
        conan = Conan(0, hp).eval()
        emf = EmformerDistillModel(hp, output_dim=100).eval()
        voc = HifiGanGenerator(voc_hp).eval()
    return hp, conan, emf, voc, voc_hp
