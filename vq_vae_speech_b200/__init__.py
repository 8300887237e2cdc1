"""Import shim: `import vq_vae_speech_b200` resolves to the hyphenated package directory `vq-vae-speech_b200/`."""
import importlib.util
import os
import sys

_REAL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'vq-vae-speech_b200')
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_REAL, '__init__.py'),
                                               submodule_search_locations=[_REAL])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
