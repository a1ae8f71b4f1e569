"""marius_b200 -- Blackwell-native (sm_100a) implementation of Marius's per-batch embedding hot path:
gather -> DistMult / ComplEx / dot scoring against chunk-shared negatives -> sparse Adagrad scatter-add.

Layers (bottom up):
  include/marius_b200.h + marius_b200/csrc/*.cu   hand-written CUDA behind a C ABI   (libmarius_b200.so)
  marius_b200/csrc/host/*.cpp                      C++/libtorch adapters keeping Marius's Storage / EdgeDecoder / Batch / Model surface
  marius_b200.ops                                  Python entry points over the C ABI (torch tensors = device memory only)
There is no CPU or eager fallback anywhere in this package: `marius_b200.ops` raises ImportError when
libmarius_b200.so has not been built (`python -m marius_b200.build`), and every op raises on a non-zero status.
"""
import importlib

__all__ = ["ops", "host", "build", "MariusB200Error"]


def __getattr__(name):  # lazy so that `python -m marius_b200.build` works before the library exists
    if name in ("ops", "host", "build", "_lib"):
        return importlib.import_module("." + name, __name__)
    if name == "MariusB200Error":
        return importlib.import_module("._lib", __name__).MariusB200Error
    raise AttributeError(name)
